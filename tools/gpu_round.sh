#!/bin/bash
# One gpurun call: a list of steps.  Outputs under gpurun_out/ (<64 MiB!).
# usage: bash tools/gpu_round.sh [tests] [bench] [benchvoid] [benchref] [incumbent] [sanitize] [probe] [launches] [ncu_s2d] [ncu_kb] [micro] [layers]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for what in "$@"; do
case $what in
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -8 gpurun_out/pytest_gpu.log;;
tests_s2d)
  timeout 900 python -m pytest tests/test_s2d_gpu.py -m gpu -x -q > gpurun_out/pytest_s2d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s2d.log
  tail -8 gpurun_out/pytest_s2d.log;;
bench)
  timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err;;
bench_eager)
  timeout 600 python bench.py --no-graph --no-cpu-baseline --no-gpu-incumbent > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; echo "bench rc=$?"
  cat gpurun_out/bench_eager.json; tail -3 gpurun_out/bench_eager.err;;
benchvoid)
  timeout 600 python bench.py --workload void > gpurun_out/bench_void.json 2> gpurun_out/bench_void.err; echo "bench void rc=$?"
  cat gpurun_out/bench_void.json; tail -3 gpurun_out/bench_void.err;;
benchref)
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
  cat gpurun_out/bench_ref.json;;
incumbent)
  timeout 600 python bench.py --impl torch-gpu --steps 20 > gpurun_out/bench_torch_gpu.json 2> gpurun_out/bench_torch_gpu.err; echo "torch-gpu rc=$?"
  cat gpurun_out/bench_torch_gpu.json; tail -3 gpurun_out/bench_torch_gpu.err;;
sanitize)
  for tool in memcheck racecheck synccheck initcheck; do
    timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/sanitize_$tool.log python tools/sanitize_target.py > gpurun_out/sanitize_$tool.out 2>&1
    echo "sanitize $tool rc=$?"; tail -3 gpurun_out/sanitize_$tool.log; tail -1 gpurun_out/sanitize_$tool.out
  done;;
micro)
  timeout 300 python tools/bench_s2d.py > gpurun_out/micro_s2d.log 2>&1; cat gpurun_out/micro_s2d.log
  timeout 300 python tools/bench_kb.py > gpurun_out/micro_kb.log 2>&1; cat gpurun_out/micro_kb.log;;
layers)
  timeout 300 python tools/profile_layers.py > gpurun_out/layers.log 2>&1; cat gpurun_out/layers.log;;
probe)
  timeout 120 tools/bin/pipe_probe > gpurun_out/pipe_probe.log 2>&1; cat gpurun_out/pipe_probe.log;;
launches)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-incumbent --no-autotune --no-graph --no-shard-check > gpurun_out/bench_ncu.log 2>&1;;
ncu_s2d)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2d_ -s 3 -c 1 -o gpurun_out/prof_s2d -f python tools/bench_s2d.py > gpurun_out/ncu_s2d.log 2>&1
  ncu -i gpurun_out/prof_s2d.ncu-rep --page raw --csv > gpurun_out/prof_s2d_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_s2d.ncu-rep --page source --csv > gpurun_out/prof_s2d_source.csv 2>/dev/null;;
ncu_kb)
  # the fused KB block on BASELINE config 2 (tools/bench_kb.py launches nothing else from the tensor-core engine)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_v2 -s 6 -c 1 -o gpurun_out/prof_kb2 -f env KBNET_AUTOTUNE=0 KBNET_V2_PLAN=2,62,1,0 python tools/bench_kb.py 8 2 nchw > gpurun_out/ncu_kb.log 2>&1
  ncu -i gpurun_out/prof_kb2.ncu-rep --page raw --csv > gpurun_out/prof_kb2_raw.csv 2>/dev/null;;
esac
done
du -sh gpurun_out; ls -la gpurun_out
