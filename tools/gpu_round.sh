#!/bin/bash
# One gpurun call: [tests] [bench] probes + ncu launch list + ncu full captures.  Outputs under gpurun_out/ (<64 MiB!).
# usage: bash tools/gpu_round.sh [tests] [bench] [probe] [launches] [ncu_s2d] [ncu_conv] [convs]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for what in "$@"; do
case $what in
tests)
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log;;
bench)
  timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err;;
benchref)
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
  cat gpurun_out/bench_ref.json;;
convs)
  timeout 300 python tools/profile_convs.py > gpurun_out/profile_convs.log 2>&1; cat gpurun_out/profile_convs.log;;
probe)
  timeout 120 tools/bin/pipe_probe > gpurun_out/pipe_probe.log 2>&1; cat gpurun_out/pipe_probe.log;;
launches)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-autotune > gpurun_out/bench_ncu.log 2>&1;;
ncu_s2d)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2d_fused -s 3 -c 1 -o gpurun_out/prof_s2d -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_s2d.log 2>&1
  ncu -i gpurun_out/prof_s2d.ncu-rep --page raw --csv > gpurun_out/prof_s2d_raw.csv 2>/dev/null;;
ncu_kb)
  # the fused KB block on BASELINE config 2 (tools/bench_kb.py launches nothing else from the tensor-core engine)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_v2 -s 6 -c 1 -o gpurun_out/prof_kb2 -f python tools/bench_kb.py > gpurun_out/ncu_kb.log 2>&1
  ncu -i gpurun_out/prof_kb2.ncu-rep --page raw --csv > gpurun_out/prof_kb2_raw.csv 2>/dev/null;;
ncu_conv)
  # one full forward = 34 conv launches; skip the 3 warm-up forwards, capture a spread of layers
  timeout 900 ncu --set full --clock-control none -k regex:conv_umma -s 102 -c 34 -o gpurun_out/prof_conv -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
  ncu -i gpurun_out/prof_conv.ncu-rep --page raw --csv > gpurun_out/prof_conv_raw.csv 2>/dev/null
  rm -f gpurun_out/prof_conv.ncu-rep;;
esac
done
du -sh gpurun_out; ls -la gpurun_out
