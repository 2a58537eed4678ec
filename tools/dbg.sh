mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_s2d_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/s2d_tma.log 2>&1; tail -5 gpurun_out/s2d_tma.log
timeout 120 python tools/bench_s2d.py 8 KITTI 2 2>&1 | tail -2
timeout 120 python tools/bench_s2d.py 8 VOID 2 2>&1 | tail -2
timeout 120 python tools/bench_s2d.py 1 KITTI 1 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:s2d_fused -s 6 -c 1 -o gpurun_out/prof_s2d_v2 -f python tools/bench_s2d.py 8 KITTI 2 > gpurun_out/ncu_s2d_v2.log 2>&1
ncu -i gpurun_out/prof_s2d_v2.ncu-rep --page raw --csv > gpurun_out/prof_s2d_v2_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_s2d_v2.ncu-rep --page source --csv > gpurun_out/prof_s2d_v2_source.csv 2>/dev/null
