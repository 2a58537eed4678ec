set -u
mkdir -p gpurun_out
bash tools/gpu_round.sh launches ncu_s2d ncu_kb micro > gpurun_out/evid1.log 2>&1
timeout 200 python tools/bench_s2d.py 8 VOID >> gpurun_out/micro_s2d.log 2>&1
timeout 200 python tools/bench_small_conv.py > gpurun_out/micro_small_conv.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:conv_mma -s 2 -c 1 -o gpurun_out/prof_cm348 -f python tools/bench_small_conv.py 3 48 > /dev/null 2>&1
ncu -i gpurun_out/prof_cm348.ncu-rep --page raw --csv > gpurun_out/prof_cm348_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:conv_mma -s 2 -c 1 -o gpurun_out/prof_cm816 -f python tools/bench_small_conv.py 8 16 > /dev/null 2>&1
ncu -i gpurun_out/prof_cm816.ncu-rep --page raw --csv > gpurun_out/prof_cm816_raw.csv 2>/dev/null
rm -f gpurun_out/prof_cm348.ncu-rep gpurun_out/prof_cm816.ncu-rep
bash tools/gpu_round.sh bench benchvoid benchref incumbent > gpurun_out/evid2.log 2>&1
tail -5 gpurun_out/evid2.log
ls -la gpurun_out | head -50
