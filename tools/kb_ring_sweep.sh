mkdir -p gpurun_out
export KBNET_AUTOTUNE=0 KBNET_V2_PLAN=2,62,1,0
for cfg in "nchw 24576 3,2,2" "nchw 24576 2,2,3" "nchw 6144 3,2,8" "nchw 6144 2,2,8" "nchw 6144 3,2,4" "nchw 18432 3,2,3" "nchw 18432 2,2,4" "of 24576 3,2,2" "of 24576 2,2,3" "of 6144 2,3,4" "of 6144 2,2,8" "of 18432 2,2,4" "of 6144 2,3,6"; do
  set -- $cfg
  echo "== mode $1 WCAP $2 RINGS $3"
  KBNET_V2_WCAP=$2 KBNET_V2_RINGS=$3 KBNET_V2_DEBUG=1 timeout 100 python tools/bench_kb.py 8 2 $1 2>&1 | grep -v "^\[conv v2\] geo [013]" | sed -e 's/algorithmic.*of measured/=/' | sort -u | tail -2
done
