# A/B of the persistent kernel's experiment flags: bash tools/gpu_flags.sh "0 1 2 3"
for f in $1; do echo "== flags $f"; PTTS_B200_MEGA_FLAGS=$f python tools/mega_trace.py 256 100 2>&1 | grep -E "^ +(0|3|23|24|25|26|27|28|29) |gemm_tc"; done
