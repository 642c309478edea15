# A/B of the attention prefetch flags on the phase trace
mkdir -p gpurun_out
timeout 600 python tools/parity_long.py 48 6 2>&1 | tail -1
for fl in 0 1 4 5; do
  echo "== PTTS_B200_MEGA_FLAGS=$fl"
  PTTS_B200_MEGA_FLAGS=$fl timeout 600 python tools/mega_trace.py 256 100 2> gpurun_out/mega_trace_f$fl.txt | tail -1 | cut -c1-220
  grep -E "ATTN|attention unit" gpurun_out/mega_trace_f$fl.txt | head -8
done
