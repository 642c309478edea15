set -x
timeout 300 python tools/ncu_step.py 256 4 0 > gpurun_out/ncu_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:k_attn_decode -s 500 -c 2 -o gpurun_out/r01_attn_decode python tools/ncu_step.py 256 4 0 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?" > gpurun_out/summary.txt
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:gemm_tc_kernel -s 2500 -c 6 -o gpurun_out/r01_gemm_flow python tools/ncu_step.py 256 4 0 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; ls -la gpurun_out/*.ncu-rep
