# usage: bash tools/gpu_prof.sh TAG [B] [F] -- ncu launch list of one FlowLM+Mimi job, then full captures of the top kernels
TAG=${1:-r01}; B=${2:-256}; F=${3:-8}
mkdir -p gpurun_out
timeout 300 python tools/ncu_step.py $B $F 1 > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_${TAG}.csv python tools/ncu_step.py $B $F 1 > gpurun_out/ncu.log 2>&1; echo "ncu list exit $?" > gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:k_attn_decode -s 8 -c 2 -o gpurun_out/${TAG}_attn_decode -f python tools/ncu_step.py $B $F 0 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:gemm_tc_kernel -s 200 -c 8 -o gpurun_out/${TAG}_gemm_flow -f python tools/ncu_step.py $B $F 0 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; ls -la gpurun_out/*.ncu-rep
