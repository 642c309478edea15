# usage: bash tools/gpu_ncu_mimi.sh TAG -- ncu --set full captures of the Mimi kernels (one launch each of the fused 64-channel block,
# the stage-3 ConvTranspose GEMM and a transformer GEMM) on a 64 x 100-frame decode
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 300 python tools/mimi_once.py > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:resblock_kernel -s 6 -c 3 -o gpurun_out/${TAG}_mimi_resblock -f python tools/mimi_once.py > gpurun_out/ncu_mimi.log 2>&1; echo "ncu resblock exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 24 -c 12 -o gpurun_out/${TAG}_mimi_gemm -f python tools/mimi_once.py >> gpurun_out/ncu_mimi.log 2>&1; echo "ncu gemm exit $?"
ls -la gpurun_out/${TAG}_mimi_*.ncu-rep
