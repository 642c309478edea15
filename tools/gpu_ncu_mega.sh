# usage: bash tools/gpu_ncu_mega.sh TAG [F] -- one ncu --set full capture of k_flow_mega (B=256, F frames), source view included
TAG=${1:-r01h}; F=${2:-48}
mkdir -p gpurun_out
timeout 300 python tools/ncu_step.py 256 $F 0 > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:k_flow_mega -c 1 -o gpurun_out/${TAG}_flow_mega -f python tools/ncu_step.py 256 $F 0 > gpurun_out/ncu_mega.log 2>&1
echo "ncu mega exit $?"; tail -3 gpurun_out/ncu_mega.log; ls -la gpurun_out/${TAG}_flow_mega.ncu-rep
