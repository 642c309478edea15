# usage: bash tools/gpu_final.sh TAG -- tests, bench (own arm + reference arm), ncu launch list + full capture of k_flow_mega
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref.err; echo "ref exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/ncu_step.py 256 100 1 > gpurun_out/ncu_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_${TAG}.csv python tools/ncu_step.py 256 100 1 > gpurun_out/ncu.log 2>&1; echo "ncu list exit $?" >> gpurun_out/summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:k_flow_mega -c 1 -o gpurun_out/${TAG}_flow_mega -f python tools/ncu_step.py 256 100 0 > gpurun_out/ncu_mega.log 2>&1; echo "ncu mega exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_${TAG}.json
