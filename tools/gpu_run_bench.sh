set -x
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/ncu_step.py 256 8 > gpurun_out/ncu_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_r01.csv python tools/ncu_step.py 256 8 > gpurun_out/ncu.log 2>&1; echo "ncu exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/summary.txt; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
