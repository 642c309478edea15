# usage: bash tools/gpu_ab.sh  -- GPU parity tests, bench with PDL, bench without PDL (A/B)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
PTTS_B200_NO_MEGA=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; echo "bench nomega exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("bench","bench_nopdl"):
    try: d=json.load(open(f'gpurun_out/{f}.json'))
    except Exception as ex: print(f,"unreadable",ex); continue
    print("==",f,"value",round(d["value"],1),"ms/step",round(d["ms_per_step"],1),"e2e",round(d["e2e"]["value"],1))
    print("roofline",d.get("roofline"))
    for k,v in d.get("kernel_classes",{}).items(): print(k,v)
    print("batch1",d.get("batch1"))
PY
