# usage: bash tools/gpu_attn.sh -- parity tests + long-horizon parity + phase trace + bench for the current build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python tools/parity_long.py 48 6 2>&1 | tail -3
timeout 600 python tools/mega_trace.py 256 100 2> gpurun_out/mega_trace.txt | tail -1; grep -E "mega trace|ATTN|attention unit" gpurun_out/mega_trace.txt | head -12
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print("value",round(d["value"],1),"ms/step",round(d["ms_per_step"],1),"e2e",round(d["e2e"]["value"],1))
print("roofline",d.get("roofline"))
for k,v in d.get("kernel_classes",{}).items(): print(k,v)
PY
echo "== B=128 trace"
timeout 600 python tools/mega_trace.py 128 100 2> gpurun_out/mega_trace_b128.txt | tail -1 | cut -c1-200; grep -E "ATTN|attention unit" gpurun_out/mega_trace_b128.txt | head -8
