mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
timeout 600 python tools/parity_long.py 48 6 2>&1 | tail -2
for i in 1 2; do
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --skip-mimi-only > gpurun_out/bench_pt$i.json 2> gpurun_out/bench.err || tail -3 gpurun_out/bench.err
  PTTS_B200_PF_ATTN_SCALAR=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --skip-mimi-only > gpurun_out/bench_ps$i.json 2> gpurun_out/bench.err || tail -3 gpurun_out/bench.err
done
python - <<'PY'
import json
for f in ("pt1","ps1","pt2","ps2"):
    d=json.load(open(f'gpurun_out/bench_{f}.json'))
    kc=d.get("kernel_classes",{})
    print(f,"ms/step",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"]),"clk",d["clocks"]["sm_mhz"],"gemm_flow",kc.get("gemm_tc_flowlm",{}).get("ms"),"attn",kc.get("attn_decode",{}).get("ms"),"row",kc.get("rownorm",{}).get("ms"),"mimi",kc.get("gemm_tc_mimi",{}).get("ms"))
PY
