# same-box A/B: bench with the tensor-core attention (PTTS_B200_ATTN_TC=1) vs the scalar one, interleaved twice
mkdir -p gpurun_out
for i in 1 2; do
  PTTS_B200_ATTN_TC=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --skip-mimi-only > gpurun_out/bench_tc$i.json 2> gpurun_out/bench.err || tail -3 gpurun_out/bench.err
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --skip-mimi-only > gpurun_out/bench_sc$i.json 2> gpurun_out/bench.err || tail -3 gpurun_out/bench.err
done
python - <<'PY'
import json
for f in ("tc1","sc1","tc2","sc2"):
    d=json.load(open(f'gpurun_out/bench_{f}.json'))
    kc=d.get("kernel_classes",{})
    print(f,"ms/step",round(d["ms_per_step"],2),"clk",d["clocks"]["sm_mhz"],"gemm_flow",kc.get("gemm_tc_flowlm",{}).get("ms"),"attn",kc.get("attn_decode",{}).get("ms"),"row",kc.get("rownorm",{}).get("ms"),"mimi",kc.get("gemm_tc_mimi",{}).get("ms"))
PY
