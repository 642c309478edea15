# usage: bash tools/gpu_ab.sh TAG "flagsA flagsB ..." -- phase trace of the persistent step kernel under each PTTS_B200_MEGA_FLAGS value
# (same box, back to back), then a short bench with the default flags
TAG=${1:-ab}; FLAGS=${2:-"0"}
mkdir -p gpurun_out
for f in $FLAGS; do
  echo "== PTTS_B200_MEGA_FLAGS=$f"
  PTTS_B200_MEGA_FLAGS=$f timeout 600 python tools/mega_trace.py 256 100 2> gpurun_out/${TAG}_trace_f$f.txt | tail -1 | cut -c1-400
  grep -E " ATTN |attention" gpurun_out/${TAG}_trace_f$f.txt | head -8
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --skip-mimi-only --total-prompts 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print("value",round(d["value"],1),"ms/step",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"],1))
print("roofline",{k:v for k,v in d.get("roofline",{}).items() if k in ("achieved","frac","avg_launch_us","us_per_frame_step","share_of_step")})
for k,v in d.get("kernel_classes",{}).items(): print(k,v)
PY
