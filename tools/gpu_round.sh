# usage: bash tools/gpu_round.sh TAG [tests|notests] [ncu|noncu]
#   GPU parity tests, bench (own arm), phase trace, ncu launch list + one full capture of k_flow_mega. Everything lands in gpurun_out/.
TAG=${1:-r02}; TESTS=${2:-tests}; NCU=${3:-ncu}
mkdir -p gpurun_out
python -c "import bench; print(bench.src_sha16())" > gpurun_out/${TAG}_src_sha16.txt 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" > gpurun_out/${TAG}_summary0.txt; tail -1 gpurun_out/${TAG}_smoke.log
if [ "$TESTS" = tests ]; then
  timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/${TAG}_summary.txt
  grep -E "max-rel|dB|passed|failed|error|Error" gpurun_out/${TAG}_pytest_gpu.log | tail -40
fi
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?" >> gpurun_out/${TAG}_summary.txt
tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python tools/mega_trace.py 256 100 2> gpurun_out/${TAG}_mega_trace.txt | tail -1 | cut -c1-300
if [ "$NCU" = ncu ]; then
  timeout 300 python tools/ncu_step.py 256 100 1 > gpurun_out/ncu_plain.log 2>&1 && \
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/${TAG}_launches_b256_f100_ncu.csv python tools/ncu_step.py 256 100 1 > gpurun_out/ncu.log 2>&1; echo "ncu list exit $?" >> gpurun_out/${TAG}_summary.txt
  timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:k_flow_mega -c 1 -o gpurun_out/${TAG}_flow_mega -f python tools/ncu_step.py 256 100 0 > gpurun_out/ncu_mega.log 2>&1; echo "ncu mega exit $?" >> gpurun_out/${TAG}_summary.txt
fi
cat gpurun_out/${TAG}_summary.txt
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print("value",round(d["value"],1),"ms/step",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"],1),"clocks",d["clocks"])
print("roofline",{k:v for k,v in d.get("roofline",{}).items() if k in ("achieved","frac","avg_launch_us","us_per_frame_step","traffic","share_of_step","tensor_frac")})
print("mimi",d.get("roofline_mimi_gemm"))
for k,v in d.get("kernel_classes",{}).items(): print(k,v)
print("batch1",d.get("batch1")); print("mimi_only",d.get("mimi_decode_only")); print("sweep",d.get("sweep")); print("cpu",d.get("cpu_baseline"))
PY
