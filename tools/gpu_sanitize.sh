# usage: bash tools/gpu_sanitize.sh TAG -- compute-sanitizer memcheck / racecheck / synccheck over tools/sanitize_job.py (reference: `make cuda-validate`
# and the sanitizer targets of the reference Makefile:155-160). Logs land in gpurun_out/; the summary lines are what profiles/ keeps.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 300 python tools/sanitize_job.py > gpurun_out/${TAG}_sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_sanitize_plain.log; exit 1; }
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_job.py > gpurun_out/${TAG}_sanitize_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize job ok|Error|error" gpurun_out/${TAG}_sanitize_$tool.log | head -12
done
