# usage: bash tools/gpu_quick.sh TAG [FLAGS_B]  - step-kernel time (default / with PTTS_B200_MEGA_FLAGS=FLAGS_B), phase trace, a short parity subset
TAG=${1:-q}; FB=${2:-32}; mkdir -p gpurun_out
timeout 200 python tools/step_time.py 256 100 3 2>&1 | tail -1
PTTS_B200_MEGA_FLAGS=$FB timeout 200 python tools/step_time.py 256 100 3 2>&1 | tail -1
timeout 200 python tools/mega_trace.py 256 100 2> gpurun_out/${TAG}_mega_trace.txt | tail -1 | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_flowlm.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
