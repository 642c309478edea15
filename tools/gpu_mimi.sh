# per-GEMM Mimi timing + the TC attention test
timeout 600 python -m pytest tests/test_gpu_flowlm.py -m gpu -x -q -p no:cacheprovider -k "tensor_core_attention" 2>&1 | tail -5
PTTS_B200_GEMM_LOG=1 timeout 600 python tools/mimi_log.py 64 100 2>&1 | tail -60
