nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,driver_version --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import torch;print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))" >> gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
for f in gemm flowlm mimi; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/t_$f.log
  echo "== $f exit ${PIPESTATUS[0]}" >> gpurun_out/summary.txt
  grep -E "passed|failed|error" gpurun_out/t_$f.log | tail -3 >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
