set -x
python -m pytest tests/test_reference_golden.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/ref_golden_gpu.log
cat gpurun_out/ref_golden_gpu.log
for rep in 1 2; do
for cfg in "65536 32" "65536 16" "262144 32" "1048576 32" "262144 100" "1048576 100" "4194304 100" "16777216 223"; do
  for v in queue sync; do printf "%-6s " $v; ./build/k1/$v $cfg; done
done; done 2>&1 | tee gpurun_out/k1_queue_vs_sync.log
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_queue.log
