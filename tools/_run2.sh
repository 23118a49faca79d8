set -x
python -m pytest tests/test_reference_golden.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/ref_golden_gpu.log
cat gpurun_out/ref_golden_gpu.log
for cfg in "65536 32" "65536 16" "262144 32" "1048576 32" "262144 100" "1048576 100" "4194304 100" "16777216 223"; do
  printf "hybrid "; ./build/k1/hybrid $cfg
  printf "always "; TFRT_K1_QUEUE=2 ./build/k1/hybrid $cfg
  printf "never  "; TFRT_K1_QUEUE=0 ./build/k1/hybrid $cfg
done 2>&1 | grep -v "^+" | tee gpurun_out/k1_hybrid.log
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
for c in cfg2 cfg3 cfg5; do python bench.py --config $c > gpurun_out/bench_$c.log 2> gpurun_out/bench_$c.err; tail -c 600 gpurun_out/bench_$c.log; done
