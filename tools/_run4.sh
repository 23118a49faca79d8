for cfg in "4194304 100" "1048576 100" "16777216 223"; do
  for v in hybrid scalar scalar_u1 scalar_u4 scalar_t256; do printf "%-12s " $v; ./build/k1/$v $cfg; done
done 2>&1 | tee gpurun_out/k1_scalar.log
