#!/bin/bash
# Build tuning variants of the search kernel into build/k1/ (run here, on the CPU box), then
# `gpurun -- tools/k1_variants.sh run` times them on the B200.  Run-time switches of every binary:
# TFRT_K1_QUEUE=0|1|2 (item queue never / auto / always), TFRT_K1_ITEMS_PER_CTA=n.
set -e
cd "$(dirname "$0")/.."
SRC="tools/k1_bench.cu tensorflowraytrace_b200/csrc/abi.cu tensorflowraytrace_b200/csrc/pack.cu tensorflowraytrace_b200/csrc/intersect3d.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17"
declare -A V
# round-2 study of the triangle-pair layout (one ray x two triangles per packed instruction):
# source order of the test, occupancy, ring depth, threads per CTA
V[pair]=""
V[pair_nr6]="-DTFRT_K1_FORCE_NR=6 -DTFRT_K1_MINBLOCKS=3"
V[pair_nr10]="-DTFRT_K1_FORCE_NR=10 -DTFRT_K1_MINBLOCKS=2"
V[pair_mb2]="-DTFRT_K1_MINBLOCKS=2"
V[pair_mb4]="-DTFRT_K1_MINBLOCKS=4"
V[pair_s4]="-DTFRT_K1_STAGES=4"
V[pair_t256]="-DTFRT_K1_THREADS=256 -DTFRT_K1_MINBLOCKS=1"
if [ "$1" = "run" ]; then
  for v in "${!V[@]}"; do printf "%-12s " $v; ./build/k1/$v ${2:-4194304} ${3:-100}; done | sort
  exit 0
fi
mkdir -p build/k1
for v in "${!V[@]}"; do
  nvcc $FLAGS ${V[$v]} -Xptxas -v $SRC -o build/k1/$v 2> build/k1/$v.ptxas &
done
wait
for v in "${!V[@]}"; do printf "%-12s " $v; grep -A1 "k_intersect3d" build/k1/$v.ptxas | grep -o "Used [0-9]* registers" | tr '\n' ' '; grep -c "spill" build/k1/$v.ptxas | tr '\n' ' '; grep "bytes spill" build/k1/$v.ptxas | grep -v " 0 bytes spill stores, 0 bytes spill loads" | head -1; echo; done
