#!/bin/bash
# Run under gpurun: ncu --set full capture of the tree-fused kernel in both modes.
# usage: tools/ncu_tree.sh <tag> [patterns]
TAG=${1:-v8}
P=${2:-600000}
mkdir -p gpurun_out
for MODE in fused fused-lnl; do
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:lk_tree -s 1 -c 1 -f \
      -o gpurun_out/prof_tree_${TAG}_$MODE \
      python bench.py --workload dna --patterns $P --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --mode $MODE --no-other-modes \
      > gpurun_out/ncu_tree_${TAG}_$MODE.log 2>&1
  grep -E "==PROF== Profiling|No kernels" gpurun_out/ncu_tree_${TAG}_$MODE.log
done
