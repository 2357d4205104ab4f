#!/bin/bash
# Run under gpurun: launch list + one full capture per dominant kernel, into gpurun_out/.
set -x
mkdir -p gpurun_out
W=${1:-dna}
P=${2:-1000000}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$W.csv \
    python bench.py --workload $W --patterns $P --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench_$W.log 2>&1
for K in prune4_kernel.*Lb0ELb0 prune4_kernel.*Lb1ELb0 prune4_kernel.*Lb0ELb1 prune4_kernel.*Lb1ELb1 root4_kernel fitch_tree_kernel; do
  name=$(echo $K | tr -cd 'a-zA-Z0-9_')
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o gpurun_out/prof_${W}_$name \
      python bench.py --workload $W --patterns $P --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 >> gpurun_out/ncu_bench_$W.log 2>&1
done
ls -la gpurun_out
