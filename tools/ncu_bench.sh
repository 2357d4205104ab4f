#!/bin/bash
# Run under gpurun: launch list + one full capture per dominant kernel, into gpurun_out/.
# usage: tools/ncu_bench.sh <workload> <patterns> [kernel-regex-on-mangled-name ...]
set -x
mkdir -p gpurun_out
W=${1:-dna}
P=${2:-1000000}
shift 2
CMD="python bench.py --workload $W --patterns $P --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-other-modes"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$W.csv \
    $CMD > gpurun_out/ncu_bench_$W.log 2>&1
for K in "$@"; do
  name=$(echo $K | tr -cd 'a-zA-Z0-9_')
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$K -s 1 -c 1 -f \
      -o gpurun_out/prof_${W}_$name $CMD >> gpurun_out/ncu_bench_$W.log 2>&1
done
grep -E "==PROF== Profiling|No kernels" gpurun_out/ncu_bench_$W.log
ls -la gpurun_out
