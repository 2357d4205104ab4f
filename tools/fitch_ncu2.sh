#!/bin/bash
# Run under gpurun: ncu --set full capture of one whole-tree Fitch kernel. usage: fitch_ncu2.sh <kernel> <patterns> <regex>
mkdir -p gpurun_out
KRN=${1:-warptile}; P=${2:-1000000}; RX=${3:-fitch_warp}
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$RX -s 3 -c 1 -f \
    -o gpurun_out/prof_fitch_${KRN}_$P python bench.py --workload fitch --patterns $P --fitch-kernel $KRN --no-cpu-baseline --e2e-steps 1 --steps 3 --warmup 3 > gpurun_out/fitch_ncu2.log 2>&1
grep -E "==PROF== Profiling|No kernels" gpurun_out/fitch_ncu2.log | head -3
