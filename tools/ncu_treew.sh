#!/bin/bash
# Run under gpurun: one ncu --set full capture of the 4-state tree-fused kernel (config 3 shape, 600 k patterns),
# summarised on the box (summary, stall samples per source line, opcode mix). usage: tools/ncu_treew.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
R=gpurun_out/prof_treew_$TAG.ncu-rep
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:lk_treew -s 2 -c 1 -f -o ${R%.ncu-rep} \
  python bench.py --workload dna --workloads none --patterns 600000 --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 1 --no-other-modes > gpurun_out/ncu_treew_$TAG.log 2>&1
grep -h "==PROF== Profiling" gpurun_out/ncu_treew_$TAG.log | head -3
B=gpurun_out/prof_treew_$TAG
python tools/ncu_summary.py rep $R > $B.txt 2>&1
echo "---- stall samples per source line" >> $B.txt
python tools/ncu_lines.py $R 45 >> $B.txt 2>&1
echo "---- opcode mix" >> $B.txt
python tools/ncu_opmix.py $R >> $B.txt 2>&1
ls -la $R; rm -f $R
tail -5 gpurun_out/ncu_treew_$TAG.log
