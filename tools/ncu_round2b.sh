#!/bin/bash
# Run under gpurun: ncu --set full captures of the kernels changed late in round 2 (Fitch tile kernel at 1 M and 64 Mi
# characters, tree-fused DMMA kernel on config 4) + launch lists of the Fitch workloads.
mkdir -p gpurun_out
for P in 1000000 67108864; do
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:fitch_tile -s 3 -c 1 -f \
      -o gpurun_out/prof_fitch_tile_$P python bench.py --workload fitch --patterns $P --no-cpu-baseline --e2e-steps 1 --steps 3 --warmup 3 > gpurun_out/ncu2b_$P.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_fitch_$P.csv \
      python bench.py --workload fitch --patterns $P --no-cpu-baseline --e2e-steps 1 --steps 3 --warmup 3 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:lk_treem -s 2 -c 1 -f \
    -o gpurun_out/prof_treem_aa_v9 python bench.py --workload aa --no-cpu-baseline --e2e-steps 1 --steps 2 --warmup 2 --no-other-modes > gpurun_out/ncu2b_aa.log 2>&1
grep -h "==PROF== Profiling" gpurun_out/ncu2b_*.log | head
ls -la gpurun_out/*.ncu-rep
# summaries on the box: the reports themselves exceed what gpurun copies back
for R in gpurun_out/prof_*.ncu-rep; do
  B=$(basename $R .ncu-rep)
  python tools/ncu_summary.py rep $R > gpurun_out/$B.txt 2>&1
  echo "---- stall samples per source line" >> gpurun_out/$B.txt
  python tools/ncu_lines.py $R 30 >> gpurun_out/$B.txt 2>&1
  ncu -i $R --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
if len(rows)>2:
    H=rows[0]
    for want in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.sum','smsp__inst_executed.sum','sm__inst_executed_pipe_lsu.sum'):
        if want in H: print(want, rows[1][H.index(want)], rows[2][H.index(want)])
" >> gpurun_out/$B.txt
  rm -f $R
done
for P in 1000000 67108864; do python tools/ncu_summary.py launches gpurun_out/launches_fitch_$P.csv > gpurun_out/launches_fitch_$P.txt 2>&1; done
ls -la gpurun_out
