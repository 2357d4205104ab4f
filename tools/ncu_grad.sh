#!/bin/bash
# Run under gpurun: one `ncu --set full` capture of param_grad4_mma_kernel (tools/grad_prof.py, 256 taxa x 131072
# patterns), summarised on the box. $1 = tag of the output files.
TAG=${1:-v1}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:param_grad4_mma -s 1 -c 1 -f \
    -o gpurun_out/prof_grad4_$TAG python tools/grad_prof.py > gpurun_out/ncu_grad_$TAG.log 2>&1
R=gpurun_out/prof_grad4_$TAG.ncu-rep
python tools/ncu_summary.py rep $R > gpurun_out/prof_grad4_$TAG.txt 2>&1
echo "---- stall samples per source line" >> gpurun_out/prof_grad4_$TAG.txt
python tools/ncu_lines.py $R 30 >> gpurun_out/prof_grad4_$TAG.txt 2>&1
ncu -i $R --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
if len(rows)>2:
    H=rows[0]
    for want in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.sum','smsp__inst_executed.sum','sm__inst_executed_pipe_lsu.sum','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_sectors_srcunit_tex_op_read.sum','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem'):
        if want in H: print(want, rows[1][H.index(want)], rows[2][H.index(want)])
" >> gpurun_out/prof_grad4_$TAG.txt
rm -f $R
tail -3 gpurun_out/ncu_grad_$TAG.log
