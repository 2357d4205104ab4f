#!/bin/bash
# Run under gpurun: launch lists + ncu --set full captures of the Fitch kernels (both regimes)
# and a fresh launch list of the default DNA bench command.
mkdir -p gpurun_out
bash tools/ncu_bench.sh fitch 1000000 fitch_tile_kernel > /dev/null 2>&1
mv gpurun_out/launches_fitch.csv gpurun_out/launches_fitch_1M.csv
bash tools/ncu_bench.sh fitch 64000000 fitch_tree_kernel > /dev/null 2>&1
mv gpurun_out/launches_fitch.csv gpurun_out/launches_fitch_64M.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_dna_default.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-other-modes --no-cpu-baseline > gpurun_out/ncu_dna_default.log 2>&1
for f in gpurun_out/prof_fitch_*.ncu-rep; do
  python tools/ncu_summary.py rep $f > ${f%.ncu-rep}.txt 2>&1
done
ls -la gpurun_out | head -30
head -12 gpurun_out/prof_fitch_fitch_tile_kernel.txt gpurun_out/prof_fitch_fitch_tree_kernel.txt
