#!/bin/bash
# Run under gpurun: ncu launch list of tools/directions_once.py (down pass, up pass, all joins, gradient).
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_directions.csv \
    python tools/directions_once.py > gpurun_out/directions_once.log 2>&1
tail -2 gpurun_out/directions_once.log
python tools/ncu_summary.py launches gpurun_out/launches_directions.csv > gpurun_out/launches_directions.txt 2>&1
cat gpurun_out/launches_directions.txt
