#!/bin/bash
# Run under gpurun: true GPU durations of the Fitch tree kernels at 1 M characters.
mkdir -p gpurun_out
for KRN in tile regwalk l2; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fitch_t -c 12 --csv --log-file gpurun_out/fitch_launches_$KRN.csv \
    python bench.py --workload fitch --fitch-kernel $KRN --no-cpu-baseline --e2e-steps 1 --steps 5 --warmup 3 > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/fitch_launches_$KRN.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
H=rows[hi]; vi=H.index("Metric Value"); ui=H.index("Metric Unit")
vals=[float(r[vi].replace(",","")) for r in rows[hi+1:] if len(r)>vi]
print("$KRN", rows[hi+1][ui], sorted(vals)[:12])
PY
done
