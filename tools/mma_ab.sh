#!/bin/bash
# Run under gpurun: the DMMA pruning kernel on configs 4 and 5 (bench) + the tests that exercise it.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "aa or codon or cfg4 or cfg5 or edge or group or large_alphabets" > gpurun_out/mma_pytest.log 2>&1
tail -5 gpurun_out/mma_pytest.log
timeout 200 python bench.py --workload aa --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-other-modes > gpurun_out/mma_aa.json 2>> gpurun_out/mma.err
timeout 200 python bench.py --workload codon --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-other-modes > gpurun_out/mma_codon.json 2>> gpurun_out/mma.err
tail -c 400 gpurun_out/mma.err
for f in gpurun_out/mma_aa.json gpurun_out/mma_codon.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.3f"%d["ms_per_step"], {k:(round(v["avg_us"],1), round(v.get("frac",0),3), round(v.get("achieved_tflops",0),1)) for k,v in d["kernels"].items() if k.startswith("prune") or k.startswith("root")}, "lnl", d["check"]["result"], "clocks", d["clocks"])
except Exception as ex: print(sys.argv[1], "parse failed", ex)
PY
done
if [ -n "$NCU_MMA" ]; then
  bash tools/ncu_bench.sh aa 500000 'prune_mma_kernelILi20EjLb0ELb0' 'prune_mma_kernelILi20EjLb1ELb0' > /dev/null 2>&1
  for f in gpurun_out/prof_aa_*.ncu-rep; do python tools/ncu_summary.py rep $f > ${f%.ncu-rep}.txt 2>&1; done
  head -30 gpurun_out/prof_aa_*Lb0ELb0*.txt
fi
