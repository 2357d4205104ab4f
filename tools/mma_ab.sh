#!/bin/bash
# Run under gpurun: A/B of prune_mma_kernel's tuning bits (PHYLO_MMA_TUNE: 1 = L2 prefetch of the
# next group, 2 = rate classes pipelined in registers) on configs 4 and 5, then the tests.
mkdir -p gpurun_out
for TUNE in ${TUNES:-0 1 2 3}; do
  PHYLO_MMA_TUNE=$TUNE timeout 200 python bench.py --workload aa --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-other-modes > gpurun_out/mma_aa_T$TUNE.json 2>> gpurun_out/mma.err
done
for TUNE in 0 1; do
  PHYLO_MMA_TUNE=$TUNE timeout 200 python bench.py --workload codon --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-other-modes > gpurun_out/mma_codon_T$TUNE.json 2>> gpurun_out/mma.err
done
for TUNE in ${TEST_TUNES:-0 3}; do
  PHYLO_MMA_TUNE=$TUNE timeout 300 python -m pytest tests -m gpu -q -x -k "aa or codon or cfg4 or cfg5 or edge or group or large_alphabets" > gpurun_out/mma_pytest_T$TUNE.log 2>&1
  tail -2 gpurun_out/mma_pytest_T$TUNE.log
done
tail -c 400 gpurun_out/mma.err
for f in gpurun_out/mma_*_T*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.3f"%d["ms_per_step"], {k:(round(v["avg_us"],1), round(v.get("frac",0),3), round(v.get("achieved_tflops",0),1)) for k,v in d["kernels"].items() if k.startswith("prune") or k.startswith("root")}, "lnl", d["check"]["result"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as ex: print(sys.argv[1], "parse failed", ex)
PY
done
