#!/bin/bash
# Run under gpurun: the 20-state tests, then a same-box A/B of the conflict-free result-store order of the tree-fused
# DMMA kernel (config 4), three rounds. PHYLO_TREEM_STSWAP=0 restores the row-order stores.
timeout 300 python -m pytest tests -m gpu -q -x -k "aa or cfg4 or large_alphabets or treem" 2>&1 | tail -2
for i in 1 2 3; do
  for P in 0 1; do
    PHYLO_TREEM_STSWAP=$P timeout 200 python bench.py --workload aa --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-other-modes 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('stswap=$P', 'ms/step %.3f'%d['ms_per_step'], d['check']['result'], d['clocks']['sm_mhz'])"
  done
done
