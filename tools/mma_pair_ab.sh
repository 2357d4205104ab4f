#!/bin/bash
# Run under gpurun: same-box A/B of the paired CLV+CLV path of the tree-fused DMMA kernel (config 4), three rounds.
for i in 1 2 3; do
  for P in 0 1; do
    PHYLO_TREEM_PAIRED=$P timeout 200 python bench.py --workload aa --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-other-modes 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('paired=$P', 'ms/step %.3f'%d['ms_per_step'], d['check']['result'], d['clocks']['sm_mhz'])"
  done
done
