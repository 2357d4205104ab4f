#!/bin/bash
# Run under gpurun: config 3 at the per-rank sizes of 8, 4 and 2 GPUs on one GPU -- resident step against the
# end-to-end call (slab plan of phylo_lk_score_alignment), lnL of both.
for P in 500000 1000000 2000000; do
  timeout 300 python bench.py --workload dna --workloads none --patterns $P --steps 10 --warmup 3 --e2e-steps 5 --no-cpu-baseline --no-other-modes 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('patterns $P', 'ms/step %.3f e2e %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']), d['check'], d['clocks']['sm_mhz'])"
done
