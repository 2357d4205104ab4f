#!/bin/bash
# Run under gpurun: up-pass tests, then the `directions` record with one launch per update (PHYLO_UPPASS_BATCH=0)
# and one launch per tree level, alternating.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_uppass.py tests/test_gpu_gradient.py -x -q 2>&1 | tail -8 | tee gpurun_out/uppass_tests.log
rm -f gpurun_out/uppass_ab.txt
for V in 0 1 0 1; do
  PHYLO_UPPASS_BATCH=$V timeout 300 python bench.py --workload dna --workloads none --patterns 600000 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());p=d['directions'];g=p['param_gradient'];print('PHYLO_UPPASS_BATCH=$V', 'down %.2f up %.2f joins %.2f grad %.2f ms'%(p['down_pass_ms'],p['up_pass_ms'],p['all_edge_joins_ms'],g['gradient_ms']), 'max rel diff of edge lnL', p['max_rel_diff_of_edge_lnl_vs_root_edge'], g['gradient'][:2])" | tee -a gpurun_out/uppass_ab.txt
done
