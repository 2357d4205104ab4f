#!/bin/bash
# Run under gpurun: gradient tests, then the `directions.param_gradient` record with the scalar per-branch
# kernel (PHYLO_GRAD_MMA=0) and the batched DMMA kernel, alternating.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gradient.py -x -q 2>&1 | tail -15 | tee gpurun_out/grad_tests.log
for V in 0 1 0 1; do
  PHYLO_GRAD_MMA=$V timeout 300 python bench.py --workload dna --workloads none --patterns 600000 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());p=d['directions'];g=p['param_gradient'];print('PHYLO_GRAD_MMA=$V', 'grad_ms %.2f fd_ms %.2f up %.2f joins %.2f'%(g['gradient_ms'],g['central_differences_ms_incl_model_setup'],p['up_pass_ms'],p['all_edge_joins_ms']), 'rel diff vs central differences %.2e'%g['max_rel_diff_vs_central_differences'], g['gradient'])" | tee -a gpurun_out/grad_mma_ab.txt
done
