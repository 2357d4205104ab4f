#!/bin/bash
# Run under gpurun: same-box A/B (ab_libs/libhead.so against the working tree's library) of the end-to-end legs of
# configs 4 and 5 (phylo_lk_score_alignment for 20 / 61 states) + the tests of that path.
mkdir -p gpurun_out
cp phylocaml_b200/lib/libphyloc_b200.so /tmp/libnew.so
timeout 400 python -m pytest tests -m gpu -q -x -k "alignment or aa or codon or cfg4 or cfg5 or large_alphabets or symbols or text" 2>&1 | tail -2
for V in head new head new; do
  if [ $V == head ]; then cp ab_libs/libhead.so phylocaml_b200/lib/libphyloc_b200.so; else cp /tmp/libnew.so phylocaml_b200/lib/libphyloc_b200.so; fi
  for W in aa codon; do
    timeout 300 python bench.py --workload $W --workloads none --steps 5 --warmup 3 --e2e-steps 5 --no-cpu-baseline --no-other-modes 2>>gpurun_out/treem_e2e.err | tail -1 > gpurun_out/treem_e2e_${W}_$V.json
    python -c "
import json;d=json.loads(open('gpurun_out/treem_e2e_${W}_$V.json').read());print('$V $W ms/step %.3f e2e %.3f lnl %r e2e lnl %r clocks %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['check']['result'], d['check'].get('result_e2e'), d['clocks']['sm_mhz']))"
  done
done
cp /tmp/libnew.so phylocaml_b200/lib/libphyloc_b200.so
tail -c 300 gpurun_out/treem_e2e.err
