#!/bin/bash
# Run under gpurun: quick A/B of a lk_treew_kernel change: the bit-identity tests, one ncu capture (cycles and
# instruction counts at 600 k patterns are box-independent), one config-3 bench line. usage: tools/treew_quick.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "tip_tables or fused_tree_equals or kernel_variants or other_rate_counts or cfg3" 2>&1 | tail -2
bash tools/ncu_treew.sh $TAG > /dev/null 2>&1
grep -E "gpu__time_duration|smsp__inst_executed|smsp__cycles_active|registers_per_thread|pipe_fp64|bank_conflicts|stall" gpurun_out/prof_treew_$TAG.txt
timeout 300 python bench.py --workload dna --workloads none --steps 10 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-other-modes 2>>gpurun_out/treew_quick.err | tail -1 > gpurun_out/treew_quick_dna_$TAG.json
python -c "
import json;d=json.loads(open('gpurun_out/treew_quick_dna_$TAG.json').read());print('dna ms/step %.3f e2e %.3f frac %.3f lnl %r clocks %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['check']['result'], d['clocks']['sm_mhz']))"
