#!/bin/bash
# Run under gpurun: same-box A/B of two builds of the library on config 3: ab_libs/libhead.so (the committed kernel)
# against the working tree's. Per build: one ncu capture at 600 k patterns (cycles, instructions) and two bench lines.
mkdir -p gpurun_out
cp phylocaml_b200/lib/libphyloc_b200.so /tmp/libnew.so
timeout 300 python -m pytest tests -m gpu -q -x -k "${TESTS:-tip_tables or fused_tree_equals or kernel_variants or other_rate_counts or cfg3}" 2>&1 | tail -2
for V in head new head new; do
  if [ $V == head ]; then cp ab_libs/libhead.so phylocaml_b200/lib/libphyloc_b200.so; else cp /tmp/libnew.so phylocaml_b200/lib/libphyloc_b200.so; fi
  timeout 300 python bench.py --workload dna --workloads none --steps 10 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-other-modes 2>>gpurun_out/treew_ab.err | tail -1 > gpurun_out/treew_ab_$V.json
  python -c "
import json;d=json.loads(open('gpurun_out/treew_ab_$V.json').read());print('$V dna ms/step %.3f e2e %.3f frac %.3f lnl %r clocks %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['check']['result'], d['clocks']['sm_mhz']))"
done
if [ -z "$SKIP_NCU" ]; then
for V in head new; do
  if [ $V == head ]; then cp ab_libs/libhead.so phylocaml_b200/lib/libphyloc_b200.so; else cp /tmp/libnew.so phylocaml_b200/lib/libphyloc_b200.so; fi
  bash tools/ncu_treew.sh ${TAG:-ab}_$V > /dev/null 2>&1
  echo "== $V"; grep -E "gpu__time_duration|smsp__inst_executed|smsp__cycles_active|registers_per_thread|stall:(long|wait|short|mio)" gpurun_out/prof_treew_${TAG:-ab}_$V.txt
done
fi
cp /tmp/libnew.so phylocaml_b200/lib/libphyloc_b200.so
