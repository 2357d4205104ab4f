cp phylocaml_b200/lib/libphyloc_b200.so /tmp/libnew.so
for V in head new head new; do
  if [ $V == head ]; then cp ab_libs/libhead.so phylocaml_b200/lib/libphyloc_b200.so; else cp /tmp/libnew.so phylocaml_b200/lib/libphyloc_b200.so; fi
  timeout 300 python bench.py --workload dna --workloads none --patterns 600000 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());p=d['directions'];print('$V', 'grad_ms %.2f fd_ms %.2f up %.2f joins %.2f'%(p['param_gradient']['gradient_ms'],p['param_gradient']['central_differences_ms_incl_model_setup'],p['up_pass_ms'],p['all_edge_joins_ms']), d['branch_loop']['loop_total_ms'])"
done
cp /tmp/libnew.so phylocaml_b200/lib/libphyloc_b200.so
