#!/bin/bash
# Run under gpurun: the three whole-tree Fitch kernels at several alignment sizes.
for P in 1000000 4000000 16000000 64000000; do
  for KRN in tile regwalk l2; do
    python bench.py --workload fitch --patterns $P --fitch-kernel $KRN --no-cpu-baseline --e2e-steps 1 --steps 20 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$P $KRN', 'ms/step %.4f'%d['ms_per_step'], 'T char-ops/s %.3f'%(d['value']/1e12), 'frac', round(d.get('roofline',{}).get('frac') or 0,3), {k:round(v.get('avg_us'),1) for k,v in d.get('kernels',{}).items()}, d['check'])"
  done
done
