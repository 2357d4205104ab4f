#!/bin/bash
# Run under gpurun: the Fitch GPU tests + the whole-tree kernels at several alignment sizes.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x -k "fitch or stream" > gpurun_out/fitch_pytest.log 2>&1
tail -3 gpurun_out/fitch_pytest.log
for P in ${SIZES:-1000000 4000000 16000000 64000000}; do
  for KRN in ${KERNELS:-tile l2}; do
    timeout 200 python bench.py --workload fitch --patterns $P --fitch-kernel $KRN --no-cpu-baseline --e2e-steps 1 --steps 20 2>>gpurun_out/fitch_ab.err | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$P $KRN', 'ms/step %.4f'%d['ms_per_step'], 'T char-ops/s %.3f'%(d['value']/1e12), 'frac', round(d.get('roofline',{}).get('frac') or 0,3), {k:round(v.get('avg_us'),1) for k,v in d.get('kernels',{}).items()}, d['check'])" | tee -a gpurun_out/fitch_ab.txt
  done
done
tail -c 600 gpurun_out/fitch_ab.err
for P in 1000000 4000000; do PHYLO_FITCH_TIMING=1 python bench.py --workload fitch --patterns $P --fitch-kernel tile --no-cpu-baseline --e2e-steps 1 --steps 20 2>&1 | grep "fitch timing" | tail -2 | tee -a gpurun_out/fitch_ab.txt; done
