#!/bin/bash
# Run under gpurun: Fitch tests (incl. the length-only mode) + both Fitch workloads with their modes.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x -k "fitch or stream" > gpurun_out/fitch_pytest.log 2>&1
tail -3 gpurun_out/fitch_pytest.log
for W in fitch fitch64; do
  timeout 200 python bench.py --workload $W --no-cpu-baseline --e2e-steps 1 --steps 20 2>>gpurun_out/fitch_lo.err | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$W', 'ms/step %.4f'%d['ms_per_step'], json.dumps(d['modes']))" | tee -a gpurun_out/fitch_lo.txt
done
PHYLO_FITCH_TIMING=1 python tools/fitch_lo_timing.py 2>&1 | grep "fitch timing" | tail -2 | tee -a gpurun_out/fitch_lo.txt
tail -c 400 gpurun_out/fitch_lo.err
