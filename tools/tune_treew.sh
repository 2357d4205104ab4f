#!/bin/bash
# Run under gpurun: PHYLO_TREEW_TUNE="slev,R,interleave" sweep of the warp-autonomous tree kernel.
for t in "0,0,1" "2,1,1"; do
  echo "== tune $t"
  PHYLO_TREEW_TUNE=$t timeout 200 python bench.py --no-cpu-baseline --steps 10 --e2e-steps 1 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print({k:round(v['ms_per_step'],2) for k,v in d['modes'].items()}, d['check']['result'])"
done
