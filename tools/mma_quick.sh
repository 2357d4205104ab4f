#!/bin/bash
# Run under gpurun: the small 20-state test first (under compute-sanitizer if it fails), then tools/mma_ab.sh
mkdir -p gpurun_out
if ! timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "aa_20_state or large_alphabets" > gpurun_out/mma_quick.log 2>&1; then
  tail -5 gpurun_out/mma_quick.log
  timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "aa_20_state" > gpurun_out/mma_sanitizer.log 2>&1
  grep -A14 "Invalid\|ERROR SUMMARY" gpurun_out/mma_sanitizer.log | head -60
  exit 1
fi
tail -2 gpurun_out/mma_quick.log
bash tools/mma_ab.sh
python -c "
import json
for w in ('aa','codon'):
    d=json.loads(open('gpurun_out/mma_%s.json'%w).read().strip().splitlines()[-1]); print(w, d['ms_per_step'], d['check'])
"
