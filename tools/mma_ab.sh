#!/bin/bash
# Run under gpurun: the DMMA pruning kernel on configs 4 and 5 (bench) + the tests that exercise it.
# (History: this script A/B-ed PHYLO_MMA_R / PHYLO_MMA_TUNE variants that were measured and removed.)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "aa or codon or cfg4 or cfg5 or edge or group or large_alphabets" > gpurun_out/mma_pytest.log 2>&1
tail -3 gpurun_out/mma_pytest.log
timeout 200 python bench.py --workload aa --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-other-modes > gpurun_out/mma_aa.json 2>> gpurun_out/mma.err
timeout 200 python bench.py --workload codon --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-other-modes > gpurun_out/mma_codon.json 2>> gpurun_out/mma.err
tail -c 400 gpurun_out/mma.err
