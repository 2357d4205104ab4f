#!/bin/bash
# Run under gpurun: compute-sanitizer over the kernels changed in session 4 (Fitch tile kernel, batched joins,
# paired DMMA groups): memcheck on the tests that drive them, racecheck on the Fitch tile kernel.
mkdir -p gpurun_out
K='fitch_tree_length or fitch_tile_program_cache or fitch_weighted or fitch_caterpillar or uppass or aa_20_state'
( timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py tests/test_gpu_uppass.py -m gpu -q -x -k "$K" ) > gpurun_out/sanitizer_memcheck_s4.log 2>&1
tail -4 gpurun_out/sanitizer_memcheck_s4.log
( timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fitch_tile_program_cache or (fitch_tree_length and tile)" ) > gpurun_out/sanitizer_racecheck_s4.log 2>&1
tail -4 gpurun_out/sanitizer_racecheck_s4.log
