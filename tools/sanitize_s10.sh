#!/bin/bash
# Run under gpurun: compute-sanitizer over the kernels added in the last session of round 2 (param_grad4_mma_kernel,
# grad_sum_branches_kernel, prune4_level_kernel): memcheck on the tests that drive them, racecheck on the gradient
# kernel (shared-memory A fragments + parity-double-buffered warp sums).
mkdir -p gpurun_out
K='tensor_core_gradient or level_batched'
( timeout 140 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_gradient.py tests/test_gpu_uppass.py -m gpu -q -x -k "$K" ) > gpurun_out/sanitizer_memcheck_s10.log 2>&1
tail -4 gpurun_out/sanitizer_memcheck_s10.log
( timeout 100 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_gradient.py -m gpu -q -x -k "tensor_core_gradient and 1025" ) > gpurun_out/sanitizer_racecheck_s10.log 2>&1
tail -4 gpurun_out/sanitizer_racecheck_s10.log
