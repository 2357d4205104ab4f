#!/bin/bash
# One gpurun call (round 2): GPU test suite, smoke, the driver's default bench line (every BASELINE
# config inside it) and the CPU arm. Everything lands in gpurun_out/; each leg has its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
if [ "$1" != "nobench" ]; then
( time timeout 400 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_default.json 2> gpurun_out/bench.err
tail -c 600 gpurun_out/bench.err
python tools/bench_summary.py gpurun_out/bench_default.json
fi
if [ "$1" != "notest" ]; then
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --durations=10 ${PYTEST_ARGS} ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
fi
if [ "$1" == "ref" ]; then
( time timeout 400 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
python tools/bench_summary.py gpurun_out/bench_reference.json
fi
