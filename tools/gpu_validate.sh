#!/bin/bash
# One gpurun call: GPU test suite, smoke, then the bench lines profiles/README.md quotes.
# Everything lands in gpurun_out/; each leg has its own timeout so one hang cannot eat the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -q --maxfail=12 --durations=15 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_dna.json 2> gpurun_out/bench.err
timeout 120 python bench.py --workload fitch > gpurun_out/bench_fitch_1M.json 2>> gpurun_out/bench.err
timeout 120 python bench.py --patterns 10000 --taxa 16 --steps 50 --warmup 5 > gpurun_out/bench_cfg1.json 2>> gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_dna_reference.json 2>> gpurun_out/bench.err
timeout 100 python bench.py --impl reference --workload fitch --steps 5 --warmup 1 > gpurun_out/bench_fitch_reference.json 2>> gpurun_out/bench.err
timeout 200 python bench.py --workload aa --steps 5 --warmup 3 --e2e-steps 2 > gpurun_out/bench_aa.json 2>> gpurun_out/bench.err
timeout 200 python bench.py --workload codon --steps 5 --warmup 3 --e2e-steps 2 > gpurun_out/bench_codon.json 2>> gpurun_out/bench.err
timeout 150 python bench.py --workload fitch --patterns 64000000 --steps 10 --e2e-steps 2 --no-cpu-baseline > gpurun_out/bench_fitch_64M.json 2>> gpurun_out/bench.err
tail -c 800 gpurun_out/bench.err
for f in gpurun_out/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" ", d.get("metric"), "%.4g"%d["value"], "ms/step %.4g"%d["ms_per_step"], "roof", (d.get("roofline") or {}).get("frac"), "modes", {k:round(v["ms_per_step"],2) for k,v in (d.get("modes") or {}).items()}, "e2e ms", (d.get("e2e") or {}).get("ms_per_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("kind"), "l2:", (d.get("config") or {}).get("l2","")[:40], "check", d.get("check"))
except Exception as ex: print("  parse failed", ex)
PY
done
