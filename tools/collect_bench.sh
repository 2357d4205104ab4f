#!/bin/bash
# Run under gpurun: the bench lines that profiles/README.md quotes, into gpurun_out/.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_dna.json 2> gpurun_out/bench_dna.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_dna_reference.json 2>> gpurun_out/bench_dna.err
python bench.py --workload fitch > gpurun_out/bench_fitch_1M.json 2>> gpurun_out/bench_dna.err
python bench.py --workload fitch --patterns 64000000 --steps 10 --e2e-steps 2 --no-cpu-baseline > gpurun_out/bench_fitch_64M.json 2>> gpurun_out/bench_dna.err
python bench.py --workload aa --steps 5 --warmup 2 --e2e-steps 2 > gpurun_out/bench_aa.json 2>> gpurun_out/bench_dna.err
python bench.py --workload codon --steps 5 --warmup 2 --e2e-steps 2 > gpurun_out/bench_codon.json 2>> gpurun_out/bench_dna.err
python bench.py --patterns 10000 --taxa 16 --steps 50 --warmup 5 > gpurun_out/bench_cfg1.json 2>> gpurun_out/bench_dna.err
tail -c 600 gpurun_out/bench_dna.err
for f in gpurun_out/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" ", d.get("metric"), "%.4g"%d["value"], d["unit"], "ms/step %.4g"%d["ms_per_step"], "roofline", (d.get("roofline") or {}).get("frac"), "modes", {k:round(v["ms_per_step"],2) for k,v in (d.get("modes") or {}).items()}, "e2e ms", (d.get("e2e") or {}).get("ms_per_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "check", d.get("check"))
except Exception as ex: print("  parse failed", ex)
PY
done
python bench.py --workload compress --steps 8 --warmup 1 > gpurun_out/bench_compress.json 2>> gpurun_out/bench_dna.err
