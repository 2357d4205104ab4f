#!/bin/bash
# Run under gpurun: tests of the directional CLVs / edge joins + the bench record that times them.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x -k "uppass or edge or direction or branch" > gpurun_out/edge_pytest.log 2>&1
tail -3 gpurun_out/edge_pytest.log
timeout 300 python bench.py --patterns 1000000 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline 2>gpurun_out/edge.err | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(json.dumps(d.get('directions'),indent=1)[:1500])" | tee gpurun_out/edge_ab.txt
tail -c 300 gpurun_out/edge.err
