#!/bin/bash
# Run under `gpurun --gpus N`: the driver's launch line for N ranks, output to gpurun_out/.
N=${1:-2}
shift
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 3 "$@" 2> gpurun_out/scale_$N.err | tail -1 > gpurun_out/scale_$N.json
python - <<PY
import json
d=json.load(open("gpurun_out/scale_$N.json"))
print("N=$N", d["value"], d["ms_per_step"], d["modes"], d["e2e"]["ms_per_step"], d["check"])
PY
