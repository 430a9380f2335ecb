#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}; COMM=${2:-fused}; TAG=${3:-n8_fused_b}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --comm $COMM --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_$TAG.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("$TAG", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1) if d["e2e"] else None)
print({k:round(v,4) for k,v in d["roofline"]["family_ms_per_launch"].items()})
PY
