#!/bin/bash
# N-GPU check (default 2): band parity report + strong-scaling bench line of the headline frame.
mkdir -p gpurun_out
N=${1:-2}
TAG=${2:-r02b}
echo "=== band parity world $N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29511 tests/band_worker.py gpurun_out/band_report_${TAG}_$N.json 2>&1 | grep -E "OK|FAIL|rror|mismatch" | cut -c1-150
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --e2e-calls 1 2>gpurun_out/bench_${TAG}_n$N.err | grep "^{" > gpurun_out/bench_${TAG}_n$N.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_n$N.json"))
print("n$N", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
print("   ", {k:round(v,4) for k,v in d["roofline"]["family_ms_per_launch"].items()})
PY
