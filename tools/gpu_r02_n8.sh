#!/bin/bash
# 8-GPU session: band parity (K = 5..31, real stop), strong-scaling bench of the headline frame at 8 GPUs, config 4 on 8 GPUs.
mkdir -p gpurun_out
N=${1:-8}
echo "=== band parity world $N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29511 tests/band_worker.py gpurun_out/band_report_r02_$N.json 2>&1 | grep -E "OK|FAIL|rror|mismatch" | cut -c1-150
run() { # tag workload
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --e2e-calls 1 --workload $2 2>gpurun_out/bench_$1.err | grep "^{" > gpurun_out/bench_r02_$1.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_r02_$1.json"))
print("$1", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "s/call", round(d["e2e"]["seconds_per_call"],4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
print("   ", {k:round(v,4) for k,v in d["roofline"]["family_ms_per_launch"].items()})
PY
}
run n${N}_c3 c3_blind_24mp_k15
run n${N}_c4 c4_blind_61mp_k31
