#!/bin/bash
# chain-kernel bring-up: stage test under a short timeout first, then solver parity, then bench
mkdir -p gpurun_out
TAG=${1:-chain}
echo "=== chain stage"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "chain_kernel" 2>&1 | tail -15
echo "=== golden + workloads"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "golden or workloads or row_fft_solver or live_reference" 2>&1 | tail -15
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-calls 1 2> gpurun_out/bench_${TAG}.err | grep '^{' > gpurun_out/bench_${TAG}.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json")); r=d["roofline"]
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", d["e2e"] and round(d["e2e"]["value"],1), d["clocks"])
print("step frac", round(r["step"]["frac_of_hbm_all_gpus"],4), {k:round(v,4) for k,v in r["family_ms_per_launch"].items()})
PY
tail -3 gpurun_out/bench_${TAG}.err
