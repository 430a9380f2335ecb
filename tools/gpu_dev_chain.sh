#!/bin/bash
# Kernel-development session on one GPU with a K = 15 only build: chain / PSF-gradient stage parity, role timing, bench.
# usage: tools/gpu_dev_chain.sh TAG [pytest -k expression]
TAG=${1:-dev}
KEXPR=${2:-"(chain_kernel or fused_residual or row_fft_stencils) and 15"}
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "$KEXPR" 2>&1 | tail -15
echo "=== roles"; timeout 300 python tools/chain_roles.py 2>&1 | tail -16
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-calls 1 2> gpurun_out/bench_${TAG}.err | grep '^{' > gpurun_out/bench_${TAG}.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json")); r=d["roofline"]
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", d["e2e"] and round(d["e2e"]["value"],1), d["clocks"])
print("step frac", round(r["step"]["frac_of_hbm_all_gpus"],4), {k:round(v,4) for k,v in r["family_ms_per_launch"].items()})
PY
