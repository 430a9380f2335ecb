#!/bin/bash
# Round-2 development session on one GPU: parity tests, headline bench, launch list, full ncu of the hot kernels.
# usage: tools/gpu_r02_dev.sh TAG [pytest -k expression]
TAG=${1:-dev}
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 ${2:+-k "$2"} 2>&1 | tail -15
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-calls 1 2> gpurun_out/bench_${TAG}.err | grep '^{' > gpurun_out/bench_${TAG}.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json")); r=d["roofline"]
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", d["e2e"] and round(d["e2e"]["value"],1), d["clocks"])
print("step frac", round(r["step"]["frac_of_hbm_all_gpus"],4), {k:round(v,4) for k,v in r["family_ms_per_launch"].items()})
PY
echo "=== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_conv_fft|k_gradk_fft|k_update' -s 6 -c 7 -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_${TAG}.log 2>&1; tail -2 gpurun_out/ncu_${TAG}.log
