#!/bin/bash
# A/B of two builds of the library on the headline bench (device-resident only): tools/gpu_r02_ab.sh TAG LIB_B
mkdir -p gpurun_out
TAG=${1:-ab}
echo "=== chain stage + golden (default lib)"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "chain_kernel or golden or workloads" 2>&1 | tail -4
for L in default "$2"; do
  if [ "$L" = default ]; then unset RLTV_LIB; else export RLTV_LIB=$PWD/$L; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/bench_${TAG}.err | grep '^{' > gpurun_out/bench_${TAG}_$(basename ${L%.so}).json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_$(basename ${L%.so}).json")); r=d["roofline"]
print("$L", "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "step frac", round(r["step"]["frac_of_hbm_all_gpus"],4), {k:round(v,4) for k,v in r["family_ms_per_launch"].items()})
PY
done
