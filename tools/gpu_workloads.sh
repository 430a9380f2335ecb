#!/bin/bash
# All BASELINE.json configs that fit one GPU, device-resident + e2e + roofline (no CPU baseline).
mkdir -p gpurun_out
for W in c1_nonblind_512_g5 c2_blind_2mp_k9 c5_nonblind_4k_kaiser7 c4_blind_61mp_k31; do
  timeout 900 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --e2e-calls 1 2>gpurun_out/wl_$W.err | grep '^{' > gpurun_out/wl_$W.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/wl_$W.json"))
    r=d["roofline"]
    print("$W", d["config"]["frame"], "K", d["config"]["psf"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"])
    print("   ", r["kernel"], "hbm frac", round(r["frac"],3), "direct-equiv/peak", round(r["fp32"]["direct_equivalent_over_peak"],3), "step hbm frac", round(r["step"]["frac_of_hbm_all_gpus"],3), {k:round(v,4) for k,v in r["family_ms_per_launch"].items()})
except Exception as e:
    print("$W failed", e); print(open("gpurun_out/wl_$W.err").read()[-800:])
PY
done
