#!/bin/bash
# A/B of the fused residual + PSF-gradient kernel (RLTV_FUSE=1, default) against the two-kernel sequence (RLTV_FUSE=0).
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for m in 1 0; do
  RLTV_FUSE=$m python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_fuse_$m.json
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_fuse_$m.json"))
print("fuse=$m", round(d["value"], 1), round(d["ms_per_step"], 4), d["gpu_launches"], d["roofline"]["kernel"], round(d["roofline"]["frac"], 3),
      {k: round(v, 4) for k, v in d["roofline"]["family_ms_per_launch"].items()})
PY
done
