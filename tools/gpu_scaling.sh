#!/bin/bash
# Strong-scaling runs of the headline workload on one box (fused peer-memory comm; NCCL baseline at N=8).
mkdir -p gpurun_out
run() { # N comm tag
  local gpus=$(seq -s, 0 $(($1-1)))
  CUDA_VISIBLE_DEVICES=$gpus timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port 2952$1 bench.py --gpus $1 --steps 10 --warmup 3 --comm $2 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_$3.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$3.json"))
print("$3", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1) if d["e2e"] else None, d["clocks"])
PY
}
run 8 fused n8_fused
run 8 nccl n8_nccl
run 4 fused n4_fused
run 2 fused n2_fused
timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('n1', round(d['value'],1), d['ms_per_step'], d['e2e']['value'], d['clocks'], d['cpu_baseline'])"
