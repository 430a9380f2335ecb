#!/bin/bash
# Strong-scaling runs of the headline workload on one box (fused peer-memory comm; NCCL baseline at N=8), band parity,
# config 4 (61 MP, 31x31) on 8 GPUs and config 5 (batch of 4K frames) sharded across 8 GPUs.
mkdir -p gpurun_out
run() { # N comm tag workload
  local gpus=$(seq -s, 0 $(($1-1)))
  CUDA_VISIBLE_DEVICES=$gpus timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port 2952$1 bench.py --gpus $1 --steps 10 --warmup 3 --comm $2 --no-cpu-baseline --workload ${4:-c3_blind_24mp_k15} --e2e-calls 1 2>/dev/null | grep '^{' > gpurun_out/bench_$3.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$3.json"))
print("$3", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1) if d["e2e"] else None, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
print("   ", {k:round(v,4) for k,v in d["roofline"]["family_ms_per_launch"].items()})
PY
}
echo "=== band parity world 8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29511 tests/band_worker.py gpurun_out/band_report_8.json 2>&1 | grep -E "OK|FAIL|rror" | head
run 8 fused n8_fused
run 8 nccl n8_nccl
run 4 fused n4_fused
run 2 fused n2_fused
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-calls 1 2>/dev/null | grep '^{' > gpurun_out/bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('n1', round(d['value'],1), d['ms_per_step'], d['e2e']['value'])"
run 8 fused c4_n8 c4_blind_61mp_k31
echo "=== config 5: 256 4K frames over 8 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --workload c5_nonblind_4k_kaiser7 --frames 256 2>/dev/null | grep '^{' > gpurun_out/bench_c5_n8.json; python -c "
import json; d=json.load(open('gpurun_out/bench_c5_n8.json')); print('c5 256 frames n8', round(d['value'],1), 'ms/frame/gpu', d['ms_per_step'])"
timeout 600 python bench.py --workload c5_nonblind_4k_kaiser7 --frames 32 2>/dev/null | grep '^{' > gpurun_out/bench_c5_n1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_c5_n1.json')); print('c5 32 frames n1', round(d['value'],1), 'ms/frame', d['ms_per_step'])"
