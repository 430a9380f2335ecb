#!/bin/bash
# Multi-GPU session: row-band parity vs single GPU, then the strong-scaling bench at N GPUs.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "=== band parity (world $N)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29511 tests/band_worker.py gpurun_out/band_report_$N.json 2>&1 | grep -v "^W10\|^\*\*\*\|OMP_NUM" | tail -25
echo "=== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --e2e-calls 1 2> gpurun_out/bench_${N}_stderr.log | tee gpurun_out/bench_$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d[\"n_gpus\"], round(d[\"value\"],1), d[\"ms_per_step\"], \"e2e\", d[\"e2e\"][\"value\"], d[\"clocks\"], {k:round(v,4) for k,v in d[\"roofline\"][\"family_ms_per_launch\"].items()})"
tail -5 gpurun_out/bench_${N}_stderr.log | cut -c1-600
