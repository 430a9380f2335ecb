#!/bin/bash
# Multi-GPU session: row-band parity vs single GPU, then the strong-scaling bench at N GPUs.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "=== band parity (world $N)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29511 tests/band_worker.py gpurun_out/band_report_$N.json 2>&1 | grep -v "^W10\|^\*\*\*\|OMP_NUM" | tail -25
echo "=== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/bench_${N}_stderr.log | tee gpurun_out/bench_$N.json | cut -c1-1500
tail -5 gpurun_out/bench_${N}_stderr.log | cut -c1-600
