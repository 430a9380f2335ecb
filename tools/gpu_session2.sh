#!/bin/bash
# GPU session: parity tests + bench (+ optional ncu of the stencil kernels).
mkdir -p gpurun_out
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -30
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_stderr.log | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_stderr.log
if [ "$1" == "ncu" ]; then
echo "=== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_conv|k_gradk|k_update' -s 5 -c 6 -o gpurun_out/prof_r01_$2 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
fi
