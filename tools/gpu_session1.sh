#!/bin/bash
# First GPU session: smoke, parity tests, micro-benchmarks, bench, ncu launch list.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40
echo "=== microbench"; timeout 120 tools/microbench | tee gpurun_out/microbench.json
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_stderr.log | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_stderr.log
echo "=== sanitizer"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -k "stages_against_oracle and (15 or 31 or 7) or whiteness" --timeout 500 2>&1 | tail -15
echo "=== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -3 gpurun_out/bench_under_ncu.log
