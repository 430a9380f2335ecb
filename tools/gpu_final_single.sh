#!/bin/bash
# Round-end single-GPU evidence: tests, smoke, bench (default flags), launch list, full ncu of the hot kernels, workloads.
mkdir -p gpurun_out
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
echo "=== bench (default flags)"; timeout 900 python bench.py 2> gpurun_out/bench_stderr.log | grep '^{' > gpurun_out/bench_final.json; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'], d['cpu_baseline']['value'], d['roofline']['frac'])"
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
echo "=== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
echo "=== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_conv_fft|k_gradk_fft|k_update' -s 6 -c 7 -o gpurun_out/prof_r01_final python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
echo "=== workloads"; bash tools/gpu_workloads.sh
