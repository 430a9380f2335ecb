#!/bin/bash
# Short round-end evidence refresh (one GPU): default bench line, launch list, full ncu of one inner step's hot kernels.
mkdir -p gpurun_out
timeout 600 python bench.py 2> gpurun_out/bench_stderr.log | grep '^{' > gpurun_out/bench_final.json
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'], d['cpu_baseline']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['gpu_launches'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_conv_fft|k_gradk_fft|k_update' -s 5 -c 6 -o gpurun_out/prof_r01_final python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
