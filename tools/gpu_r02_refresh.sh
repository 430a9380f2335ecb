#!/bin/bash
# Final refresh of the round-2 single-GPU evidence with the committed sources: ncu capture -> stamped traffic -> the bench
# line that quotes it -> launch list -> the whole GPU test suite -> the other configs.
mkdir -p gpurun_out
echo "=== ncu full"; rm -f gpurun_out/prof_r02_final.ncu-rep; timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_chain_fft|k_gradk_fft|k_update' -s 8 -c 8 -o gpurun_out/prof_r02_final python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02_final.log 2>&1; tail -1 gpurun_out/ncu_r02_final.log
python tools/ncu_summary.py gpurun_out/prof_r02_final.ncu-rep gpurun_out/ncu_traffic_r02.json c3_blind_24mp_k15 > gpurun_out/ncu_r02_final_summary.txt 2>&1; cp gpurun_out/ncu_traffic_r02.json profiles/ncu_traffic_r02.json
python tools/ncu_roles.py gpurun_out/prof_r02_final.ncu-rep > gpurun_out/ncu_r02_chain_roles.txt 2>&1
echo "=== bench (driver flags)"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bench_r02_final.err | grep '^{' > gpurun_out/bench_r02_final_n1.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r02_final_n1.json")); r=d["roofline"]
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d["clocks"], "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],3))
print("step frac", round(r["step"]["frac_of_hbm_all_gpus"],4), "dominant", r["kernel"], round(r["frac"],4), "traffic", r["traffic"], {k:round(v,4) for k,v in r["family_ms_per_launch"].items()})
PY
echo "=== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1500 2>&1 | tail -4
echo "=== workloads"
for W in c1_nonblind_512_g5 c2_blind_2mp_k9 c5_nonblind_4k_kaiser7 c4_blind_61mp_k31; do
  timeout 900 python bench.py --workload $W --steps 6 --warmup 3 --no-cpu-baseline --e2e-calls 1 2>gpurun_out/wl_r02_$W.err | grep '^{' > gpurun_out/bench_r02_workload_$W.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r02_workload_$W.json")); r=d["roofline"]
    print("$W", "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "step hbm frac", round(r["step"]["frac_of_hbm_all_gpus"],3), d["clocks"]["reasons"])
except Exception as e:
    print("$W failed", e)
PY
done
