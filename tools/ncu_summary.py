#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small text table for profiles/."""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_act_%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp_insts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"), ("lts__t_bytes.sum", "l2_bytes")]


FAMILIES = [("k_chain_fft<", "", "conv_adj"), ("k_conv_fft<", ", 0>", "conv_fwd"), ("k_conv_fft<", ", 1>", "conv_adj"), ("k_conv<", ", 0>", "conv_fwd"),
            ("k_conv<", ", 1>", "conv_adj"), ("k_update", "", "update"), ("k_gradk_fft<", "", "gradk"), ("k_gradk<", "", "gradk")]


def to_bytes(val, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(val.replace(",", "")) * scale[unit]


def family_of(name):
    for a, b, fam in FAMILIES:
        if a in name and (not b or b in name.split("(")[0]):
            return fam
    return None


def main(rep, json_out=None, workload=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    traffic = {}
    for d in data:
        fam = family_of(d[idx["Kernel Name"]])
        if fam and "dram__bytes_read.sum" in idx:
            rd, wr = idx["dram__bytes_read.sum"], idx["dram__bytes_write.sum"]
            traffic.setdefault(fam, []).append(to_bytes(d[rd], units[rd]) + to_bytes(d[wr], units[wr]))
        print("kernel:", d[idx["Kernel Name"]][:90])
        for k, short in KEYS:
            if k in idx:
                print(f"  {short:16s} {d[idx[k]]:>18s} {units[idx[k]]}")
        top = sorted(((float(d[idx[s]].replace(",", "") or 0), s) for s in stall), reverse=True)[:5]
        print("  top stalls (warps per issue):", ", ".join(
            f"{s.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}" for v, s in top))
        print()


    if json_out:
        import json
        import pathlib
        sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
        import bench                                   # the stamp bench.py checks before quoting these numbers
        json.dump({"workload": workload, "source": rep.split("/")[-1], "csrc_sha16": bench.csrc_hash(),
                   "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches), ncu --set full",
                   "per_launch_dram_bytes": {f: sum(v) / len(v) for f, v in traffic.items()},
                   "launches_captured": {f: len(v) for f, v in traffic.items()}}, open(json_out, "w"), indent=1)


if __name__ == "__main__":
    # ncu_summary.py REPORT [JSON_OUT WORKLOAD]
    main(*sys.argv[1:4])
