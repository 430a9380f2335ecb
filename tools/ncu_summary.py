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


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    for d in data:
        print("kernel:", d[idx["Kernel Name"]][:90])
        for k, short in KEYS:
            if k in idx:
                print(f"  {short:16s} {d[idx[k]]:>18s} {units[idx[k]]}")
        top = sorted(((float(d[idx[s]].replace(",", "") or 0), s) for s in stall), reverse=True)[:5]
        print("  top stalls (warps per issue):", ", ".join(
            f"{s.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}" for v, s in top))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
