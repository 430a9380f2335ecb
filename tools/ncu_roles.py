#!/usr/bin/env python
"""Stall samples and executed instructions of the chain kernel per ROLE: the SASS between two setmaxnreg instructions is
one role's loop (csrc/rltv_chain_fft.cuh).  usage: ncu_roles.py report.ncu-rep [kernel_regex]"""
import collections, csv, subprocess, sys

def num(s):
    try: return int(float(s.replace(',', '')))
    except Exception: return 0

def main(rep, regex="k_chain_fft"):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}", "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    seen, d = set(), []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[0] in seen: continue
        seen.add(r[0]); d.append(r)
    keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(num(r[ix["# Samples"]]) for r in d) or 1
    marks = [i for i, r in enumerate(d) if "USETMAXREG" in r[ix["Source"]]]
    print(len(d), "SASS instructions; stall samples", tot, "; setmaxnreg at", marks)
    bounds = [0] + marks + [len(d)]
    for a, b in zip(bounds[:-1], bounds[1:]):
        blk = d[a:b]
        s = sum(num(r[ix["# Samples"]]) for r in blk); n = sum(num(r[ix["Instructions Executed"]]) for r in blk)
        if s < 0.002 * tot: continue
        src0 = blk[0][ix["Source"]].replace(";", " ").split()
        regs = str(int(src0[-1], 16)) if "USETMAXREG" in blk[0][ix["Source"]] else "-"
        ops = collections.Counter()
        for r in blk:
            t = r[ix["Source"]].split(); op = t[1] if t[0].startswith('@') else t[0]; ops[op.split('.')[0]] += num(r[ix["Instructions Executed"]])
        st = collections.Counter({k[6:]: sum(num(r[ix[k]]) for r in blk) for k in keys})
        print(f"[{a:5d},{b:5d}) regs {regs:>5s} samples {s / tot:.3f} warp-instr {n:10d} top ops {[(k, round(v / max(n, 1), 2)) for k, v in ops.most_common(5)]}")
        print("        stalls", [(k, round(v / max(s, 1), 2)) for k, v in st.most_common(8)])

if __name__ == "__main__":
    main(*sys.argv[1:])
