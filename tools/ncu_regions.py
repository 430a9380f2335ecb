#!/usr/bin/env python
"""Stall samples and executed instructions along the SASS of one kernel, in chunks of N instructions (which part of a
role-specialised kernel the time goes to).  usage: ncu_regions.py report.ncu-rep kernel_regex [chunk]"""
import collections, csv, subprocess, sys

def num(s):
    try: return int(float(s.replace(',', '')))
    except Exception: return 0

def main(rep, regex, chunk="150"):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}", "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    seen, d = set(), []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[0] in seen: continue
        seen.add(r[0]); d.append(r)
    tot_s = sum(num(r[ix["# Samples"]]) for r in d) or 1
    tot_i = sum(num(r[ix["Instructions Executed"]]) for r in d) or 1
    print(len(d), "instructions; samples", tot_s, "warp instructions", tot_i)
    CH = int(chunk)
    keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for a in range(0, len(d), CH):
        blk = d[a:a + CH]
        s = sum(num(r[ix["# Samples"]]) for r in blk); n = sum(num(r[ix["Instructions Executed"]]) for r in blk)
        if s < 0.002 * tot_s and n < 0.002 * tot_i: continue
        ops = collections.Counter()
        for r in blk:
            t = r[ix["Source"]].split(); op = t[1] if t[0].startswith('@') else t[0]; ops[op.split('.')[0]] += 1
        st = collections.Counter({k[6:]: sum(num(r[ix[k]]) for r in blk) for k in keys})
        print(f"[{a:5d}] samples {s / tot_s:.3f} instr {n / tot_i:.3f} ops {ops.most_common(4)} stalls {st.most_common(3)}")

if __name__ == "__main__":
    main(*sys.argv[1:])
