#!/usr/bin/env python
"""Per-instruction view of one kernel from an .ncu-rep: opcode mix, shared-memory excess wavefronts, stall samples.
usage: ncu_source.py report.ncu-rep kernel_regex [launch_skip]"""
import csv, subprocess, sys

def num(s):
    try: return int(float(s.replace(',', '')))
    except Exception: return 0

def main(rep, regex, skip="0"):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}",
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[0][1][:110] if rows and len(rows[0]) > 1 else "")
    hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) >= len(rows[hi])]
    ix = {h: i for i, h in enumerate(hdr)}
    tot, by_op, conf, samples = 0, {}, [], []
    for r in data:
        src = r[ix["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG")) and "." in op else "")
        n = num(r[ix["Instructions Executed"]])
        tot += n
        by_op[op] = by_op.get(op, 0) + n
        ex = num(r[ix["L1 Wavefronts Shared Excessive"]])
        if ex > 0:
            conf.append((ex, num(r[ix["L1 Wavefronts Shared"]]), num(r[ix["L1 Wavefronts Shared Ideal"]]), n, src[:80]))
        samples.append((num(r[ix["# Samples"]]), src[:80]))
    print("total warp instructions", tot)
    for k, v in sorted(by_op.items(), key=lambda kv: -kv[1])[:18]:
        print(f"  {k:24s} {v:12d} {v / tot:.4f}")
    print("excess shared wavefronts: total", sum(c[0] for c in conf), " (excess, actual, ideal, executions, instruction)")
    for c in sorted(conf, reverse=True)[:10]:
        print("  ", c)
    cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {c: sum(num(r[ix[c]]) for r in data) for c in cols}
    s = sum(agg.values()) or 1
    print("stall samples:", ", ".join(f"{k[6:]}={v / s:.3f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
    print("hottest instructions by samples:")
    for c in sorted(samples, reverse=True)[:8]:
        print("  ", c)

if __name__ == "__main__":
    main(*sys.argv[1:])
