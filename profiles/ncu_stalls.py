#!/usr/bin/env python3
"""Stall samples per source line of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_stalls.py REPORT KERNEL_REGEX [TOP]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None; cur = None; agg = {}
cols = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_not_selected", "stall_math", "stall_branch_resolving"]
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        I = lambda k: int(d[k]) if d.get(k, "").strip().isdigit() else 0
        a = agg.setdefault((cur, int(r[0]), r[1].strip()[:80]), [0] * (len(cols) + 2))
        a[0] += I("# Samples"); a[1] += I("Instructions Executed")
        for i, c in enumerate(cols): a[2 + i] += I(c)
tot = sum(a[0] for a in agg.values())
print(f"kernel {kern}: {tot} samples; " + ", ".join(f"{c[6:]} {sum(a[2+i] for a in agg.values())}" for i, c in enumerate(cols)))
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{100*a[0]/tot:5.1f}% samp  bar {a[2]:6d} lsb {a[3]:6d} ssb {a[4]:6d} wait {a[5]:6d}  inst {a[1]:10d}  {k[0]}:{k[1]}: {k[2]}")
