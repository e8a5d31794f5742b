#!/usr/bin/env python3
"""Per-source-line summary of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py REPORT KERNEL_REGEX [TOP]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; out = []; tot = 0; tots = 0
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 9 and r[0].isdigit():
        I = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
        n, t, s = I(r[7]), I(r[8]), I(r[6])
        if n or s: out.append((n, t, s, cur, r[0], r[1].strip()[:100])); tot += n; tots += s
print(f"kernel {kern}: {tot} warp-instructions, {tots} samples")
for n, t, s, f, ln, src in sorted(out, key=lambda x: -x[0])[:top]:
    print(f"{100*n/max(tot,1):5.1f}% inst  {100*s/max(tots,1):5.1f}% samp  thr/inst {t/max(n,1):5.1f}  {f}:{ln}: {src}")
