#!/usr/bin/env python3
"""GFF/GTF gene annotation -> RMT file that blocks every gene (reference: data/scripts/gff_genes_2_rmt.py:1-38).

    python -m mutation_simulator_b200.tools.gff_genes_2_rmt annotation.gff3 0.01

writes annotation.rmt: `std / it None / sn <rate>`, then one `chr N #<seqid>` section per sequence (in order of first
appearance, like the reference) with a `start-stop None` line per feature of type `gene`.  Gene intervals overlap in
every real annotation; this package's planner lets blocked ranges win (mutation_simulator_b200/plan.py), so the
result is usable as it is.  Unlike the reference script this one streams the input (gzip allowed) instead of reading
it whole, and reports malformed lines instead of dying on them.
"""
from __future__ import annotations

import gzip
import sys
from pathlib import Path

USAGE = "USAGE: python3 gff3_genes_2_rmt.py [gff_file] [rate]"


def gene_intervals(lines):
    """{seqid: [(start, stop), ...]} for feature type 'gene' (case-insensitive), seqids in order of appearance."""
    by_seq = {}
    for no, line in enumerate(lines, 1):
        if not line.strip() or line.startswith("#"):
            continue
        f = line.rstrip("\n").split("\t")
        if len(f) < 5:
            raise ValueError(f"line {no}: expected at least 5 tab-separated columns")
        if f[2].lower() == "gene":
            by_seq.setdefault(f[0], []).append((f[3], f[4]))
    return by_seq


def rmt_text(by_seq, rate) -> str:
    out = [f"std\nit None\nsn {rate}\n\n"]
    for i, (seqid, ivs) in enumerate(by_seq.items()):
        out.append(f"chr {i+1} #{seqid}\n")
        out.extend(f"{a}-{b} None\n" for a, b in ivs)
    return "".join(out)


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2:
        print(USAGE)
        return 1
    path = Path(argv[0])
    try:
        rate = float(argv[1])
    except ValueError:
        print(USAGE)
        return 1
    try:
        opener = gzip.open if path.suffix == ".gz" else open
        with opener(path, "rt") as fh:
            by_seq = gene_intervals(fh)
    except FileNotFoundError:
        print(f"ERROR: cannot find {path}")
        return 1
    except ValueError as e:
        print(f"ERROR: {path}: {e}")
        return 1
    path.with_suffix(".rmt").write_text(rmt_text(by_seq, rate))
    return 0


if __name__ == "__main__":
    sys.exit(main())
