"""Shared test helpers: golden-case loading with a minimal FASTA reader.

The reader here is deliberately independent from the product's loader so the
oracle tests do not depend on the package under test."""
from __future__ import annotations

import json
from pathlib import Path

from oracle import pyref

GOLDEN = Path(__file__).resolve().parent / "golden"
MS_CASES = ["args_all", "iupac", "c1_small", "rmt_ranges", "rmt_it", "edges", "tiny"]
IT_CASES = ["it_basic", "rmt_it"]


def read_fasta_simple(path):
    """-> [(name, long_name, upper-cased sequence, bases per line)] (all bytes)"""
    out = []
    name = None
    lines = []
    for raw in Path(path).read_bytes().split(b"\n"):
        if raw.startswith(b">"):
            if name is not None:
                out.append((name, long_name, b"".join(lines).upper(), len(lines[0]) if lines else 0))
            long_name = raw[1:]
            name = long_name.split()[0]
            lines = []
        elif raw:
            lines.append(raw)
    if name is not None:
        out.append((name, long_name, b"".join(lines).upper(), len(lines[0]) if lines else 0))
    return out


def load_muts(case):
    """-> list (per contig, FASTA order) of list[pyref.Mut]"""
    data = json.loads((GOLDEN / case / "muts.json").read_text())
    res = []
    for c in data:
        ms = []
        for m in c["muts"]:
            ms.append(pyref.Mut(key=m["key"], type=m["type"], start=m["start"], stop=m["stop"],
                                reverse=m["reverse"],
                                alt=m["alt"].encode() if "alt" in m else None,
                                insert=m["insert"].encode() if "insert" in m else None))
        res.append(ms)
    return res


def load_case(case):
    d = GOLDEN / case
    contigs = read_fasta_simple(d / "in.fa")
    return d, contigs


def vcf_body(text: bytes) -> bytes:
    return b"".join(l for l in text.splitlines(keepends=True) if not l.startswith(b"#"))


def vcf_head(text: bytes) -> bytes:
    return b"".join(l for l in text.splitlines(keepends=True) if l.startswith(b"#"))
