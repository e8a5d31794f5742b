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


# ---- GPU-side helpers -------------------------------------------------------------
def engine_for(contigs, device=0):
    """contigs: [(name, long_name, seq_upper, bpl)] -> (Engine with the genome resident, goff, lengths)"""
    import numpy as np
    from mutation_simulator_b200 import records as R
    from mutation_simulator_b200.engine import Engine
    eng = Engine(device)
    seqs = [c[2] for c in contigs]
    genome, goff = R.pack_genome(seqs)
    lens = [len(s) for s in seqs]
    eng.upload_genome(genome[:int(goff[-1])], lens, [c[3] for c in contigs], [c[1] for c in contigs], [c[0] for c in contigs])
    return eng, genome, goff, np.array(lens)


_EMU = None


def emu_lib():
    """tests/emu/emu.cpp compiled with g++ (host build of the kernels' shared host/device headers)."""
    global _EMU
    if _EMU is None:
        import ctypes as C
        import subprocess
        here = Path(__file__).resolve().parent
        src, lib = here / "emu" / "emu.cpp", here / "emu" / "_build" / "libemu.so"
        lib.parent.mkdir(exist_ok=True)
        deps = [src] + list((here.parent / "mutation_simulator_b200" / "csrc").glob("*.h"))
        if not lib.exists() or lib.stat().st_mtime < max(p.stat().st_mtime for p in deps):
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(lib), str(src)])
        _EMU = C.CDLL(str(lib))
    return _EMU


def rand_insert(seed, gid, pos, n) -> bytes:
    """The random insert a K_RAND record expands to (Philox, keyed by seed / global contig id / position)."""
    import ctypes as C
    out = C.create_string_buffer(max(1, n))
    emu_lib().emu_rand_insert(C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), C.c_uint32(gid), C.c_uint32(pos), C.c_uint32(n), out)
    return out.raw[:n]


def recs_to_muts(recs, lit, goff, seed=0, gids=None):
    """Decode device records back into per-contig lists of pyref.Mut (for the oracle)."""
    from mutation_simulator_b200 import records as R
    n_contigs = len(goff) - 1
    out = [[] for _ in range(n_contigs)]
    lit = bytes(lit.tobytes()) if hasattr(lit, "tobytes") else bytes(lit)
    for r in recs:
        c, pos, t = int(r["contig"]), int(r["pos"]), int(r["type"])
        name = R.TYPE_NAME[t]
        m = pyref.Mut(key=pos, type=name, start=pos, stop=pos)
        if t == R.T_SN:
            m.alt = bytes([int(r["alt"])])
        elif t == R.T_IN:
            if int(r["kind"]) == 6:   # K_RAND: generated where consumed
                m.insert = rand_insert(seed, c if gids is None else int(gids[c]), pos, int(r["prod"]))
            else:
                m.insert = lit[int(r["src"]):int(r["src"]) + int(r["prod"])]
            m.stop = pos + int(r["prod"]) - 1
        elif t in (R.T_DE, R.T_TL, R.T_IV):
            m.stop = pos + int(r["cons"]) - 1
        elif t == R.T_DU:
            m.stop = pos + int(r["prod"]) - 1
        elif t == R.T_TLI:
            m.start = int(r["src"]) - int(goff[c])
            m.stop = m.start + int(r["prod"]) - 1
            m.reverse = int(r["kind"]) == R.K_RC
        out[c].append(m)
    return out
