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


def oracle_contig_from_recs(seq: bytes, name: bytes, long_name: bytes, bpl: int, recs, lit, goff_c: int, seed: int, gid: int,
                            written):
    """One contig through the C oracle (orc_walk + orc_wrap) given the device's records of that contig — the vectorised
    form of recs_to_muts + c_oracle.mutate_genome for chromosome-sized contigs.  `written` is the ctypes line state that
    orc_wrap carries from contig to contig.  -> (fasta bytes of this contig as the oracle's writer emits them, vcf bytes)"""
    import ctypes as C
    import numpy as np
    from mutation_simulator_b200 import records as R
    from oracle import c_oracle
    lib = c_oracle.lib()
    n = len(recs)
    dt = np.dtype([("key", "<i8"), ("start", "<i8"), ("stop", "<i8"), ("lit_off", "<i8"),
                   ("type", "<i4"), ("reverse", "u1"), ("alt", "u1"), ("pad", "u1", 2)])
    m = np.zeros(max(n, 1), dtype=dt)
    pos = recs["pos"].astype(np.int64); cons = recs["cons"].astype(np.int64); prod = recs["prod"].astype(np.int64)
    typ = recs["type"].astype(np.int32)
    m["key"][:n] = pos
    m["type"][:n] = typ
    start = pos.copy()
    tli = typ == R.T_TLI
    start[tli] = recs["src"][tli] - goff_c
    stop = pos.copy()
    by_prod = (typ == R.T_IN) | (typ == R.T_DU)
    stop[by_prod] = pos[by_prod] + prod[by_prod] - 1
    by_cons = (typ == R.T_DE) | (typ == R.T_TL) | (typ == R.T_IV)
    stop[by_cons] = pos[by_cons] + cons[by_cons] - 1
    stop[tli] = start[tli] + prod[tli] - 1
    m["start"][:n] = start; m["stop"][:n] = stop
    m["reverse"][:n] = (tli & (recs["kind"] == R.K_RC)).astype(np.uint8)
    m["alt"][:n] = np.where(typ == R.T_SN, recs["alt"], 0)
    # insert strings: literal ones from the pool, random ones regenerated from (seed, contig id, position)
    ins = np.flatnonzero(typ == R.T_IN)
    ilen = prod[ins]
    off = np.concatenate(([0], np.cumsum(ilen)[:-1])).astype(np.int64) if len(ins) else np.zeros(0, np.int64)
    pool = np.zeros(int(ilen.sum()) + 16, dtype=np.uint8)
    m["lit_off"][ins] = off
    rnd = recs["kind"][ins] == 6
    if rnd.any():
        p32 = np.ascontiguousarray(recs["pos"][ins][rnd], np.uint32); n32 = np.ascontiguousarray(ilen[rnd], np.uint32)
        o64 = np.ascontiguousarray(off[rnd], np.int64)
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        emu_lib().emu_rand_inserts(C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), C.c_uint32(gid), C.c_int64(len(p32)), P(p32), P(n32), P(o64), P(pool))
    litb = np.asarray(lit, dtype=np.uint8)
    for i in np.flatnonzero(~rnd):
        r = recs[ins[i]]
        pool[off[i]:off[i] + ilen[i]] = litb[int(r["src"]):int(r["src"]) + int(r["prod"])]
    body, vcf = C.c_void_p(), C.c_void_p()
    bl, vl = C.c_int64(), C.c_int64()
    rc = lib.orc_walk(seq, C.c_int64(len(seq)), name, m.ctypes.data_as(C.c_void_p), C.c_int64(n), pool.ctypes.data_as(C.c_void_p),
                      C.byref(body), C.byref(bl), C.byref(vcf), C.byref(vl))
    assert rc == 0, rc
    dst = C.create_string_buffer(bl.value + bl.value // max(1, bpl) + len(long_name) + 8)
    k = lib.orc_wrap(long_name, C.c_int64(len(long_name)), body, C.c_int64(bl.value), C.c_int64(bpl), C.byref(written), dst)
    fa = dst.raw[:k]
    lib.orc_free(body)
    lines = C.string_at(vcf, vl.value) if vl.value else b""
    lib.orc_free(vcf)
    return fa, lines
