"""ctypes wrapper around oracle/ms_oracle.c (the plain-C restatement).

TEST INFRASTRUCTURE — NOT A PRODUCT PATH (see ms_oracle.c header)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libms_oracle.so"
TYPE_CODE = {"SN": 0, "IN": 1, "DE": 2, "IV": 3, "DU": 4, "TL": 5, "TLI": 6}
TYPE_NAME = {v: k for k, v in TYPE_CODE.items()}


class OrcMut(C.Structure):
    _fields_ = [("key", C.c_int64), ("start", C.c_int64), ("stop", C.c_int64), ("lit_off", C.c_int64),
                ("type", C.c_int32), ("reverse", C.c_uint8), ("alt", C.c_uint8), ("pad", C.c_uint8 * 2)]


class OrcRange(C.Structure):
    _fields_ = [("start", C.c_int64), ("stop", C.c_int64), ("k", C.c_int64), ("cdf", C.c_double * 7),
                ("minlen", C.c_int32 * 7), ("maxlen", C.c_int32 * 7)]


def build(force: bool = False) -> Path:
    src = HERE / "ms_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        LIB.parent.mkdir(exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", str(LIB), str(src)])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.orc_walk.restype = C.c_int
        _lib.orc_wrap.restype = C.c_int64
        _lib.orc_sample_contig.restype = C.c_int64
        _lib.orc_mutate_contig.restype = C.c_int64
        _lib.orc_free.restype = None
        _lib.orc_free.argtypes = [C.c_void_p]
    return _lib


def _take(ptr, n):
    buf = C.string_at(ptr, n) if n else b""
    lib().orc_free(ptr)
    return buf


def walk(seq: bytes, name: bytes, muts) -> tuple[bytes, bytes]:
    """muts: iterable of oracle.pyref.Mut -> (mutated bases, vcf body)"""
    ms = sorted(muts, key=lambda m: m.key)
    arr = (OrcMut * max(1, len(ms)))()
    lit = bytearray()
    for i, m in enumerate(ms):
        arr[i].key, arr[i].start, arr[i].stop = m.key, m.start, m.stop
        arr[i].type = TYPE_CODE[m.type]
        arr[i].reverse = 1 if m.reverse else 0
        if m.type == "SN":
            arr[i].alt = m.alt[0]
        if m.type == "IN":
            arr[i].lit_off = len(lit)
            lit += m.insert
    body, vcf = C.c_void_p(), C.c_void_p()
    bl, vl = C.c_int64(), C.c_int64()
    rc = lib().orc_walk(seq, C.c_int64(len(seq)), name, arr, C.c_int64(len(ms)), bytes(lit),
                        C.byref(body), C.byref(bl), C.byref(vcf), C.byref(vl))
    if rc:
        raise RuntimeError(f"orc_walk failed: {rc}")
    return _take(body, bl.value), _take(vcf, vl.value)


def mutate_genome(contigs, muts_per_contig) -> tuple[bytes, bytes]:
    """Same contract as oracle.pyref.mutate_genome, through the C restatement."""
    fa, vcf = [], []
    written = C.c_int64(0)
    for (name, long_name, seq, bpl), muts in zip(contigs, muts_per_contig):
        body, lines = walk(seq, name, muts)
        dst = C.create_string_buffer(len(body) + len(body) // max(1, bpl) + len(long_name) + 8)
        n = lib().orc_wrap(long_name, C.c_int64(len(long_name)), body, C.c_int64(len(body)),
                           C.c_int64(bpl), C.byref(written), dst)
        fa.append(dst.raw[:n])
        vcf.append(lines)
    return b"".join(fa), b"".join(vcf)


def make_ranges(ranges):
    """ranges: list of dict(start, stop, k, cdf[7], minlen[7], maxlen[7])"""
    arr = (OrcRange * max(1, len(ranges)))()
    for i, r in enumerate(ranges):
        arr[i].start, arr[i].stop, arr[i].k = r["start"], r["stop"], r["k"]
        for t in range(7):
            arr[i].cdf[t] = r["cdf"][t]
            arr[i].minlen[t] = r["minlen"][t]
            arr[i].maxlen[t] = r["maxlen"][t]
    return arr


def sample_contig(seq: bytes, ranges, block, min_dist, titv, seed):
    """-> (numpy structured view of mutations, literal pool)"""
    arr = make_ranges(ranges)
    blk = (C.c_int32 * 7)(*block)
    muts, lit, litn = C.POINTER(OrcMut)(), C.c_void_p(), C.c_int64()
    n = lib().orc_sample_contig(seq, C.c_int64(len(seq)), arr, C.c_int32(len(ranges)), blk,
                                C.c_int32(min_dist), C.c_double(titv), C.c_uint64(seed),
                                C.byref(muts), C.byref(lit), C.byref(litn))
    if n < 0:
        raise ValueError(f"orc_sample_contig failed: {n}")
    dt = np.dtype([("key", "<i8"), ("start", "<i8"), ("stop", "<i8"), ("lit_off", "<i8"),
                   ("type", "<i4"), ("reverse", "u1"), ("alt", "u1"), ("pad", "u1", 2)])
    out = np.frombuffer(C.string_at(muts, n * C.sizeof(OrcMut)), dtype=dt).copy() if n else np.zeros(0, dt)
    pool = _take(lit, litn.value)
    lib().orc_free(muts)
    return out, pool


def mutate_contig(seq: bytes, name: bytes, header: bytes, bpl: int, ranges, block, min_dist, titv, seed,
                  written: int = 0):
    """End-to-end for one contig -> (fasta bytes, vcf bytes, counts[7], written)"""
    arr = make_ranges(ranges)
    blk = (C.c_int32 * 7)(*block)
    fa, vcf = C.c_void_p(), C.c_void_p()
    fl, vl = C.c_int64(), C.c_int64()
    w = C.c_int64(written)
    counts = (C.c_int64 * 7)()
    n = lib().orc_mutate_contig(seq, C.c_int64(len(seq)), name, header, C.c_int64(len(header)), C.c_int64(bpl),
                                arr, C.c_int32(len(ranges)), blk, C.c_int32(min_dist), C.c_double(titv),
                                C.c_uint64(seed), C.byref(w), C.byref(fa), C.byref(fl), C.byref(vcf), C.byref(vl),
                                counts)
    if n < 0:
        raise ValueError(f"orc_mutate_contig failed: {n}")
    return _take(fa, fl.value), _take(vcf, vl.value), list(counts), w.value
