"""BEDPE output of interchromosomal swaps (bedpe_writer.py:36-55): one row per pair of
consecutive breakpoints, written once from each partner's side, no header."""
from __future__ import annotations

import numpy as np


class BedpeWriterError(Exception):
    """Raised when the writer can not write to a file."""


def rows(chrom: str, bp, chrom_len: int, partner: str, bpp, partner_len: int) -> bytes:
    bp = np.asarray(bp, dtype=np.int64)
    bpp = np.asarray(bpp, dtype=np.int64)
    n = bp.size
    if n == 0:
        return b""
    a = np.append(bp, chrom_len)         # an odd count closes the last row with the contig lengths
    b = np.append(bpp, partner_len)
    idx = np.arange(0, n, 2)
    if n % 2 == 0:
        idx = idx[idx != n - 1]
    lines = [f"{chrom}\t{a[i]}\t{a[i+1]}\t{partner}\t{b[i]}\t{b[i+1]}\n" for i in idx]
    return "".join(lines).encode("latin-1")


def breakpoints_from_rows(text, names, lengths):
    """Inverse of :func:`rows` / BedpeWriter.write (bedpe_writer.py:36-55): BEDPE text -> (partners, breakpoints) as
    ITMutator keeps them — partners[c] = index of c's partner, breakpoints[c] = {"self": [...], "partner": [...]}.
    Every swap is listed from both sides (it_mutator.py:205-211 calls the writer once per member of a pair); the rows of
    one side are enough, both are accepted.  A row that ends at the contig lengths is the closing row of an odd count,
    not a breakpoint (breakpoints lie in [1, len-1], it_mutator.py:108-111)."""
    if isinstance(text, (bytes, bytearray)):
        text = bytes(text).decode("latin-1")
    index = {(n.decode("latin-1") if isinstance(n, (bytes, bytearray)) else str(n)): i for i, n in enumerate(names)}
    partners, bps = {}, {}
    for ln, line in enumerate(text.splitlines(), 1):
        if not line or line[0] == "#":
            continue
        f = line.split("\t")
        if len(f) != 6 or f[0] not in index or f[3] not in index:
            raise ValueError(f"BEDPE line {ln}: expected chrom1 start1 stop1 chrom2 start2 stop2 with known contigs")
        a, b = index[f[0]], index[f[3]]
        if partners.setdefault(a, b) != b:
            raise ValueError(f"BEDPE line {ln}: {f[0]} is paired with more than one contig")
        d = bps.setdefault(a, {"self": [], "partner": []})
        s1, e1, s2, e2 = int(f[1]), int(f[2]), int(f[4]), int(f[5])
        d["self"].append(s1); d["partner"].append(s2)
        if not (e1 == int(lengths[a]) and e2 == int(lengths[b])):
            d["self"].append(e1); d["partner"].append(e2)
    for a, b in list(partners.items()):
        if partners.get(b, a) != a:
            raise ValueError("BEDPE: pairing is not symmetric")
        if b not in partners:          # only one side listed: mirror it
            partners[b] = a
            bps[b] = {"self": list(bps[a]["partner"]), "partner": list(bps[a]["self"])}
    return partners, {c: {k: np.asarray(v, dtype=np.uint32) for k, v in d.items()} for c, d in bps.items()}


class BedpeWriter:
    def __init__(self, fname):
        try:
            self._f = open(fname, "wb")
        except OSError as e:
            raise BedpeWriterError(f"Cannot write to BEDPE file {fname} {e}")

    def __del__(self):
        self.close()

    def close(self):
        if getattr(self, "_f", None) is not None and not self._f.closed:
            self._f.close()

    def write_header(self):
        self._f.write(b"#chrom1\tstart1\tstop1\tchrom2\tstart2\tstop2\n")

    def write(self, chrom, bp_chrom, chrom_len_pre_it, partner, bp_partner, partner_len_pre_it):
        self._f.write(rows(chrom, bp_chrom, chrom_len_pre_it, partner, bp_partner, partner_len_pre_it))
