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
