"""SimulationSettings -> the flat range table the GPU samples from (ms_range).

Keeps the reference's exact arithmetic where it decides counts: the candidate
count of a range is ``int(((stop-start)+1) * sum(rates))`` in float64 with the
summation order of the settings dict (mutator.py:225, :160), and the type
probabilities are numpy.random.choice's normalised cumulative sum
(mutator.py:170-174)."""
from __future__ import annotations

import numpy as np

from ._lib import MsRange
from .mut_types import DEVICE_CODE, DEVICE_ORDER, MutType


class RangeOverlapError(ValueError):
    """Two RMT ranges that both carry mutations overlap (the reference crashes later
    with an uncaught ValueError from random.sample)."""


def block_list(sim) -> list:
    return [int(sim.mut_block[t]) for t in DEVICE_ORDER]


def p_transition(titv: float) -> float:
    return titv * (1 / (titv + 1))   # mutator.py:436


def build_ranges(sim, lengths, contig_ids=None):
    """-> (ctypes array of MsRange, n) for the contigs in ``contig_ids`` (global indices, in
    engine order; default all).  ``lengths[i]`` = length of global contig i."""
    ids = list(range(len(lengths))) if contig_ids is None else list(contig_ids)
    local = {g: i for i, g in enumerate(ids)}
    rows = []
    for chrom in sim.chromosomes:
        if chrom.number not in local:
            continue
        L = int(lengths[chrom.number])
        rds = chrom.range_definitions
        # exclusive end an SV may reach: start of the next blocked (None) range, else the contig end
        limits = [L] * len(rds)
        nxt = L
        for i in range(len(rds) - 1, -1, -1):
            limits[i] = nxt
            if not rds[i].mutation_settings.has_mutations and rds[i].stop >= rds[i].start:
                nxt = rds[i].start
        prev_stop = -1
        blocked_until = -1     # last base covered by a blocked (None) range seen so far (ranges are sorted by start)
        for rd, limit in zip(rds, limits):
            ms = rd.mutation_settings
            if not ms.has_mutations:
                blocked_until = max(blocked_until, int(rd.stop))
                continue
            # Overlapping explicit ranges (gene annotations overlap; every RMT shipped with the reference has them)
            # make the reference's gap filler produce negative-length ranges and die in random.sample.  Here blocked
            # ranges win: a range with mutations is cut back to the part no None range covers; nothing is left of a
            # negative-length filler.
            start = max(int(rd.start), blocked_until + 1)
            stop = min(int(rd.stop), int(limit) - 1)
            if stop < start:
                continue
            total = sum(ms.mut_rates.values())
            k = int(((stop - start) + 1) * total)
            if k <= 0:
                continue
            if start <= prev_stop:
                raise RangeOverlapError(f"Range {rd.start+1}-{rd.stop+1} of chromosome {chrom.number+1} overlaps the previous range")
            prev_stop = stop
            p = np.zeros(7)
            for t, c in ms.mut_chances.items():
                p[DEVICE_CODE[t]] = c
            cdf = np.cumsum(p)
            cdf /= cdf[-1]
            lo = [1] * 7
            hi = [1] * 7
            for t in DEVICE_ORDER[1:]:
                src = MutType.TL if t is MutType.TLI else t
                if ms.mut_lengs and src in ms.mut_lengs["min"] and src in ms.mut_lengs["max"]:
                    lo[DEVICE_CODE[t]] = int(ms.mut_lengs["min"][src])
                    hi[DEVICE_CODE[t]] = int(ms.mut_lengs["max"][src])
            rows.append((local[chrom.number], start, stop, k, int(limit), cdf, lo, hi))
    rows.sort(key=lambda r: (r[0], r[1]))
    arr = (MsRange * max(1, len(rows)))()
    for a, (c, start, stop, k, limit, cdf, lo, hi) in zip(arr, rows):
        a.contig, a.start, a.stop, a.k, a.limit = c, start, stop, k, limit
        for t in range(7):
            a.cdf[t], a.minlen[t], a.maxlen[t] = float(cdf[t]), lo[t], hi[t]
    return arr, len(rows)
