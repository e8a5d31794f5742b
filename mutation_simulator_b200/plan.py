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


def _type_table(ms, cache):
    """(cdf[7], minlen[7], maxlen[7]) of one MutationSettings: numpy.random.choice's normalised cumulative sum
    (mutator.py:170-174) and the length bounds, in device type order."""
    hit = cache.get(id(ms))
    if hit is not None:
        return hit
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
    cache[id(ms)] = ([float(x) for x in cdf], lo, hi)
    return cache[id(ms)]


def _build_single_range(sim, lengths, local, tables):
    """ARGS / IT-free genomes with many contigs: every chromosome has exactly one range.  Same table as the general
    loop below, computed column-wise (the per-range Python loop costs ~7 us per contig: 1.5 s at C5's 200 k)."""
    chroms = [c for c in sim.chromosomes if c.number in local]
    if len(chroms) < 1024 or any(len(c.range_definitions) != 1 for c in chroms):
        return None
    rds = [c.range_definitions[0] for c in chroms]
    ms_ids = np.fromiter((id(rd.mutation_settings) for rd in rds), dtype=np.int64, count=len(rds))
    uniq, inv = np.unique(ms_ids, return_inverse=True)
    by_id = {}
    for rd in rds:
        by_id.setdefault(id(rd.mutation_settings), rd.mutation_settings)
    has = np.array([bool(by_id[int(u)].has_mutations) for u in uniq])
    total = np.array([sum(by_id[int(u)].mut_rates.values()) if by_id[int(u)].has_mutations else 0.0 for u in uniq], dtype=np.float64)
    start = np.fromiter((rd.start for rd in rds), dtype=np.int64, count=len(rds))
    stop = np.fromiter((rd.stop for rd in rds), dtype=np.int64, count=len(rds))
    num = np.fromiter((local[c.number] for c in chroms), dtype=np.int64, count=len(rds))
    L = np.asarray(lengths, dtype=np.int64)[np.fromiter((c.number for c in chroms), dtype=np.int64, count=len(rds))]
    k = np.trunc(((stop - start) + 1).astype(np.float64) * total[inv]).astype(np.int64)    # int(float64 product), mutator.py:225
    keep = has[inv] & (k > 0) & (stop >= start)
    order = np.argsort(num[keep], kind="stable")
    sel = np.flatnonzero(keep)[order]
    n = int(sel.size)
    arr = (MsRange * max(1, n))()
    if n:
        view = np.frombuffer(arr, dtype=_MSRANGE_DTYPE, count=n)
        view["contig"] = num[sel]; view["start"] = start[sel]; view["stop"] = stop[sel]; view["k"] = k[sel]; view["limit"] = L[sel]
        tabs = [_type_table(by_id[int(u)], tables) if h else ([1.0] * 7, [1] * 7, [1] * 7) for u, h in zip(uniq, has)]
        view["cdf"] = np.array([t[0] for t in tabs], dtype=np.float64)[inv[sel]]
        view["minlen"] = np.array([t[1] for t in tabs], dtype=np.int32)[inv[sel]]
        view["maxlen"] = np.array([t[2] for t in tabs], dtype=np.int32)[inv[sel]]
    return arr, n


def build_ranges(sim, lengths, contig_ids=None):
    """-> (ctypes array of MsRange, n) for the contigs in ``contig_ids`` (global indices, in
    engine order; default all).  ``lengths[i]`` = length of global contig i."""
    ids = list(range(len(lengths))) if contig_ids is None else list(contig_ids)
    local = {g: i for i, g in enumerate(ids)}
    tables = {}     # per MutationSettings object (ARGS mode shares one among all contigs: C5 has 200 k of them)
    fast = _build_single_range(sim, lengths, local, tables)
    if fast is not None:
        return fast
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
            cdf, lo, hi = _type_table(ms, tables)
            rows.append((local[chrom.number], start, stop, k, int(limit), cdf, lo, hi))
    rows.sort(key=lambda r: (r[0], r[1]))
    # fill the ms_range table column-wise through a numpy view (per-field ctypes stores cost seconds at 200 k ranges)
    n = len(rows)
    arr = (MsRange * max(1, n))()
    if n:
        view = np.frombuffer(arr, dtype=_MSRANGE_DTYPE, count=n)
        view["contig"] = [r[0] for r in rows]
        view["start"] = [r[1] for r in rows]
        view["stop"] = [r[2] for r in rows]
        view["k"] = [r[3] for r in rows]
        view["limit"] = [r[4] for r in rows]
        uniq = {}
        idx = np.fromiter((uniq.setdefault(id(r[5]), len(uniq)) for r in rows), dtype=np.int64, count=n)
        first = {}
        for r in rows:
            first.setdefault(id(r[5]), r)
        order = sorted(uniq, key=uniq.get)
        view["cdf"] = np.array([first[u][5] for u in order], dtype=np.float64)[idx]
        view["minlen"] = np.array([first[u][6] for u in order], dtype=np.int32)[idx]
        view["maxlen"] = np.array([first[u][7] for u in order], dtype=np.int32)[idx]
    return arr, n


_MSRANGE_DTYPE = np.dtype({"names": ["contig", "start", "stop", "k", "limit", "cdf", "minlen", "maxlen"],
                           "formats": ["<u4", "<u4", "<u4", "<u4", "<i8", ("<f8", 7), ("<i4", 7), ("<i4", 7)],
                           "offsets": [MsRange.contig.offset, MsRange.start.offset, MsRange.stop.offset, MsRange.k.offset,
                                       MsRange.limit.offset, MsRange.cdf.offset, MsRange.minlen.offset, MsRange.maxlen.offset],
                           "itemsize": __import__("ctypes").sizeof(MsRange)})
