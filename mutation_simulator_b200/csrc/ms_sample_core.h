// ms_sample_core.h — per-candidate logic of the sampling kernels (K2/K3/K4b),
// host/device shared.  Restates mutator.py:166-182 (type), :229-265 (length /
// stop), :204-209 (blocking reach) and :429-463 (SNP substitution).
#pragma once
#include "ms_records.h"

namespace ms {

// One RMT range that has mutations (rmt.py RangeDefinition + MutationSettings),
// flattened by the host.  Ranges are sorted by (contig, start).
struct Range {
    int64_t gstart;     // genome index of the range's first base
    int64_t cand_lo;    // first candidate slot of this range
    int64_t store_lo;   // first slot of this range's sort buckets in the bucket store
    int64_t limit;      // contig-relative end (exclusive) an SV may extend to: contig length,
                        // or the start of the next blocked (None) range — SURVEY.md Q3 stance
    uint32_t start;     // contig-relative, inclusive
    uint32_t stop;      // contig-relative, inclusive
    uint32_t k;         // candidates: int(((stop-start)+1)*sum(rates)) computed on the host in float64 (mutator.py:225)
    uint32_t n;         // population of util.py:104: stop-(k-1)*d-start
    uint32_t contig;    // local contig index
    uint32_t bucket_lo; // first sort bucket
    uint32_t nb;        // number of sort buckets
    uint32_t bscale;    // floor(2^32 * nb / n): bucket of value v = mulhi32(v, bscale) (monotone, < nb)
    double cdf[7];      // cumulative mut_chances in canonical type order (numpy.random.choice, mutator.py:170-174)
    int32_t minlen[7];
    int32_t maxlen[7];
};

struct RangeParams {
    double cdf[7];
    int32_t minlen[7];
    int32_t maxlen[7];
    int64_t limit;
};

// Type and length of the candidate at contig-relative `pos`.
// type = T_DEAD for an inversion that does not fit (mutator.py:243-244).
template <class RP>
MS_HD void draw_type_len(Seed seed, uint32_t gid, uint32_t pos, const RP& rp, uint8_t& type, uint32_t& len) {
    const U4 r = draw(seed, gid, P_TYPE_LEN, pos);
    const double u = unit_double(r.x, r.y);
    int t = 0;
    while (t < 6 && rp.cdf[t] <= u) ++t;  // searchsorted(cdf, u, side='right')
    const uint64_t span = (uint64_t)(rp.maxlen[t] - rp.minlen[t]) + 1;
    uint32_t l = (uint32_t)rp.minlen[t] + (uint32_t)bounded(u64_of(r.z, r.w), span);  // randint(min, max)
    const int64_t lim = rp.limit;
    switch (t) {
        case T_SN: l = 1; break;
        case T_IV:
            if ((int64_t)pos + rp.maxlen[T_IV] >= lim - 1) { type = T_DEAD; len = 0; return; }
            break;
        case T_IN: break;  // never clipped (the stop only encodes the insert length)
        case T_DU: case T_DE: case T_TL:
            if ((int64_t)pos + l > lim) l = (uint32_t)(lim - pos);
            break;
        default: l = 0; break;  // TLI placeholder, linked later
    }
    type = (uint8_t)t;
    len = l;
}

// Exclusive end of the blocked stretch an accepted candidate leaves behind
// (mutator.py:204-209), contig-relative.  TLI placeholders have stop = 0 in the
// reference, hence the absolute 1 + block.
MS_HD int64_t block_reach(uint8_t type, uint32_t pos, uint32_t len, const int32_t* block) {
    switch (type) {
        case T_SN: case T_IN: return (int64_t)pos + 1 + block[type];
        case T_TLI: return 1 + (int64_t)block[T_TLI];
        case T_DEAD: return 0;
        default: return (int64_t)pos + len + block[type];
    }
}

// mutator.py:429-463.  `ref` is already IUPAC-converted.
MS_HD uint8_t draw_snp(Seed seed, uint32_t gid, uint32_t pos, uint8_t ref, double p_ti, const uint8_t* trans) {
    const U4 r = draw(seed, gid, P_SNP, pos);
    const double u = unit_double(r.x, r.y);
    if (u <= p_ti) return trans[ref];
    return transversion(ref, r.z & 1u);
}

// mutator.py:466-471: base j of the insert at `pos`; 64 bases per Philox block.
MS_HD uint8_t insert_base(const U4& blk, uint32_t j) {
    const uint32_t wsel = (j >> 4) & 3u;
    const uint32_t wv = wsel == 0 ? blk.x : wsel == 1 ? blk.y : wsel == 2 ? blk.z : blk.w;
    const uint32_t two = (wv >> ((j & 15u) * 2u)) & 3u;
    return (uint8_t)("ATGC"[two]);
}

// mutator.py:307-316
MS_HD bool draw_tl_reverse(Seed seed, uint32_t gid, uint32_t tli_pos, uint32_t len) {
    const U4 r = draw(seed, gid, P_TL_REV, tli_pos);
    return (r.x & 1u) != 0 && len >= 2;
}

}  // namespace ms
