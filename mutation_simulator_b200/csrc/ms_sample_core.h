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
    uint32_t top_levels; // levels of the bucket tree whose nodes span more than SPLIT_LEAF buckets (k_split_top); 0 for most ranges
    uint32_t pad;
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

// ---- k distinct positions of a range, generated in order (util.py:94-109: random.sample(range(n), k), sorted) --------
// The value span [0, n) of a range is cut into nb equal sort buckets.  A uniform k-subset of [0, n) puts a
// multivariate-hypergeometric number of values into each bucket; those counts come from a binary split tree over
// the buckets (the samples of a node go left with a hypergeometric law), each split a pure function of
// (range key, node) — so every bucket knows its count without any value having been drawn or scattered.  Inside a
// bucket the values are then a uniform c-subset of its span (k_sort_emit).
//
// hypergeom(): number of marked items among `n` drawn without replacement from `N` of which `K` are marked.
//   variance >= 900: rounded normal with the exact mean and variance (Sheppard-corrected for the rounding); splits are
//     near p = 1/2, so the skewness (1-2p)/sigma is ~0 and the pmf is matched to O(1/sigma^2) < 1e-3 relative;
//   otherwise: inversion by walking outwards from the mode with the exact pmf ratios.  The mode's pmf comes from log
//     factorials (Stirling's series where every argument is >= 16: truncation < 3e-12; lgamma otherwise) — the
//     normalisation is good to ~2e-5 for N ~ 2^31 and ~1e-9 for the node sizes that take this path when the range is
//     dense; the ratios are formed in single precision (a walk of ~1.6 sigma <= 50 steps accumulates < 1e-5 relative).
//     A shortfall of the walked mass falls back to the mode.
MS_HD double log_fact_stirling(double x, double lx) {   // ln x!, x >= 16, lx = ln x
    const double i = 1.0 / x, i2 = i * i;
    return x * lx - x + 0.5 * lx + 0.9189385332046727 + i * (1.0 / 12.0 - i2 * (1.0 / 360.0 - i2 * (1.0 / 1260.0)));
}
MS_HD double log_choose(double n, double k) {
    const double m = n - k;
    if (k >= 16.0 && m >= 16.0) {
        const double ln = log(n), lk = log(k), lm = log(m);
        return log_fact_stirling(n, ln) - log_fact_stirling(k, lk) - log_fact_stirling(m, lm);
    }
    return lgamma(n + 1.0) - lgamma(k + 1.0) - lgamma(m + 1.0);
}

// u >= pmf(m): walk outwards from the mode, up and down alternately, until the accumulated mass passes u
template <class F>
MS_HD int64_t hypergeom_walk(double u, double pm, int64_t m, int64_t lo, int64_t hi, double dN, double dK, double dn) {
    const F fK = (F)dK, fn = (F)dn, fR = (F)(dN - dK - dn);   // R = N - K - n
    const F one = (F)1, tiny = (F)1e-19;
    F xb = (F)m, xa = (F)m;      // [a, b] walked so far
    F pa = (F)pm, pb = pa;       // pmf at a and at b
    double acc = pm;
    int64_t a = m, b = m;
    bool up = b < hi, down = a > lo;
    while (up || down) {
        if (up) {               // p(x+1)/p(x) = (K-x)(n-x) / ((x+1)(N-K-n+x+1))
            pb *= ((fK - xb) * (fn - xb)) / ((xb + one) * (fR + xb + one));
            ++b; xb += one;
            acc += (double)pb;
            if (u < acc) return b;
            up = b < hi && pb > tiny;
        }
        if (down) {             // p(x-1)/p(x) = x (N-K-n+x) / ((K-x+1)(n-x+1))
            pa *= (xa * (fR + xa)) / ((fK - xa + one) * (fn - xa + one));
            --a; xa -= one;
            acc += (double)pa;
            if (u < acc) return a;
            down = a > lo && pa > tiny;
        }
    }
    return m;
}

MS_HD uint32_t hypergeom(const U4& r, uint32_t N, uint32_t K, uint32_t n) {
    if (n == 0u || K == 0u) return 0u;
    if (K >= N) return n;
    if (n >= N) return K;
    const int64_t lo = (int64_t)n + (int64_t)K - (int64_t)N > 0 ? (int64_t)n + (int64_t)K - (int64_t)N : 0;
    const int64_t hi = n < K ? n : K;
    if (lo == hi) return (uint32_t)lo;
    const double dN = (double)N, dK = (double)K, dn = (double)n;
    const double p = dK / dN;
    const double mean = dn * p;
    const double var = mean * (1.0 - p) * ((dN - dn) / (dN - 1.0));
    if (var >= 900.0) {
        const double u1 = ((double)(u64_of(r.x, r.y) >> 11) + 1.0) * (1.0 / 9007199254740992.0);   // (0, 1]
        const double u2 = unit_double(r.z, r.w);
        const double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
        double x = floor(mean + sqrt(var - 1.0 / 12.0) * z + 0.5);     // rounding adds 1/12 to the variance
        if (x < (double)lo) x = (double)lo;
        if (x > (double)hi) x = (double)hi;
        return (uint32_t)x;
    }
    int64_t m = (int64_t)(((dn + 1.0) * (dK + 1.0)) / (dN + 2.0));
    if (m < lo) m = lo;
    if (m > hi) m = hi;
    const double pm = exp(log_choose(dK, (double)m) + log_choose(dN - dK, dn - (double)m) - log_choose(dN, dn));
    // walk outwards, up and down alternately: any fixed order of the support is an inversion as long as it is the same
    // for every u; a side stops once its pmf is below 1e-19
    const double u = unit_double(r.x, r.y);
    double acc = pm;
    if (u < acc) return (uint32_t)m;
    // (single precision holds the walked values exactly only below 2^24)
    if (hi < (1 << 23)) return (uint32_t)hypergeom_walk<float>(u, pm, m, lo, hi, dN, dK, dn);
    return (uint32_t)hypergeom_walk<double>(u, pm, m, lo, hi, dN, dK, dn);
}

// samples of the bucket-tree node [lo, hi) that go to its left child [lo, mid): N values in the node, K of them left
MS_HD uint32_t split_left(Seed range_key, uint32_t lo, uint32_t hi, uint32_t N, uint32_t K, uint32_t n) {
    const U4 r = philox4x32_10(U4{lo, hi, 0x53504C54u, 0u}, range_key.k0, range_key.k1);
    return hypergeom(r, N, K, n);
}

// j-th value drawn in redraw round `round` for sort bucket `bl` of a range: uniform in [0, width).  One Philox block
// serves draws 2i and 2i+1.
MS_HD U4 bucket_block(Seed range_key, uint32_t bl, uint32_t round, uint32_t pair) {
    return philox4x32_10(U4{bl, round, pair, 0x56414C55u}, range_key.k0, range_key.k1);
}
MS_HD uint32_t bucket_value_of(const U4& r, uint32_t j, uint32_t width) {
    return (uint32_t)bounded((j & 1u) ? u64_of(r.z, r.w) : u64_of(r.x, r.y), (uint64_t)width);
}
MS_HD uint32_t bucket_value(Seed range_key, uint32_t bl, uint32_t round, uint32_t j, uint32_t width) {
    return bucket_value_of(bucket_block(range_key, bl, round, j >> 1), j, width);
}

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
MS_HD uint8_t snp_of_block(const U4& r, uint8_t ref, double p_ti, const uint8_t* trans) {   // r = draw(seed, gid, P_SNP, pos)
    const double u = unit_double(r.x, r.y);
    if (u <= p_ti) return trans[ref];
    return transversion(ref, r.z & 1u);
}
MS_HD uint8_t draw_snp(Seed seed, uint32_t gid, uint32_t pos, uint8_t ref, double p_ti, const uint8_t* trans) {
    return snp_of_block(draw(seed, gid, P_SNP, pos), ref, p_ti, trans);
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
