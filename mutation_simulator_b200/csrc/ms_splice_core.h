// ms_splice_core.h — per-16-byte-group logic of the splice/emit kernel (K6).
//
// Restates the walk of mutator.py:318-426 fused with the line wrapping of
// fasta_writer.py:49-58 as a pure function  output byte -> value:
//   body byte q of a contig is '\n' iff q mod (bpl+1) == bpl, otherwise it is
//   mutated base b = q - q/(bpl+1), and mutated base b is found through the
//   record with the largest `out` <= b:  inside its payload, or in the copy run
//   that follows it with a constant input shift.
//
// Host/device shared so that tests/emu can run exactly this code on the CPU.
#pragma once
#include "ms_records.h"

namespace ms {

struct SpliceView {
    const uint8_t* genome;  // padded base array (>= 32 readable bytes past the last base)
    const uint8_t* lit;     // literal pool
    const Rec* recs;
    const int64_t* S;       // running length delta before each record (device plan); NULL: Rec.out is filled in (emulation)
    const uint8_t* conv;    // 256-entry tables (global or shared memory)
    const uint8_t* comp;
    Seed seed;              // for K_RAND payloads
    int64_t local_cap = INT64_MAX;   // genome indices outside [0, local_cap) address another GPU's buffer (ms_peer_open)
};

MS_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

// contig-relative output base index of record i's payload: its position shifted by the length deltas before it
MS_HD uint32_t rec_out_of(const SpliceView& v, const Contig& c, int64_t i) {
    return v.S ? (uint32_t)((int64_t)v.recs[i].pos + (v.S[i] - v.S[c.rec_lo])) : v.recs[i].out;
}

// Record cursor: the record governing output base b and its cached fields.
struct Cursor {
    uint32_t pos, gid;  // for K_RAND payloads
    int64_t i;        // absolute record index, rec_lo-1 when b precedes every record
    uint32_t out;     // payload start
    uint32_t prod;    // payload length
    uint32_t next;    // out of the following record (or out_len)
    int64_t run_src;  // genome index of the first base of the trailing copy run
    int64_t src;      // payload source
    uint8_t kind, alt;
};

MS_HD void cursor_load(const SpliceView& v, const Contig& c, int64_t i, Cursor& k) {
    k.i = i;
    k.gid = c.gid; k.pos = 0;
    if (i < c.rec_lo) {
        k.out = 0; k.prod = 0; k.kind = K_NONE; k.alt = 0; k.src = 0;
        k.run_src = c.goff;
    } else {
        const Rec r = v.recs[i];
        k.out = rec_out_of(v, c, i); k.prod = r.prod; k.kind = r.kind; k.alt = r.alt; k.src = r.src; k.pos = r.pos;
        k.run_src = c.goff + (int64_t)r.pos + (int64_t)r.cons;
    }
    k.next = (i + 1 < c.rec_hi) ? rec_out_of(v, c, i + 1) : (uint32_t)c.out_len;
}

// the record governing output base b: the last one of the contig with out <= b (rec_lo - 1 if none)
MS_HD int64_t rec_find(const SpliceView& v, const Contig& c, uint32_t b) {
    int64_t lo = c.rec_lo, hi = c.rec_hi;   // first record with out > b
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (rec_out_of(v, c, mid) <= b) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

// Position the cursor on output base b (b < out_len).
MS_HD void cursor_seek(const SpliceView& v, const Contig& c, uint32_t b, Cursor& k) {
    cursor_load(v, c, rec_find(v, c, b), k);
}

MS_HD void cursor_advance(const SpliceView& v, const Contig& c, uint32_t b, Cursor& k) {
    if (b < k.next) return;
    int64_t i = k.i;
    while (i + 1 < c.rec_hi && rec_out_of(v, c, i + 1) <= b) ++i;
    cursor_load(v, c, i, k);
}

MS_HD uint8_t payload_byte(const SpliceView& v, const Cursor& k, uint32_t rel) {
    switch (k.kind) {
        case K_SNP:  return k.alt;
        case K_LIT:  return v.lit[k.src + rel];
        case K_RAW:  return v.genome[k.src + rel];
        case K_CONV: return v.conv[v.genome[k.src + rel]];
        case K_RC:   return v.comp[v.conv[v.genome[k.src + (int64_t)(k.prod - 1 - rel)]]];
        case K_RAND: return rel < 32u ? cached_insert_base(k.src, rel) : rand_insert_base(v.seed, k.gid, k.pos, rel);
        default:     return '?';
    }
}

MS_HD uint8_t base_at(const SpliceView& v, const Cursor& k, uint32_t b) {
    uint32_t rel = b - k.out;
    if (rel < k.prod) return payload_byte(v, k, rel);
    return v.genome[k.run_src + (int64_t)(rel - k.prod)];
}

// ---- generic path: n (<=16) consecutive body bytes starting at q_lo -----------------
// Writes byte k of the span into bit lanes of w[4] starting at lane `lane0`.
MS_HD void group_slow(const SpliceView& v, const Contig& c, uint32_t q_lo, int n, int lane0, uint32_t w[4]) {
    const uint32_t w1 = (uint32_t)c.bpl + 1u;
    uint32_t line = q_lo / w1;
    uint32_t col = q_lo - line * w1;
    uint32_t b = q_lo - line;  // index of the next base to emit
    Cursor k;
    bool have = false;
    for (int t = 0; t < n; ++t) {
        uint8_t ch;
        if (col == (uint32_t)c.bpl) {
            ch = '\n';
            col = 0;
        } else {
            if (!have) { cursor_seek(v, c, b, k); have = true; } else cursor_advance(v, c, b, k);
            ch = base_at(v, k, b);
            ++b;
            ++col;
        }
        const int lane = lane0 + t;
        w[lane >> 2] |= (uint32_t)ch << (8 * (lane & 3));
    }
}

// ---- fast path: a full 16-byte group that is one shifted copy (plus SNP patches) ------
// Returns false when the group needs the generic path.  LoadWin loads the 32
// bytes at a 16-byte aligned genome index into eight 32-bit words.
template <class LoadWin>
MS_HD bool group_fast(const SpliceView& v, const Contig& c, uint32_t q0, uint32_t w[4], LoadWin load_win) {
    const uint32_t bpl = (uint32_t)c.bpl;
    if (bpl < 16u) return false;  // more than one line break per group: generic path
    const uint32_t w1 = bpl + 1u;
    const uint32_t line = q0 / w1;
    const uint32_t col = q0 - line * w1;
    const uint32_t j = bpl - col;                 // group lane of the line break (if < 16)
    const uint32_t nb = (j < 16u) ? 15u : 16u;    // bases in the group
    const uint32_t bF = q0 - line;
    const uint32_t bL = bF + nb - 1u;

    // governing record of the first base
    const int64_t i = rec_find(v, c, bF);
    int64_t src0;
    uint32_t patch_pos[4];
    uint8_t patch_val[4];
    int npatch = 0;
    bool scan_next = true;
    if (i < c.rec_lo) {
        src0 = c.goff + (int64_t)bF;
    } else {
        const Rec r = v.recs[i];
        const uint32_t rel = bF - rec_out_of(v, c, i);
        if (rel < r.prod) {
            if (r.kind == K_RAW) {
                if (rel + nb > r.prod) return false;
                src0 = r.src + (int64_t)rel;
                scan_next = false;  // the next record starts at or after out+prod > bL
            } else if (r.kind == K_SNP) {
                src0 = c.goff + (int64_t)r.pos;
                patch_pos[0] = 0; patch_val[0] = r.alt; npatch = 1;
            } else {
                return false;
            }
        } else {
            src0 = c.goff + (int64_t)r.pos + (int64_t)r.cons + (int64_t)(rel - r.prod);
        }
    }
    if (scan_next) {
        for (int64_t n = i + 1; n < c.rec_hi; ++n) {
            const uint32_t o = rec_out_of(v, c, n);
            if (o > bL) break;
            const Rec r = v.recs[n];
            if (r.kind != K_SNP || npatch == 4) return false;
            patch_pos[npatch] = o - bF; patch_val[npatch] = r.alt; ++npatch;
        }
    }

    // 16 (unaligned) source bytes -> x[4]
    uint32_t win[8];
    load_win(src0 & ~(int64_t)15, win);
    const uint32_t o = (uint32_t)(src0 & 15);
    const uint32_t ws = o >> 2, bs = (o & 3u) * 8u;
    uint32_t u0, u1, u2, u3, u4;
    switch (ws) {
        case 0:  u0 = win[0]; u1 = win[1]; u2 = win[2]; u3 = win[3]; u4 = win[4]; break;
        case 1:  u0 = win[1]; u1 = win[2]; u2 = win[3]; u3 = win[4]; u4 = win[5]; break;
        case 2:  u0 = win[2]; u1 = win[3]; u2 = win[4]; u3 = win[5]; u4 = win[6]; break;
        default: u0 = win[3]; u1 = win[4]; u2 = win[5]; u3 = win[6]; u4 = win[7]; break;
    }
    uint32_t x[4];
    x[0] = funnel_r(u0, u1, bs); x[1] = funnel_r(u1, u2, bs);
    x[2] = funnel_r(u2, u3, bs); x[3] = funnel_r(u3, u4, bs);

    for (int p = 0; p < npatch; ++p) {
        const uint32_t t = patch_pos[p];
        const uint32_t sh = (t & 3u) * 8u;
        const uint32_t m = ~(0xFFu << sh), val = (uint32_t)patch_val[p] << sh;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t q = 0; q < 4; ++q)
            if (q == (t >> 2)) x[q] = (x[q] & m) | val;
    }

    if (j >= 16u) {
        w[0] = x[0]; w[1] = x[1]; w[2] = x[2]; w[3] = x[3];
        return true;
    }
    // insert '\n' at lane j: lanes < j keep x, lanes > j take the base one lane earlier
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t q = 0; q < 4; ++q) {
        const uint32_t lo = 4u * q;
        if (j >= lo + 4u) { w[q] = x[q]; continue; }
        const uint32_t s = (q == 0) ? (x[0] << 8) : ((x[q] << 8) | (x[q - 1] >> 24));
        if (j < lo) { w[q] = s; continue; }
        const uint32_t t = j - lo;
        const uint32_t keep = t ? (0xFFFFFFFFu >> (32u - 8u * t)) : 0u;
        const uint32_t nl = 0xFFu << (8u * t);
        w[q] = (x[q] & keep) | (0x0Au << (8u * t)) | (s & ~(keep | nl));
    }
    return true;
}

}  // namespace ms
