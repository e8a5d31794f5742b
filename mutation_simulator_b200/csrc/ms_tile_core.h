// ms_tile_core.h — index logic of the splice/emit kernel (K6, second generation), host/device shared.
//
// The walk of mutator.py:318-426 fused with the line wrapping of fasta_writer.py:49-58, per 16 KiB tile of the output
// file image.  Records are split into two position-ordered streams by the plan pass:
//   * SvRec  (32 B)  every record that moves bases (IN, DE, IV, DU, TL, TLI, IT segments): where its payload starts in
//                    the output, how long it is, and where the copy run that follows it starts in the input;
//   * Snp8   (8 B)   single-base substitutions: output index + substituted base.  SNPs move nothing, so between two
//                    consecutive SvRecs the output is ONE shifted copy of the input ("run", ~290 bases at human rates).
// A tile is assembled in shared memory as the final file bytes (bases AND line breaks) and leaves with one bulk store:
//   prep   one thread per SvRec: span start, payload end, and the offset of its run in the staged input span
//   walk   one thread per CELL (= one line of the file, or 60 bytes of a long line): find the governing SvRec (sorted
//          starts), then copy the cell's bases run by run — each piece is one contiguous byte copy with a constant
//          shift, so no line break is ever inserted into a vector and run boundaries need no special pass; payloads
//          that have to be generated (inserts, inversions, translocation inserts) are queued as jobs
//   jobs   half a warp per queued payload piece;  snp: one thread per Snp8 scatters the substituted base
// This header holds everything that is index arithmetic, so tests/emu runs exactly this code on the CPU against the
// reference's golden files; the CUDA kernel (ms_apply.cu) supplies the memory operations.
#pragma once
#include "ms_records.h"

namespace ms {

#if defined(__CUDACC__) && defined(MS_TILE_RAND_OOL)
__device__ uint8_t rand_base_ool(Seed seed, uint32_t gid, uint32_t pos, uint32_t j);
#endif

struct alignas(16) SvRec {
    uint32_t out;      // contig-relative output base index where the payload starts
    uint32_t prod;     // payload bases
    uint32_t run_in;   // contig-relative input index where the trailing copy run starts (pos + cons)
    uint32_t pos;      // input position of the record (key of random inserts)
    int64_t  src;      // payload source (Rec.src)
    uint32_t kind;     // Kind
    uint32_t pad;
};
static_assert(sizeof(SvRec) == 32, "SvRec must be 32 bytes");

struct alignas(8) Snp8 { uint32_t out; uint32_t alt; };   // alt: substituted base in the low byte
static_assert(sizeof(Snp8) == 8, "Snp8 must be 8 bytes");

#ifndef MS_TILE_SHIFT
#define MS_TILE_SHIFT 14
#endif
constexpr int TL_TILE_SHIFT = MS_TILE_SHIFT;
constexpr int TL_TILE = 1 << TL_TILE_SHIFT;    // file bytes per tile
constexpr int TL_STAGE_CAP = TL_TILE + TL_TILE / 8 + 512;   // staged input span (tile + what deletions skip + alignment slack)
constexpr int TL_SV_CAP = TL_TILE / 32;        // SvRecs per tile the staged path handles (incl. the governing one)
constexpr int TL_POOL = TL_TILE / 2;           // bytes of SvRec + Snp8 per tile
constexpr int TL_JOB_CAP = TL_TILE / 64;       // queued generated-payload pieces per tile
constexpr int TL_PIECE_CAP = TL_TILE / 32 + TL_TILE / 128;   // queued copy pieces per tile
constexpr int32_t TL_DIRECT = INT32_MIN;       // source lies outside the staged span: read from global memory

enum : uint32_t { PD_GOV_VIRTUAL = 1u, PD_FALLBACK = 2u };

// Everything a tile's CTA needs to start, precomputed by a fully parallel kernel (k_piece_desc).
struct alignas(16) PieceDesc {
    int64_t f_lo, f_hi;      // file bytes of the piece (tile ∩ contig body)
    int64_t body_off;        // file offset of body byte 0 of the contig
    int64_t goff;            // genome index of base 0 of the contig
    int64_t in_lo;           // 16-byte aligned genome index where the staged input span starts
    int64_t sv_lo;           // first SvRec to load: the governing one (or, if that is virtual, the first inside the tile)
    int64_t snp_lo;          // first Snp8 inside the tile
    uint32_t in_bytes;       // staged bytes (multiple of 16, <= TL_STAGE_CAP)
    uint32_t n_sv;           // SvRecs of the tile incl. the governing one (slot 0)
    uint32_t n_snp;
    uint32_t flags;          // PD_*
    uint32_t bpl, gid, cidx;
    uint32_t b_lo, b_hi;     // mutated bases [b_lo, b_hi) live in this piece
    uint32_t line_lo, col_lo;  // f_lo - body_off = line_lo * (bpl + 1) + col_lo
    uint32_t cb;             // b_lo % bpl
    int32_t  k0;             // (b_lo + b_lo / bpl) - (f_lo - body_off): 1 if the piece starts on a line break, else 0
    float rcp_w1, rcp_bpl;   // tl_rcp(bpl + 1), tl_rcp(bpl)
    uint32_t pad[3];
};
static_assert(sizeof(PieceDesc) == 128, "PieceDesc layout");

struct alignas(16) TileRec {
    uint32_t pe;     // tile-relative base index where the payload ends and the run starts (clipped into the span)
    int32_t  ro;     // offset in the staged span of the run's first base, or TL_DIRECT
    int32_t  po;     // K_RAW: offset in the staged span of the payload's first (clipped) base, or TL_DIRECT
    uint32_t kind;
};

// Shared-memory view of one tile (device: __shared__; emu: heap).
struct TileShared {
    uint8_t* stage;      // staged input span [in_lo, in_lo + in_bytes) (+32 readable bytes)
    uint8_t* image;      // file bytes [f_lo & ~15, ...) of the piece
    SvRec* sv;           // [n_sv]
    const Snp8* snp;     // entry i of the tile = snp[snp_skip + i]
    uint32_t* rs;        // [n_sv + 1] tile-relative span start of each SvRec (ascending), rs[n_sv] = bases in the piece
    TileRec* dv;         // [n_sv]
};

// what the kernel (or the emulation) needs besides the tile itself
struct TileView {
    const uint8_t* genome;
    const uint8_t* lit;
    const uint8_t* conv;
    const uint8_t* comp;
    Seed seed;
};

// x / d for x < d + 2^15 (tile-relative numerators), exact.  rcp = tl_rcp(d).
MS_HD float tl_rcp(uint32_t d) { return d >= (1u << 20) ? 0.f : (1.0f / (float)d) * 0.99999976f; }
MS_HD uint32_t div_small(uint32_t x, uint32_t d, float rcp) {
    if (d >= (1u << 20)) return x >= d ? 1u : 0u;            // x < 2 d
    uint32_t t = (uint32_t)((float)x * rcp);                 // x < 2^21: exact in float; t is the quotient or one less
    if (x - t * d >= d) ++t;
    return t;
}

// ---- prep: derived fields of SvRec j -----------------------------------------------------------------------------
MS_HD void tile_prep_rec(const PieceDesc& d, const TileShared& sh, uint32_t j) {
    const SvRec r = sh.sv[j];
    const uint32_t b_lo = d.b_lo, b_hi = d.b_hi;
    uint32_t nxt = (j + 1u < d.n_sv) ? sh.sv[j + 1u].out : b_hi;
    if (nxt > b_hi) nxt = b_hi;
    const uint32_t pe_abs = r.out + r.prod;
    const uint32_t s = r.out > b_lo ? r.out : b_lo;                  // span start
    uint32_t pe = pe_abs < s ? s : pe_abs;
    if (pe > nxt) pe = nxt;                                          // payload end clipped into [s, nxt]
    sh.rs[j] = s - b_lo;
    TileRec t;
    t.pe = pe - b_lo; t.kind = r.kind; t.ro = TL_DIRECT; t.po = TL_DIRECT;
    if (nxt > pe) {          // trailing run [pe, nxt): base pe  <->  input run_in + (pe - pe_abs)
        const int64_t rel = d.goff + (int64_t)r.run_in + (int64_t)(pe - pe_abs) - d.in_lo;
        if (rel >= 0 && rel + (int64_t)(nxt - pe) <= (int64_t)d.in_bytes) t.ro = (int32_t)rel;
    }
    if (r.kind == K_RAW && pe > s) {   // raw payload [s, pe): base s  <->  src + (s - out)
        const int64_t rel = r.src + (int64_t)(s - r.out) - d.in_lo;
        if (rel >= 0 && rel + (int64_t)(pe - s) <= (int64_t)d.in_bytes) t.po = (int32_t)rel;
    }
    sh.dv[j] = t;
}

// largest j in [0, n) with rs[j] <= r   (rs ascending, rs[0] == 0)
MS_HD uint32_t tile_find(const uint32_t* rs, uint32_t n, uint32_t r) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1u) {
        const uint32_t mid = (lo + hi) >> 1;
        if (rs[mid] <= r) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- walk: the piece is cut into CELLS, one thread each ------------------------------------------------------------
// Lines of up to 96 bytes: a cell is one line of the file (its bases + the line break), so the shift between source
// and destination is constant along the whole cell and nothing ever has to be inserted into a vector of bases; with
// the usual odd line pitch (61, 71, 81 bytes) the threads of a warp also fall on different shared-memory banks.
// Longer lines are cut into 60-byte cells counted from the piece start (a cell then holds at most one line break).
struct TileGeom {
    uint32_t e, img_end;     // image bytes [e, img_end) belong to the piece (image byte 0 = file offset f_lo & ~15)
    uint32_t w1, bpl;
    uint32_t cw;             // cell pitch in bytes
    int32_t  x_org;          // image offset of cell 0 (row-aligned: may lie before e)
    uint32_t n_cells;
    float rcp_w1;
};

MS_HD TileGeom tile_geom(const PieceDesc& d) {
    TileGeom g;
    const int64_t g0 = d.f_lo & ~(int64_t)15;
    g.e = (uint32_t)(d.f_lo - g0); g.img_end = (uint32_t)(d.f_hi - g0);
    g.bpl = d.bpl; g.w1 = d.bpl + 1u;
    g.rcp_w1 = d.rcp_w1;
    const uint32_t len = g.img_end - g.e;
    if (g.w1 <= 96u) {
        g.cw = g.w1;
        g.x_org = (int32_t)g.e - (int32_t)d.col_lo;
        g.n_cells = len ? div_small(d.col_lo + len - 1u, g.w1, g.rcp_w1) + 1u : 0u;
    } else {
        g.cw = 60u;
        g.x_org = (int32_t)g.e;
        g.n_cells = (len + 59u) / 60u;
    }
    return g;
}

// genome byte g: from the staged span when it lies inside (an inversion's source is the stretch it replaces, so it
// usually does), else from global memory
MS_HD uint8_t tile_genome_byte(const PieceDesc& d, const TileShared& sh, const TileView& v, int64_t g) {
    const int64_t rel = g - d.in_lo;
    return (rel >= 0 && rel < (int64_t)d.in_bytes) ? sh.stage[rel] : v.genome[g];
}

// byte r of SvRec j's payload region, r = tile-relative base index (generated payloads; raw ones are copied)
MS_HD uint8_t tile_payload_byte(const PieceDesc& d, const TileShared& sh, const TileView& v, uint32_t j, uint32_t r) {
    const SvRec q = sh.sv[j];
    const uint32_t rel = (r + d.b_lo) - q.out;
    switch (q.kind) {
        case K_RAW:  return tile_genome_byte(d, sh, v, q.src + rel);
        case K_LIT:  return v.lit[q.src + rel];
        case K_CONV: return v.conv[tile_genome_byte(d, sh, v, q.src + rel)];
        case K_RC:   return v.comp[v.conv[tile_genome_byte(d, sh, v, q.src + (int64_t)(q.prod - 1u - rel))]];
        case K_RAND:
            if (rel < 32u) return cached_insert_base(q.src, rel);
#if defined(__CUDA_ARCH__) && defined(MS_TILE_RAND_OOL)
            return rand_base_ool(v.seed, d.gid, q.pos, rel);      // Philox out of line: rare, and it would cost registers here
#else
            return rand_insert_base(v.seed, d.gid, q.pos, rel);
#endif
        default:     return (uint8_t)'?';
    }
}

// n bases starting at tile-relative base r go to image offset x.  Ops supplies the memory operations:
//   copy_stage(x, s_off, n)   n bytes from the staged span at s_off
//   copy_global(x, g, n)      n bytes from genome index g
//   job(x, j, r, n)           n generated payload bytes of SvRec j from base r (queued; produced by tile_payload_byte)
// j: the SvRec governing r on entry, the one governing r + n (or the next) on exit.
template <class Ops>
MS_HD void tile_emit_bases(const PieceDesc& d, const TileShared& sh, Ops& ops, uint32_t x, uint32_t r, uint32_t n, uint32_t& j) {
    while (n) {
        while (r >= sh.rs[j + 1u]) ++j;
        const TileRec t = sh.dv[j];
        const uint32_t end = sh.rs[j + 1u];
        uint32_t m;
        if (r < t.pe) {                                        // inside the payload
            m = t.pe - r < n ? t.pe - r : n;
            if (t.kind == K_RAW) {
                const uint32_t rs = sh.rs[j];
                if (t.po != TL_DIRECT) ops.copy_stage(x, (uint32_t)(t.po + (int32_t)(r - rs)), m);
                else { const SvRec q = sh.sv[j]; ops.copy_global(x, q.src + ((int64_t)r + (int64_t)d.b_lo - (int64_t)q.out), m); }
            } else {
                ops.job(x, j, r, m);
            }
        } else {                                               // in the trailing copy run
            m = end - r < n ? end - r : n;
            if (t.ro != TL_DIRECT) ops.copy_stage(x, (uint32_t)(t.ro + (int32_t)(r - t.pe)), m);
            else {
                const SvRec q = sh.sv[j];
                ops.copy_global(x, d.goff + (int64_t)q.run_in + ((int64_t)r + (int64_t)d.b_lo - (int64_t)(q.out + q.prod)), m);
            }
        }
        x += m; r += m; n -= m;
    }
}

// One cell: its bases (through tile_emit_bases) and line breaks (ops.put(x, '\n')).
template <class Ops>
MS_HD void tile_cell(const PieceDesc& d, const TileShared& sh, const TileGeom& g, Ops& ops, uint32_t cell) {
    int32_t xs = g.x_org + (int32_t)(cell * g.cw);
    int32_t xe = xs + (int32_t)g.cw;
    if (xs < (int32_t)g.e) xs = (int32_t)g.e;
    if (xe > (int32_t)g.img_end) xe = (int32_t)g.img_end;
    if (xe <= xs) return;
    uint32_t x = (uint32_t)xs, left = (uint32_t)(xe - xs);
    const uint32_t dd = x - g.e;                                // offset from f_lo
    uint32_t dl = div_small(d.col_lo + dd, g.w1, g.rcp_w1);
    uint32_t col = d.col_lo + dd - dl * g.w1;
    uint32_t r = dd - dl;                                       // tile-relative index of the next base
    uint32_t j = 0u;
    bool have = false;
    while (left) {
        if (col == g.bpl) {
            ops.put(x, (uint8_t)'\n');
            ++x; --left; col = 0u;
            continue;
        }
        const uint32_t nb = g.bpl - col < left ? g.bpl - col : left;
        if (!have) { j = tile_find(sh.rs, d.n_sv, r); have = true; }
        tile_emit_bases(d, sh, ops, x, r, nb, j);
        x += nb; r += nb; left -= nb; col += nb;
    }
}

// ---- snp: offset from f_lo of the file byte of output base `out` ------------------------------------------------
MS_HD uint32_t tile_snp_offset(const PieceDesc& d, uint32_t out, float rcp_bpl) {
    const uint32_t r = out - d.b_lo;
    return r + (uint32_t)d.k0 + div_small(d.cb + r, d.bpl, rcp_bpl);
}

// ---- the piece descriptor (k_piece_desc; one thread per piece) -----------------------------------------------------
// sv / snp: the two streams; [sv_c0, sv_c1) / [snp_c0, snp_c1) the contig's entries.  f_lo / f_hi already clipped to
// the contig body.
MS_HD PieceDesc tile_describe(const Contig& k, uint32_t cidx, int64_t f_lo, int64_t f_hi, const SvRec* sv, int64_t sv_c0, int64_t sv_c1,
                              const Snp8* snp, int64_t snp_c0, int64_t snp_c1) {
    PieceDesc d{};
    d.f_lo = f_lo; d.f_hi = f_hi; d.body_off = k.body_off; d.goff = k.goff;
    d.bpl = (uint32_t)k.bpl; d.gid = k.gid; d.cidx = cidx;
    const uint32_t w1 = d.bpl + 1u;
    const uint32_t q_lo = (uint32_t)(f_lo - k.body_off), q_hi = (uint32_t)(f_hi - k.body_off);
    d.line_lo = q_lo / w1; d.col_lo = q_lo - d.line_lo * w1;
    d.b_lo = q_lo - d.line_lo; d.b_hi = q_hi - q_hi / w1;
    d.cb = d.b_lo % d.bpl;
    d.k0 = (int32_t)((int64_t)d.b_lo + (int64_t)(d.b_lo / d.bpl) - (int64_t)q_lo);
    d.rcp_w1 = tl_rcp(w1); d.rcp_bpl = tl_rcp(d.bpl);
    // governing SvRec: the last one with out <= b_lo
    int64_t lo = sv_c0, hi = sv_c1;          // first index with out > b_lo
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (sv[mid].out <= d.b_lo) lo = mid + 1; else hi = mid; }
    const int64_t gov = lo - 1;
    // first index with out >= b_hi: a tile holds a few dozen SvRecs, so gallop from the start instead of bisecting
    // the whole contig again (each probe is a dependent load: this kernel is pure latency)
    int64_t e = lo; hi = sv_c1;
    for (int64_t step = 64; e + step < hi; step <<= 1) {
        if (sv[e + step].out < d.b_hi) e += step + 1; else { hi = e + step; break; }
    }
    while (e < hi) { const int64_t mid = (e + hi) >> 1; if (sv[mid].out < d.b_hi) e = mid + 1; else hi = mid; }
    const bool virt = gov < sv_c0;
    d.flags = virt ? PD_GOV_VIRTUAL : 0u;
    d.sv_lo = virt ? lo : gov;
    const int64_t n_sv = (e - lo) + 1;       // tile records + the governing slot
    lo = snp_c0; hi = snp_c1;                // first snp with out >= b_lo
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (snp[mid].out < d.b_lo) lo = mid + 1; else hi = mid; }
    const int64_t s0 = lo;
    hi = snp_c1;
    for (int64_t step = 256; lo + step < hi; step <<= 1) {
        if (snp[lo + step].out < d.b_hi) lo += step + 1; else { hi = lo + step; break; }
    }
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (snp[mid].out < d.b_hi) lo = mid + 1; else hi = mid; }
    d.snp_lo = s0;
    const int64_t n_snp = lo - s0;
    // pool layout: SvRecs first (32 B each), then the Snp8s from an even stream index (16-byte aligned bulk copy)
    const int64_t pool = 32 * n_sv + 8 * (n_snp + 2);
    (void)pool;
    if (n_sv > TL_SV_CAP) d.flags |= PD_FALLBACK;
    d.n_sv = (uint32_t)(n_sv > 0x7fffffff ? 0x7fffffff : n_sv);
    d.n_snp = (uint32_t)(n_snp > 0x7fffffff ? 0x7fffffff : n_snp);
    // staged span: the contiguous, monotone stretch of input the tile's copy runs read — from the source of its first
    // base to the source of its last base.  A tile that starts inside a raw far copy (interchromosomal segment,
    // it_mutator.py:133-137) stages that payload's source instead when it covers most of the tile.
    auto run_src = [&](int64_t ri, uint32_t b) -> int64_t {       // source of base b governed by SvRec ri (not in a payload)
        if (ri < sv_c0) return k.goff + (int64_t)b;
        const SvRec q = sv[ri];
        const uint32_t rel = b - q.out;
        return k.goff + (int64_t)q.run_in + (rel >= q.prod ? (int64_t)(rel - q.prod) : 0);
    };
    int64_t s_lo, s_hi;
    bool raw_span = false;
    if (!virt) {
        const SvRec q = sv[gov];
        const uint32_t rel = d.b_lo - q.out;
        if (q.kind == K_RAW && rel < q.prod) {
            const uint32_t take = (q.prod - rel) < (d.b_hi - d.b_lo) ? (q.prod - rel) : (d.b_hi - d.b_lo);
            if (2u * take >= d.b_hi - d.b_lo) { s_lo = q.src + rel; s_hi = s_lo + take; raw_span = true; }
        }
    }
    if (!raw_span) {
        const uint32_t t = d.b_hi > d.b_lo ? d.b_hi - 1u : d.b_lo;
        s_lo = run_src(gov, d.b_lo);
        s_hi = run_src(e - 1 >= sv_c0 && e - 1 >= gov ? e - 1 : gov, t) + 1;
        if (s_hi > k.goff + k.len) s_hi = k.goff + k.len;
    }
    d.in_lo = s_lo & ~(int64_t)15;
    int64_t nb = s_hi > d.in_lo ? ((s_hi + 15) & ~(int64_t)15) - d.in_lo : 0;
    if (nb > (int64_t)TL_STAGE_CAP) nb = TL_STAGE_CAP;
    d.in_bytes = (uint32_t)nb;
    return d;
}

}  // namespace ms
