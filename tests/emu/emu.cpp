// emu.cpp — CPU emulation of the per-thread logic of the CUDA kernels.
//
// TEST TOOLING ONLY.  The build container has no GPU, so the index math that the
// kernels share with this file through the host/device headers
// (ms_rng.h, ms_splice_core.h, ms_vcf_core.h) is exercised here against the
// oracle before GPU time is spent.  It is compiled by tests/test_emu.py with g++
// and is never imported by the product package (which has no CPU fallback).
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>
#include "../../mutation_simulator_b200/csrc/ms_rng.h"
#include "../../mutation_simulator_b200/csrc/ms_records.h"
#include "../../mutation_simulator_b200/csrc/ms_splice_core.h"
#include "../../mutation_simulator_b200/csrc/ms_vcf_core.h"
#include "../../mutation_simulator_b200/csrc/ms_sample_core.h"
#include "../../mutation_simulator_b200/csrc/ms_tile_core.h"

using namespace ms;

extern "C" {

void emu_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    U4 r = philox4x32_10(U4{ctr[0], ctr[1], ctr[2], ctr[3]}, key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

void emu_prp(uint64_t seed, uint32_t contig, uint32_t purpose, uint64_t idx, uint32_t n, uint32_t count, uint32_t* out) {
    Prp p = make_prp(make_seed(seed), contig, purpose, idx, n);
    for (uint32_t j = 0; j < count; ++j) out[j] = prp_apply(p, j);
}

void emu_ndigits(uint32_t count, const uint32_t* v, uint32_t* out) { for (uint32_t i = 0; i < count; ++i) out[i] = ndigits(v[i]); }

// count draws of hypergeom(N, K, n), draw i from Philox counter (i, 0, 0, 0) under `seed`
void emu_hypergeom(uint64_t seed, uint32_t N, uint32_t K, uint32_t n, uint32_t count, uint32_t* out) {
    const Seed s = make_seed(seed);
    for (uint32_t i = 0; i < count; ++i) out[i] = hypergeom(philox4x32_10(U4{i, 0, 0, 0}, s.k0, s.k1), N, K, n);
}

// Bucket counts of one range through the split tree, the way k_range_keys + k_split_level compute them on the device:
// nb buckets with value spans starting at vlo[0..nb) (vlo[nb] = n), k samples in total.
void emu_split_counts(uint64_t seed, uint32_t gid, uint32_t start, uint32_t nb, const uint32_t* vlo, uint32_t k, uint32_t* cnt) {
    const U4 kk = draw(make_seed(seed), gid, P_RANGE_KEY, start);
    const Seed key{kk.x, kk.y};
    std::vector<std::pair<uint32_t, uint32_t>> nodes{{0u, nb}};
    cnt[0] = k;
    while (!nodes.empty()) {
        std::vector<std::pair<uint32_t, uint32_t>> next;
        for (auto [lo, hi] : nodes) {
            if (hi - lo < 2u) continue;
            const uint32_t mid = (lo + hi) >> 1;
            const uint32_t left = split_left(key, lo, hi, vlo[hi] - vlo[lo], vlo[mid] - vlo[lo], cnt[lo]);
            cnt[mid] = cnt[lo] - left;
            cnt[lo] = left;
            next.push_back({lo, mid});
            next.push_back({mid, hi});
        }
        nodes.swap(next);
    }
}

// Layout + splice + VCF for a whole (small) genome.
//   genome: concatenated contigs, contig c at goff[c] (16-aligned, >=32 pad bytes at the end)
//   recs:   sorted by (contig, pos), `out` not yet filled
// Returns 0, fills fasta (size *fasta_len) and vcf.
int emu_apply(const uint8_t* genome, int32_t n_contigs, const int64_t* goff, const int64_t* clen, const int32_t* bpl,
              const uint8_t* headers, const int64_t* hdr_off, const uint8_t* names, const int64_t* name_off,
              Rec* recs, int64_t n_recs, const uint8_t* lit,
              uint8_t* fasta, int64_t fasta_cap, int64_t* fasta_len, uint8_t* vcf, int64_t vcf_cap, int64_t* vcf_len,
              int64_t tile_bytes, int64_t* n_fast_groups, int64_t* n_slow_groups) {
    Tables tab; fill_tables(tab);
    std::vector<Contig> C(n_contigs);
    // delta scan + per-contig record ranges (K5)
    int64_t ri = 0, file = 0, piece_total = 0;
    for (int c = 0; c < n_contigs; ++c) {
        Contig& k = C[c];
        memset(&k, 0, sizeof(k));
        k.goff = goff[c]; k.len = clen[c]; k.bpl = bpl[c] > 0 ? bpl[c] : 60; k.gid = c;
        k.hdr_src = hdr_off[c]; k.hdr_len = (int32_t)(hdr_off[c + 1] - hdr_off[c]);
        k.name_src = name_off[c]; k.name_len = (int32_t)(name_off[c + 1] - name_off[c]);
        k.rec_lo = ri;
        int64_t delta = 0;
        while (ri < n_recs && recs[ri].contig == (uint32_t)c) {
            recs[ri].out = (uint32_t)((int64_t)recs[ri].pos + delta);
            delta += (int64_t)recs[ri].prod - (int64_t)recs[ri].cons;
            ++ri;
        }
        k.rec_hi = ri;
        k.out_len = k.len + delta;
        k.body_bytes = k.out_len + k.out_len / k.bpl;
        k.sep = (k.out_len % k.bpl != 0 && c != n_contigs - 1) ? 1 : 0;
        k.hdr_off = file;
        k.body_off = file + 1 + k.hdr_len + 1;
        file = k.body_off + k.body_bytes + k.sep;
        k.piece_lo = piece_total;
        if (k.body_bytes > 0) piece_total += (k.body_off + k.body_bytes - 1) / tile_bytes - k.body_off / tile_bytes + 1;
    }
    if (file > fasta_cap) return -1;
    *fasta_len = file;
    SpliceView v{genome, lit, recs, nullptr, tab.conv, tab.comp, Seed{0, 0}};
    memset(fasta, 0, (size_t)file);
    int64_t nf = 0, ns = 0;
    for (int c = 0; c < n_contigs; ++c) {
        const Contig& k = C[c];
        // header kernel
        fasta[k.hdr_off] = '>';
        memcpy(fasta + k.hdr_off + 1, headers + k.hdr_src, (size_t)k.hdr_len);
        fasta[k.hdr_off + 1 + k.hdr_len] = '\n';
        if (k.sep) fasta[k.body_off + k.body_bytes] = '\n';
        if (k.body_bytes == 0) continue;
        // pieces = tiles of the file image intersected with the contig body
        int64_t t0 = k.body_off / tile_bytes, t1 = (k.body_off + k.body_bytes - 1) / tile_bytes;
        for (int64_t t = t0; t <= t1; ++t) {
            int64_t f_lo = t * tile_bytes, f_hi = f_lo + tile_bytes;
            if (f_lo < k.body_off) f_lo = k.body_off;
            if (f_hi > k.body_off + k.body_bytes) f_hi = k.body_off + k.body_bytes;
            for (int64_t g = f_lo & ~(int64_t)15; g < f_hi; g += 16) {
                int64_t a = g < f_lo ? f_lo : g, b = g + 16 > f_hi ? f_hi : g + 16;
                uint32_t w[4] = {0, 0, 0, 0};
                bool fast = false;
                if (b - a == 16) {
                    fast = group_fast(v, k, (uint32_t)(a - k.body_off), w, [&](int64_t idx, uint32_t win[8]) {
                        memcpy(win, genome + idx, 32);
                    });
                }
                if (fast) ++nf; else { ++ns; w[0] = w[1] = w[2] = w[3] = 0; group_slow(v, k, (uint32_t)(a - k.body_off), (int)(b - a), (int)(a - g), w); }
                for (int64_t x = a; x < b; ++x) fasta[x] = (uint8_t)(w[(x - g) >> 2] >> (8 * ((x - g) & 3)));
            }
        }
    }
    *n_fast_groups = nf; *n_slow_groups = ns;
    // VCF (K7): sizes -> offsets -> bytes
    VcfView vv{genome, lit, names, tab.conv, tab.comp, Seed{0, 0}};
    int64_t off = 0;
    std::vector<int64_t> offs(n_recs + 1);
    for (int64_t i = 0; i < n_recs; ++i) { offs[i] = off; off += vcf_line_size(vv, C[recs[i].contig], recs[i]); }
    offs[n_recs] = off;
    if (off > vcf_cap) return -2;
    for (int64_t i = 0; i < n_recs; ++i) {
        if (offs[i + 1] == offs[i]) continue;
        WriteSink s{vcf + offs[i]};
        vcf_emit(s, vv, C[recs[i].contig], recs[i]);
        if (s.p != vcf + offs[i + 1]) return -3;
    }
    *vcf_len = off;
    return 0;
}

// The second-generation splice kernel (ms_tile_core.h) emulated one "thread" at a time: per piece the same phases the
// CTA runs — prep, chunk copy with a dirty queue, dirty bytes, SNP scatter, store — with shared memory on the heap.
// Pieces flagged PD_FALLBACK take the generic per-group path, as on the device.
int emu_apply_tiles(const uint8_t* genome, int32_t n_contigs, const int64_t* goff, const int64_t* clen, const int32_t* bpl,
                    const uint8_t* headers, const int64_t* hdr_off, Rec* recs, int64_t n_recs, const uint8_t* lit,
                    uint8_t* fasta, int64_t fasta_cap, int64_t* fasta_len, int64_t tile_bytes, int64_t* n_clean, int64_t* n_dirty,
                    int64_t* n_fallback) {
    Tables tab; fill_tables(tab);
    std::vector<Contig> C(n_contigs);
    std::vector<SvRec> sv;
    std::vector<Snp8> snp;
    std::vector<int64_t> sv_c(n_contigs + 1), snp_c(n_contigs + 1);
    int64_t ri = 0, file = 0;
    for (int c = 0; c < n_contigs; ++c) {
        Contig& k = C[c];
        memset(&k, 0, sizeof(k));
        k.goff = goff[c]; k.len = clen[c]; k.bpl = bpl[c] > 0 ? bpl[c] : 60; k.gid = c;
        k.hdr_src = hdr_off[c]; k.hdr_len = (int32_t)(hdr_off[c + 1] - hdr_off[c]);
        k.rec_lo = ri;
        sv_c[c] = (int64_t)sv.size(); snp_c[c] = (int64_t)snp.size();
        int64_t delta = 0;
        while (ri < n_recs && recs[ri].contig == (uint32_t)c) {
            Rec& r = recs[ri];
            r.out = (uint32_t)((int64_t)r.pos + delta);
            delta += (int64_t)r.prod - (int64_t)r.cons;
            if (r.kind == K_SNP) snp.push_back(Snp8{r.out, r.alt});
            else sv.push_back(SvRec{r.out, r.prod, r.pos + r.cons, r.pos, r.src, r.kind, 0u});
            ++ri;
        }
        k.rec_hi = ri;
        k.out_len = k.len + delta;
        k.body_bytes = k.out_len + k.out_len / k.bpl;
        k.sep = (k.out_len % k.bpl != 0 && c != n_contigs - 1) ? 1 : 0;
        k.hdr_off = file;
        k.body_off = file + 1 + k.hdr_len + 1;
        file = k.body_off + k.body_bytes + k.sep;
    }
    sv_c[n_contigs] = (int64_t)sv.size(); snp_c[n_contigs] = (int64_t)snp.size();
    sv.resize(sv.size() + 4); snp.resize(snp.size() + 4);
    if (file > fasta_cap) return -1;
    *fasta_len = file;
    memset(fasta, 0, (size_t)file);
    SpliceView v{genome, lit, recs, nullptr, tab.conv, tab.comp, Seed{0, 0}};
    TileView tv{genome, lit, tab.conv, tab.comp, Seed{0, 0}};
    int64_t nc = 0, nd = 0, nf = 0;
    std::vector<uint8_t> stage(TL_STAGE_CAP + 64), image(tile_bytes + 64);
    std::vector<SvRec> ssv(TL_SV_CAP + 1);
    std::vector<uint32_t> rs(TL_SV_CAP + 2);
    std::vector<TileRec> dv(TL_SV_CAP + 1);
    for (int c = 0; c < n_contigs; ++c) {
        const Contig& k = C[c];
        fasta[k.hdr_off] = '>';
        memcpy(fasta + k.hdr_off + 1, headers + k.hdr_src, (size_t)k.hdr_len);
        fasta[k.hdr_off + 1 + k.hdr_len] = '\n';
        if (k.sep) fasta[k.body_off + k.body_bytes] = '\n';
        if (k.body_bytes == 0) continue;
        int64_t t0 = k.body_off / tile_bytes, t1 = (k.body_off + k.body_bytes - 1) / tile_bytes;
        for (int64_t t = t0; t <= t1; ++t) {
            int64_t f_lo = t * tile_bytes, f_hi = f_lo + tile_bytes;
            if (f_lo < k.body_off) f_lo = k.body_off;
            if (f_hi > k.body_off + k.body_bytes) f_hi = k.body_off + k.body_bytes;
            const PieceDesc d = tile_describe(k, (uint32_t)c, f_lo, f_hi, sv.data(), sv_c[c], sv_c[c + 1], snp.data(), snp_c[c], snp_c[c + 1]);
            const int64_t g0 = f_lo & ~(int64_t)15;
            if (d.flags & PD_FALLBACK) {
                ++nf;
                for (int64_t g = g0; g < f_hi; g += 16) {
                    int64_t a = g < f_lo ? f_lo : g, b = g + 16 > f_hi ? f_hi : g + 16;
                    uint32_t w[4] = {0, 0, 0, 0};
                    group_slow(v, k, (uint32_t)(a - k.body_off), (int)(b - a), (int)(a - g), w);
                    for (int64_t x = a; x < b; ++x) fasta[x] = (uint8_t)(w[(x - g) >> 2] >> (8 * ((x - g) & 3)));
                }
                continue;
            }
            // "TMA": staged span, records
            memset(stage.data(), 0xEE, stage.size());
            memcpy(stage.data(), genome + d.in_lo, d.in_bytes);
            memset(image.data(), 0xDD, image.size());
            if (d.flags & PD_GOV_VIRTUAL) { ssv[0] = SvRec{0u, 0u, 0u, 0u, 0, K_NONE, 0u}; for (uint32_t j = 1; j < d.n_sv; ++j) ssv[j] = sv[d.sv_lo + j - 1]; }
            else for (uint32_t j = 0; j < d.n_sv; ++j) ssv[j] = sv[d.sv_lo + j];
            TileShared sh{stage.data(), image.data(), ssv.data(), snp.data() + d.snp_lo, rs.data(), dv.data()};
            for (uint32_t j = 0; j < d.n_sv; ++j) tile_prep_rec(d, sh, j);
            rs[d.n_sv] = d.b_hi - d.b_lo;
            const float rcp_bpl = d.rcp_bpl;
            const TileGeom geo = tile_geom(d);
            const uint32_t e = geo.e, img_end = geo.img_end;
            struct Job { uint32_t x, j, r, n; };
            struct HostOps {
                uint8_t* image; const uint8_t* stage; const uint8_t* genome; std::vector<Job>* jobs; int64_t* nc;
                void copy_stage(uint32_t x, uint32_t s, uint32_t n) { memcpy(image + x, stage + s, n); ++*nc; }
                void copy_global(uint32_t x, int64_t g, uint32_t n) { memcpy(image + x, genome + g, n); ++*nc; }
                void job(uint32_t x, uint32_t j, uint32_t r, uint32_t n) { jobs->push_back(Job{x, j, r, n}); }
                void put(uint32_t x, uint8_t c) { image[x] = c; }
            };
            std::vector<Job> jobs;
            HostOps ops{image.data(), stage.data(), genome, &jobs, &nc};
            for (uint32_t cell = 0; cell < geo.n_cells; ++cell) tile_cell(d, sh, geo, ops, cell);
            for (const Job& q : jobs) {
                ++nd;
                for (uint32_t t2 = 0; t2 < q.n; ++t2) image[q.x + t2] = tile_payload_byte(d, sh, tv, q.j, q.r + t2);
            }
            for (uint32_t i = 0; i < d.n_snp; ++i) {
                const Snp8 sp = sh.snp[i];
                image[e + tile_snp_offset(d, sp.out, rcp_bpl)] = (uint8_t)sp.alt;
            }
            for (uint32_t X = e; X < img_end; ++X) fasta[g0 + X] = image[X];
        }
    }
    *n_clean = nc; *n_dirty = nd; *n_fallback = nf;
    return 0;
}

// Sampling per-candidate logic (K2) and the chain walk (K3) on the CPU.
int emu_type_len(uint64_t seed, uint32_t gid, uint32_t pos, const double cdf[7], const int32_t minlen[7],
                 const int32_t maxlen[7], int64_t limit, uint8_t* type, uint32_t* len) {
    RangeParams rp;
    for (int t = 0; t < 7; ++t) { rp.cdf[t] = cdf[t]; rp.minlen[t] = minlen[t]; rp.maxlen[t] = maxlen[t]; }
    rp.limit = limit;
    draw_type_len(make_seed(seed), gid, pos, rp, *type, *len);
    return 0;
}

// bases of the random insert at (gid, pos): what K_RAND payloads expand to
void emu_rand_insert(uint64_t seed, uint32_t gid, uint32_t pos, uint32_t n, uint8_t* out) {
    for (uint32_t j = 0; j < n; ++j) out[j] = rand_insert_base(make_seed(seed), gid, pos, j);
}

// the same for many inserts at once: insert i = n[i] bases at out + off[i] (full-size parity tests)
void emu_rand_inserts(uint64_t seed, uint32_t gid, int64_t count, const uint32_t* pos, const uint32_t* n, const int64_t* off, uint8_t* out) {
    const Seed s = make_seed(seed);
    for (int64_t i = 0; i < count; ++i)
        for (uint32_t j = 0; j < n[i]; ++j) out[off[i] + j] = rand_insert_base(s, gid, pos[i], j);
}

int emu_snp(uint64_t seed, uint32_t gid, uint32_t pos, uint8_t ref, double p_ti, uint8_t* alt) {
    Tables tab; fill_tables(tab);
    *alt = draw_snp(make_seed(seed), gid, pos, ref, p_ti, tab.trans);
    return 0;
}

}  // extern "C"
