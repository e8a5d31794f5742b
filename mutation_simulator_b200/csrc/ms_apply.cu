// ms_apply.cu — K5 (plan: delta scan, layout, block index), K6 (splice + SNP +
// line wrap -> FASTA image), K7 (VCF lines).  Replaces Mutator.__mutate_sequence
// (mutator.py:318-426), FastaWriter (fasta_writer.py:40-65) and VcfWriter.write
// (vcf_writer.py:118-126).
#include "ms_common.cuh"
#include "ms_scan.cuh"
#include "ms_splice_core.h"
#include "ms_vcf_core.h"

namespace ms {

struct Gap { int64_t start; uint32_t count; uint32_t value; };
constexpr uint32_t GAP_INLINE = 32;

__device__ inline void raise_error(Totals* t, int64_t code, int64_t arg) {
    if (atomicCAS((unsigned long long*)&t->error, 0ull, (unsigned long long)code) == 0ull) t->error_arg = arg;
}

// ---- record ranges per contig -------------------------------------------------------
__global__ void k_rec_bounds(const Rec* recs, int64_t n_recs, Contig* contigs, int32_t n_contigs) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_contigs) return;
    int64_t lo = 0, hi = n_recs;  // first record with contig >= c
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (recs[mid].contig < (uint32_t)c) lo = mid + 1; else hi = mid;
    }
    if (c < n_contigs) contigs[c].rec_lo = lo;
    if (c > 0) contigs[c - 1].rec_hi = lo;
}

// ---- layout of the output file image ---------------------------------------------------
// One CTA walks the contigs in FASTA order (they are few: 24 .. 200k) carrying the
// running file offset, block-index offset and piece count.
__global__ void __launch_bounds__(SCAN_THREADS)
k_contig_layout(Contig* contigs, int32_t n_contigs, const int64_t* S, int64_t* piece_lo, int64_t tile_bytes,
                Gap* gaps, int64_t gap_cap, Totals* tot, int64_t n_recs, const int64_t* V) {
    __shared__ I64x2 sm2[2 * SCAN_THREADS / 32];
    __shared__ int64_t sm1[2 * SCAN_THREADS / 32];
    I64x2 carry{0, 0};
    for (int base = 0; base < n_contigs; base += SCAN_THREADS) {
        const int c = base + threadIdx.x;
        I64x2 v{0, 0};
        if (c < n_contigs) {
            Contig& k = contigs[c];
            const int64_t delta = S[k.rec_hi] - S[k.rec_lo];
            const int64_t out_len = k.len + delta;
            if (out_len < 0) raise_error(tot, MS_ERR_OVERLAP, c);
            k.out_len = out_len;
            const int64_t bpl = k.bpl;
            k.body_bytes = out_len + out_len / bpl;
            if (k.body_bytes >= (int64_t)0xFFFFFFF0ll || out_len >= (int64_t)0xFFFFFFF0ll) raise_error(tot, MS_ERR_LIMIT, c);
            k.sep = (out_len % bpl != 0 && c != n_contigs - 1) ? 1u : 0u;
            v.a = (int64_t)k.hdr_len + 2 + k.body_bytes + k.sep;
            v.b = (out_len >> BLK_SHIFT) + 1;
        }
        I64x2 total;
        I64x2 ex = block_excl_scan(v, I64x2{0, 0}, SumOp(), total, sm2);
        if (c < n_contigs) {
            Contig& k = contigs[c];
            k.hdr_off = carry.a + ex.a;
            k.body_off = k.hdr_off + k.hdr_len + 2;
            k.blk_lo = carry.b + ex.b;
        }
        carry = carry + total;
    }
    __syncthreads();
    int64_t pcarry = 0;
    for (int base = 0; base < n_contigs; base += SCAN_THREADS) {
        const int c = base + threadIdx.x;
        int64_t np = 0;
        if (c < n_contigs) {
            const Contig& k = contigs[c];
            if (k.body_bytes > 0) np = (k.body_off + k.body_bytes - 1) / tile_bytes - k.body_off / tile_bytes + 1;
        }
        int64_t total;
        int64_t ex = block_excl_scan(np, (int64_t)0, SumOp(), total, sm1);
        if (c < n_contigs) { contigs[c].piece_lo = pcarry + ex; piece_lo[c] = pcarry + ex; }
        pcarry += total;
    }
    if (threadIdx.x == 0) {
        piece_lo[n_contigs] = pcarry;
        tot->fasta_bytes = carry.a;
        tot->n_blk = carry.b;
        tot->n_pieces = pcarry;
        tot->vcf_bytes = V[n_recs];
        tot->n_recs = n_recs;
    }
}

// ---- out positions, validation and the coarse block index ----------------------------
// blk[k] of a contig = number of its records with out < k*BLK_BASES.
__device__ inline void fill_blk(uint32_t* blk, int64_t start, int64_t count, uint32_t value, Gap* gaps, int64_t gap_cap, Totals* tot) {
    if (count <= 0) return;
    if (count <= GAP_INLINE) {
        for (int64_t k = 0; k < count; ++k) blk[start + k] = value;
    } else {
        while (count > 0) {  // split very long gaps so one CTA never fills more than 1M entries
            const int64_t n = count > (1 << 20) ? (1 << 20) : count;
            const unsigned long long slot = atomicAdd((unsigned long long*)&tot->n_long_gaps, 1ull);
            if ((int64_t)slot < gap_cap) gaps[slot] = Gap{start, (uint32_t)n, value};
            else raise_error(tot, MS_ERR_INTERNAL, 1);
            start += n; count -= n;
        }
    }
}

__global__ void __launch_bounds__(256)
k_rec_out(Rec* recs, int64_t n_recs, const Contig* contigs, const int64_t* S, uint32_t* blk, Gap* gaps, int64_t gap_cap, Totals* tot) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_recs) return;
    Rec r = recs[i];
    const Contig& k = contigs[r.contig];
    const int64_t s0 = S[k.rec_lo];
    const int64_t out = (int64_t)r.pos + (S[i] - s0);
    if ((int64_t)r.pos + r.cons > k.len) raise_error(tot, MS_ERR_OVERLAP, i);
    int64_t prev_blk = -1;  // block of the previous record's out
    if (i > k.rec_lo) {
        const Rec p = recs[i - 1];
        if ((int64_t)r.pos < (int64_t)p.pos + p.cons || r.pos == p.pos) raise_error(tot, MS_ERR_OVERLAP, i);
        prev_blk = ((int64_t)p.pos + (S[i - 1] - s0)) >> BLK_SHIFT;
    }
    recs[i].out = (uint32_t)out;
    const uint32_t j = (uint32_t)(i - k.rec_lo);
    const int64_t my_blk = out >> BLK_SHIFT;
    fill_blk(blk, k.blk_lo + prev_blk + 1, my_blk - prev_blk, j, gaps, gap_cap, tot);
    if (i + 1 == k.rec_hi) {
        const int64_t nblk = (k.out_len >> BLK_SHIFT) + 1;
        fill_blk(blk, k.blk_lo + my_blk + 1, nblk - (my_blk + 1), j + 1, gaps, gap_cap, tot);
    }
}

__global__ void k_empty_contig_gaps(const Contig* contigs, int32_t n_contigs, uint32_t* blk, Gap* gaps, int64_t gap_cap, Totals* tot) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_contigs) return;
    const Contig& k = contigs[c];
    if (k.rec_lo == k.rec_hi) fill_blk(blk, k.blk_lo, (k.out_len >> BLK_SHIFT) + 1, 0u, gaps, gap_cap, tot);
}

__global__ void __launch_bounds__(256) k_fill_gaps(uint32_t* blk, const Gap* gaps, const Totals* tot) {
    const int64_t n = tot->n_long_gaps;
    for (int64_t g = blockIdx.x; g < n; g += gridDim.x) {
        const Gap gp = gaps[g];
        for (uint32_t k = threadIdx.x; k < gp.count; k += blockDim.x) blk[gp.start + k] = gp.value;
    }
}

// ---- K6: splice + SNP + line wrap -----------------------------------------------------
// One CTA per piece (= 16 KiB tile of the output file image intersected with one
// contig body).  Pass A: every thread tries the vector path for its 16-byte groups
// (shifted copy + SNP patches + one line break, 2 x LDG.128 -> 1 x STG.128) and
// queues the rest; pass B: the queued groups are drained densely through the
// generic per-byte path.
constexpr int SPLICE_THREADS = 256;
constexpr int MAX_TILE_GROUPS = 2048;

struct LoadWinGlobal {
    const uint8_t* genome;
    __device__ __forceinline__ void operator()(int64_t idx, uint32_t win[8]) const {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(genome + idx));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(genome + idx + 16));
        win[0] = a.x; win[1] = a.y; win[2] = a.z; win[3] = a.w;
        win[4] = b.x; win[5] = b.y; win[6] = b.z; win[7] = b.w;
    }
};

__global__ void __launch_bounds__(SPLICE_THREADS)
k_splice(SpliceView v, const Contig* contigs, int32_t n_contigs, const int64_t* piece_lo, const Tables* tables,
         uint8_t* fasta, int64_t tile_bytes) {
    __shared__ Contig sc;
    __shared__ uint16_t dirty[MAX_TILE_GROUPS];
    __shared__ int n_dirty;
    __shared__ uint8_t s_conv[256], s_comp[256];
    const int tid = threadIdx.x;
    const int64_t p = blockIdx.x;
    if (tid == 0) {
        int lo = 0, hi = n_contigs;  // last c with piece_lo[c] <= p
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (piece_lo[mid] <= p) lo = mid; else hi = mid;
        }
        sc = contigs[lo];
        n_dirty = 0;
    }
    s_conv[tid] = tables->conv[tid];
    s_comp[tid] = tables->comp[tid];
    __syncthreads();
    const Contig& k = sc;
    v.conv = s_conv;
    v.comp = s_comp;
    const int64_t tile = k.body_off / tile_bytes + (p - k.piece_lo);
    int64_t f_lo = tile * tile_bytes, f_hi = f_lo + tile_bytes;
    if (f_lo < k.body_off) f_lo = k.body_off;
    if (f_hi > k.body_off + k.body_bytes) f_hi = k.body_off + k.body_bytes;
    const int64_t g0 = f_lo & ~(int64_t)15;
    const int ngroups = (int)((f_hi - g0 + 15) >> 4);
    const LoadWinGlobal loader{v.genome};

    for (int gi = tid; gi < ngroups; gi += SPLICE_THREADS) {
        const int64_t g = g0 + ((int64_t)gi << 4);
        bool done = false;
        if (g >= f_lo && g + 16 <= f_hi) {
            uint32_t w[4];
            if (group_fast(v, k, (uint32_t)(g - k.body_off), w, loader)) {
                *reinterpret_cast<uint4*>(fasta + g) = make_uint4(w[0], w[1], w[2], w[3]);
                done = true;
            }
        }
        if (!done) dirty[atomicAdd(&n_dirty, 1)] = (uint16_t)gi;
    }
    __syncthreads();
    const int nd = n_dirty;
    for (int e = tid; e < nd; e += SPLICE_THREADS) {
        const int64_t g = g0 + ((int64_t)dirty[e] << 4);
        const int64_t a = g < f_lo ? f_lo : g;
        const int64_t b = g + 16 > f_hi ? f_hi : g + 16;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        group_slow(v, k, (uint32_t)(a - k.body_off), (int)(b - a), (int)(a - g), w);
        if (b - a == 16) {
            *reinterpret_cast<uint4*>(fasta + g) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (int64_t x = a; x < b; ++x) {
                const int lane = (int)(x - g);
                fasta[x] = (uint8_t)(w[lane >> 2] >> (8 * (lane & 3)));
            }
        }
    }
}

// ">header\n" of every contig and the "\n" that closes a partial last line
// (fasta_writer.py:40-47).
__global__ void k_headers(const Contig* contigs, int32_t n_contigs, const uint8_t* headers, uint8_t* fasta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_contigs) return;
    const Contig& k = contigs[c];
    uint8_t* d = fasta + k.hdr_off;
    d[0] = '>';
    for (int i = 0; i < k.hdr_len; ++i) d[1 + i] = headers[k.hdr_src + i];
    d[1 + k.hdr_len] = '\n';
    if (k.sep) fasta[k.body_off + k.body_bytes] = '\n';
}

// ---- K7: VCF lines -----------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_vcf_write(VcfView v, const Rec* recs, int64_t n_recs, const Contig* contigs, const Tables* tables, const int64_t* V, uint8_t* vcf) {
    __shared__ uint8_t s_conv[256], s_comp[256];
    s_conv[threadIdx.x] = tables->conv[threadIdx.x];
    s_comp[threadIdx.x] = tables->comp[threadIdx.x];
    __syncthreads();
    v.conv = s_conv; v.comp = s_comp;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_recs) return;
    const int64_t a = V[i], b = V[i + 1];
    if (b == a) return;
    const Rec r = recs[i];
    WriteSink s{vcf + a};
    vcf_emit(s, v, contigs[r.contig], r);
}

__global__ void k_store_total2(const I64x2* total, int64_t* S_end, int64_t* V_end) {
    *S_end = total->a;
    *V_end = total->b;
}

// ---- host orchestration --------------------------------------------------------------
int apply_pipeline(ms_ctx* c) {
    if (c->n_contigs <= 0 || !c->genome.p) MS_FAIL(c, MS_ERR_STATE, "ms_apply: no genome resident");
    const int64_t M = c->n_recs;
    Contig* d_contigs = c->contigs.as<Contig>();
    Rec* d_recs = c->recs.as<Rec>();
    const Tables* d_tab = c->tables.as<Tables>();
    Totals* d_tot = c->totals.as<Totals>();
    cudaStream_t st = c->stream;

    stage_begin(c, ST_PLAN);
    MS_CUDA(c, cudaMemsetAsync(d_tot, 0, sizeof(Totals), st));
    MS_CUDA(c, c->svec.ensure((size_t)(M + 1) * sizeof(int64_t)));
    MS_CUDA(c, c->vcf_off.ensure((size_t)(M + 1) * sizeof(int64_t)));
    MS_CUDA(c, c->piece_lo.ensure((size_t)(c->n_contigs + 1) * sizeof(int64_t)));
    MS_CUDA(c, c->recs.ensure(32));  // M == 0: keep pointers valid
    d_recs = c->recs.as<Rec>();
    int64_t* S = c->svec.as<int64_t>();
    int64_t* V = c->vcf_off.as<int64_t>();

    k_rec_bounds<<<(unsigned)ceil_div(c->n_contigs + 1, 128), 128, 0, st>>>(d_recs, M, d_contigs, c->n_contigs);
    MS_LAUNCH_CHECK(c);

    VcfView vv{c->genome.as<uint8_t>(), c->lit.as<uint8_t>(), c->names.as<uint8_t>(), d_tab->conv, d_tab->comp};
    {
        const Rec* recs = d_recs;
        const Contig* contigs = d_contigs;
        auto in = [=] __device__(int64_t i) -> I64x2 {
            const Rec r = recs[i];
            return I64x2{(int64_t)r.prod - (int64_t)r.cons, (int64_t)vcf_line_size(vv, contigs[r.contig], r)};
        };
        auto out = [=] __device__(int64_t i, I64x2 ex, I64x2) { S[i] = ex.a; V[i] = ex.b; };
        I64x2* d_total = nullptr;
        MS_CUDA(c, (device_scan<I64x2>(c, in, out, M, I64x2{0, 0}, SumOp(), c->scan_tmp, &d_total)));
        k_store_total2<<<1, 1, 0, st>>>(d_total, S + M, V + M);
        MS_LAUNCH_CHECK(c);
    }
    // gap list capacity is bounded by the block count; size it from the input side (exact bound needs out_len)
    const int64_t gap_cap = c->total_bases / (BLK_BASES * GAP_INLINE) + 4 * (int64_t)c->n_contigs + M / GAP_INLINE + 1024;
    MS_CUDA(c, c->long_gaps.ensure((size_t)gap_cap * sizeof(Gap)));
    k_contig_layout<<<1, SCAN_THREADS, 0, st>>>(d_contigs, c->n_contigs, S, c->piece_lo.as<int64_t>(), (int64_t)c->tile_bytes,
                                                c->long_gaps.as<Gap>(), gap_cap, d_tot, M, V);
    MS_LAUNCH_CHECK(c);
    MS_CUDA(c, cudaMemcpyAsync(c->h_totals, d_tot, sizeof(Totals), cudaMemcpyDeviceToHost, st));
    stage_end(c, ST_PLAN);
    MS_CUDA(c, cudaStreamSynchronize(st));
    Totals t = *c->h_totals;
    if (t.error) MS_FAIL(c, (int)t.error, "ms_apply: layout failed (code %lld at contig %lld)", (long long)t.error, (long long)t.error_arg);
    c->fasta_bytes = t.fasta_bytes; c->vcf_bytes = t.vcf_bytes; c->n_blk = t.n_blk; c->n_pieces = t.n_pieces;

    MS_CUDA(c, c->blk.ensure((size_t)(t.n_blk + 1) * sizeof(uint32_t)));
    MS_CUDA(c, c->fasta.ensure((size_t)t.fasta_bytes + 64));
    MS_CUDA(c, c->vcf.ensure((size_t)t.vcf_bytes + 64));
    uint32_t* d_blk = c->blk.as<uint32_t>();

    stage_begin(c, ST_INDEX);
    if (M > 0) {
        k_rec_out<<<(unsigned)ceil_div(M, 256), 256, 0, st>>>(d_recs, M, d_contigs, S, d_blk, c->long_gaps.as<Gap>(), gap_cap, d_tot);
        MS_LAUNCH_CHECK(c);
    }
    k_empty_contig_gaps<<<(unsigned)ceil_div(c->n_contigs, 128), 128, 0, st>>>(d_contigs, c->n_contigs, d_blk, c->long_gaps.as<Gap>(), gap_cap, d_tot);
    MS_LAUNCH_CHECK(c);
    k_fill_gaps<<<NUM_SMS_B200 * 4, 256, 0, st>>>(d_blk, c->long_gaps.as<Gap>(), d_tot);
    MS_LAUNCH_CHECK(c);
    stage_end(c, ST_INDEX);

    stage_begin(c, ST_SPLICE);
    SpliceView sv{c->genome.as<uint8_t>(), c->lit.as<uint8_t>(), d_recs, d_blk, d_tab->conv, d_tab->comp};
    if (t.n_pieces > 0) {
        k_splice<<<(unsigned)t.n_pieces, SPLICE_THREADS, 0, st>>>(sv, d_contigs, c->n_contigs, c->piece_lo.as<int64_t>(), d_tab,
                                                                 c->fasta.as<uint8_t>(), (int64_t)c->tile_bytes);
        MS_LAUNCH_CHECK(c);
    }
    k_headers<<<(unsigned)ceil_div(c->n_contigs, 128), 128, 0, st>>>(d_contigs, c->n_contigs, c->headers.as<uint8_t>(), c->fasta.as<uint8_t>());
    MS_LAUNCH_CHECK(c);
    stage_end(c, ST_SPLICE);

    stage_begin(c, ST_VCF);
    if (M > 0) {
        k_vcf_write<<<(unsigned)ceil_div(M, 256), 256, 0, st>>>(vv, d_recs, M, d_contigs, d_tab, V, c->vcf.as<uint8_t>());
        MS_LAUNCH_CHECK(c);
    }
    MS_CUDA(c, cudaMemcpyAsync(c->h_totals, d_tot, sizeof(Totals), cudaMemcpyDeviceToHost, st));
    stage_end(c, ST_VCF);
    MS_CUDA(c, cudaStreamSynchronize(st));
    t = *c->h_totals;
    if (t.error) MS_FAIL(c, (int)t.error, "ms_apply: records invalid (code %lld at record %lld): overlapping or out of bounds",
                         (long long)t.error, (long long)t.error_arg);
    c->last_totals.fasta_bytes = t.fasta_bytes;
    c->last_totals.vcf_bytes = t.vcf_bytes;
    c->last_totals.n_recs = M;
    return MS_OK;
}

}  // namespace ms
