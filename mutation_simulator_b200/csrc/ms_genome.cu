// ms_genome.cu — synthetic genome generation on the device (benchmarks only;
// SURVEY.md §8d "Synthetic inputs": iid uniform ACGT, one centromere-like N run
// per contig and N telomeres).
#include <vector>
#include "ms_common.cuh"

namespace ms {

// 64 bases per thread from one Philox block (2 bits per base).  Base j of the contig with global index gid is a pure
// function of (seed, gid, j): the genome a rank synthesises for its share of the contigs is the same genome a single
// GPU synthesises for all of them (bench.py's partition-invariance check relies on it).
struct SynthCtg { int64_t goff, len, chunk_lo; uint32_t gid, pad; };

__global__ void __launch_bounds__(256) k_synth(uint8_t* g, const SynthCtg* ctg, int32_t n_contigs, int64_t n_chunks, Seed seed) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_chunks) return;
    int lo = 0, hi = n_contigs;          // last contig with chunk_lo <= q
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (ctg[mid].chunk_lo <= q) lo = mid; else hi = mid; }
    const SynthCtg k = ctg[lo];
    const int64_t j0 = (q - k.chunk_lo) * 64;
    const U4 r = draw(seed, k.gid, P_GENOME, (uint64_t)(q - k.chunk_lo));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint8_t* dst = g + k.goff + j0;
    const int n = (int)(k.len - j0 < 64 ? k.len - j0 : 64);
    if (n == 64 && ((k.goff + j0) & 15) == 0) {
        uint32_t out[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int j = i * 4 + b;
                v |= (uint32_t)("ACGT"[(w[j >> 4] >> ((j & 15) * 2)) & 3u]) << (8 * b);
            }
            out[i] = v;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(dst)[i] = make_uint4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
    } else {
        for (int j = 0; j < n; ++j) dst[j] = (uint8_t)("ACGT"[(w[j >> 4] >> ((j & 15) * 2)) & 3u]);
    }
}

// ---- content hash of byte ranges (bench.py: N-GPU output == 1-GPU output without leaving the devices) -------------
__device__ __forceinline__ uint64_t mix64(uint64_t x) {      // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
    return x;
}

// One thread per 8 buffer bytes; each byte adds mix(offset-in-range << 8 | byte) to its range's sum.
__global__ void __launch_bounds__(256)
k_hash_ranges(const uint8_t* buf, int64_t lo, int64_t hi, const int64_t* start, const int64_t* end, int32_t n, unsigned long long* out) {
    const int64_t b0 = lo + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    uint64_t acc = 0;
    int r = -1;
    if (b0 < hi) {
        int a = 0, z = n;                 // last range with start <= b0 (or 0)
        while (z - a > 1) { const int mid = (a + z) >> 1; if (start[mid] <= b0) a = mid; else z = mid; }
        r = a;
        for (int t = 0; t < 8 && b0 + t < hi; ++t) {
            const int64_t x = b0 + t;
            while (r + 1 < n && start[r + 1] <= x) {      // crossed into the next range: flush
                if (acc) atomicAdd(&out[r], (unsigned long long)acc);
                acc = 0; ++r;
            }
            if (x >= start[r] && x < end[r]) acc += mix64(((uint64_t)(x - start[r]) << 8) | buf[x]);
        }
    }
    // most warps lie inside one range: one atomic per warp
    const int r0 = __shfl_sync(0xffffffffu, r, 0);
    if (__all_sync(0xffffffffu, r == r0)) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
        if ((threadIdx.x & 31) == 0 && r0 >= 0 && acc) atomicAdd(&out[r0], (unsigned long long)acc);
    } else if (r >= 0 && acc) {
        atomicAdd(&out[r], (unsigned long long)acc);
    }
}

int hash_ranges(ms_ctx* c, const uint8_t* buf, int32_t n, const int64_t* start, const int64_t* end, uint64_t* out) {
    MS_CUDA(c, c->scan_tmp2.ensure((size_t)n * 24 + 64));
    int64_t* d_start = c->scan_tmp2.as<int64_t>();
    int64_t* d_end = d_start + n;
    unsigned long long* d_out = reinterpret_cast<unsigned long long*>(d_end + n);
    cudaStream_t st = c->stream;
    MS_CUDA(c, cudaMemcpyAsync(d_start, start, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    MS_CUDA(c, cudaMemcpyAsync(d_end, end, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    MS_CUDA(c, cudaMemsetAsync(d_out, 0, (size_t)n * 8, st));
    const int64_t lo = start[0], hi = end[n - 1];
    if (hi > lo) {
        k_hash_ranges<<<(unsigned)ceil_div(ceil_div(hi - lo, 8), 256), 256, 0, st>>>(buf, lo, hi, d_start, d_end, n, d_out);
        MS_LAUNCH_CHECK(c);
    }
    MS_CUDA(c, cudaMemcpyAsync(out, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    MS_CUDA(c, cudaStreamSynchronize(st));
    return MS_OK;
}

}  // namespace ms

using namespace ms;

extern "C" int ms_genome_synth(ms_ctx* c, uint64_t seed, int32_t n_contigs, const int64_t* contig_len, const int32_t* bpl,
                               const uint32_t* gid, double n_fraction, int64_t telomere_n, const uint8_t* headers,
                               const int64_t* hdr_off, const uint8_t* names, const int64_t* name_off) {
    if (!c || n_contigs <= 0 || !contig_len) return MS_ERR_ARG;
    MS_CUDA(c, cudaSetDevice(c->device));
    int64_t total = 0;
    for (int i = 0; i < n_contigs; ++i) total += contig_len[i];
    MS_CUDA(c, c->genome.ensure((size_t)total + 192 + (size_t)c->foreign_cap + 64));
    std::vector<SynthCtg> tab((size_t)n_contigs);
    int64_t chunks = 0, goff = 0;
    for (int i = 0; i < n_contigs; ++i) {
        tab[i] = SynthCtg{goff, contig_len[i], chunks, gid ? gid[i] : (uint32_t)i, 0u};
        chunks += ceil_div(contig_len[i], 64);
        goff += contig_len[i];
    }
    MS_CUDA(c, c->tmp_contigs.ensure(sizeof(SynthCtg) * (size_t)n_contigs));
    MS_CUDA(c, cudaMemcpyAsync(c->tmp_contigs.p, tab.data(), sizeof(SynthCtg) * (size_t)n_contigs, cudaMemcpyHostToDevice, c->stream));
    if (chunks > 0) {
        k_synth<<<(unsigned)ceil_div(chunks, 256), 256, 0, c->stream>>>(c->genome.as<uint8_t>(), c->tmp_contigs.as<SynthCtg>(), n_contigs,
                                                                       chunks, make_seed(seed));
        MS_LAUNCH_CHECK(c);
    }
    MS_CUDA(c, cudaStreamSynchronize(c->stream));      // tab is a host temporary
    if (n_contigs <= 4096 && (n_fraction > 0 || telomere_n > 0)) {
        int64_t off = 0;
        for (int i = 0; i < n_contigs; ++i) {
            const int64_t L = contig_len[i];
            const int64_t tel = telomere_n < L / 4 ? telomere_n : L / 4;
            if (tel > 0) {
                MS_CUDA(c, cudaMemsetAsync(c->genome.as<uint8_t>() + off, 'N', (size_t)tel, c->stream));
                MS_CUDA(c, cudaMemsetAsync(c->genome.as<uint8_t>() + off + L - tel, 'N', (size_t)tel, c->stream));
            }
            const int64_t cen = (int64_t)(n_fraction * (double)L);
            if (cen > 0) MS_CUDA(c, cudaMemsetAsync(c->genome.as<uint8_t>() + off + (L * 2) / 5, 'N', (size_t)cen, c->stream));
            off += L;
        }
    }
    return ms_genome_adopt(c, c->genome.as<uint8_t>(), total, n_contigs, contig_len, bpl, gid, headers, hdr_off, names, name_off);
}
