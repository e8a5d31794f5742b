// ms_genome.cu — synthetic genome generation on the device (benchmarks only;
// SURVEY.md §8d "Synthetic inputs": iid uniform ACGT, one centromere-like N run
// per contig and N telomeres).
#include "ms_common.cuh"

namespace ms {

// 64 bases per thread from one Philox block (2 bits per base).
__global__ void __launch_bounds__(256) k_synth(uint8_t* g, int64_t total, Seed seed) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t base = q * 64;
    if (base >= total) return;
    const U4 r = draw(seed, 0u, P_GENOME, (uint64_t)q);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint32_t out[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = i * 4 + b;
            const uint32_t two = (w[j >> 4] >> ((j & 15) * 2)) & 3u;
            v |= (uint32_t)("ACGT"[two]) << (8 * b);
        }
        out[i] = v;
    }
    uint4* dst = reinterpret_cast<uint4*>(g + base);  // buffer is padded to a multiple of 64
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_uint4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
}

}  // namespace ms

using namespace ms;

extern "C" int ms_genome_synth(ms_ctx* c, uint64_t seed, int32_t n_contigs, const int64_t* contig_len, const int32_t* bpl,
                               double n_fraction, int64_t telomere_n, const uint8_t* headers, const int64_t* hdr_off,
                               const uint8_t* names, const int64_t* name_off) {
    if (!c || n_contigs <= 0 || !contig_len) return MS_ERR_ARG;
    MS_CUDA(c, cudaSetDevice(c->device));
    int64_t total = 0;
    for (int i = 0; i < n_contigs; ++i) total += contig_len[i];
    MS_CUDA(c, c->genome.ensure((size_t)total + 192));
    const int64_t chunks = ceil_div(total, 64);
    if (chunks > 0) {
        k_synth<<<(unsigned)ceil_div(chunks, 256), 256, 0, c->stream>>>(c->genome.as<uint8_t>(), total, make_seed(seed));
        MS_LAUNCH_CHECK(c);
    }
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
    return ms_genome_adopt(c, c->genome.as<uint8_t>(), total, n_contigs, contig_len, bpl, nullptr, headers, hdr_off, names, name_off);
}
