// ms_scan.cuh — generic, order-preserving device-wide scan in ONE pass (chained scan with decoupled look-back).
//
// in(i) -> T produces the i-th input (any fused transform), out(i, excl, val)
// consumes the exclusive prefix, so compaction / scatter fuses into the same
// pass.  Works for non-commutative associative operators.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ms_common.cuh"

namespace ms {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

struct I64x2 { int64_t a, b; };
struct I64x3 { int64_t a, b, c; };
__host__ __device__ inline I64x2 operator+(const I64x2& x, const I64x2& y) { return I64x2{x.a + y.a, x.b + y.b}; }
__host__ __device__ inline I64x3 operator+(const I64x3& x, const I64x3& y) { return I64x3{x.a + y.a, x.b + y.b, x.c + y.c}; }

__device__ inline int64_t shfl_up_t(int64_t v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ inline uint64_t shfl_up_t(uint64_t v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ inline uint32_t shfl_up_t(uint32_t v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ inline I64x2 shfl_up_t(I64x2 v, int d) { return I64x2{shfl_up_t(v.a, d), shfl_up_t(v.b, d)}; }
__device__ inline I64x3 shfl_up_t(I64x3 v, int d) { return I64x3{shfl_up_t(v.a, d), shfl_up_t(v.b, d), shfl_up_t(v.c, d)}; }

struct SumOp { template <class T> __device__ T operator()(const T& a, const T& b) const { return a + b; } };
struct MaxOp { __device__ int64_t operator()(int64_t a, int64_t b) const { return a > b ? a : b; } };

// Exclusive scan of one value per thread across the CTA, in thread order.
template <class T, class Op>
__device__ inline T block_excl_scan(T v, T identity, Op op, T& total, T* sm /* 2 * SCAN_THREADS/32 */) {
    constexpr int NW = SCAN_THREADS / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = shfl_up_t(inc, d);
        if (lane >= d) inc = op(o, inc);
    }
    T excl = shfl_up_t(inc, 1);
    if (lane == 0) excl = identity;
    if (lane == 31) sm[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        T w = lane < NW ? sm[lane] : identity;
#pragma unroll
        for (int d = 1; d < NW; d <<= 1) {
            T o = shfl_up_t(w, d);
            if (lane >= d) w = op(o, w);
        }
        if (lane < NW) sm[NW + lane] = w;
    }
    __syncthreads();
    total = sm[NW + NW - 1];
    T r = wid ? op(sm[NW + wid - 1], excl) : excl;
    __syncthreads();  // sm may be reused by the caller's next call
    return r;
}

// ---- single pass: chained scan with decoupled look-back ----------------------------------------------------------
// One launch: tiles are numbered by an atomic ticket (so a tile only ever waits for tiles that already run), each
// publishes its aggregate, then its inclusive prefix, in global memory; a tile's first warp looks back over its
// predecessors 32 at a time.  The input is read once (the reduce / scan-of-totals / down-sweep form above reads it
// twice and has a single-CTA middle kernel).  Works for non-commutative operators: look-back values are combined
// oldest-first.
template <class T> __device__ __forceinline__ T ld_cg_t(const T* p) {
    static_assert(sizeof(T) % 8 == 0, "scan value must be a multiple of 8 bytes");
    T r;
    const unsigned long long* s = reinterpret_cast<const unsigned long long*>(p);
    unsigned long long* d = reinterpret_cast<unsigned long long*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 8); ++i) d[i] = __ldcg(s + i);
    return r;
}
template <class T> __device__ __forceinline__ void st_cg_t(T* p, const T& v) {
    unsigned long long* d = reinterpret_cast<unsigned long long*>(p);
    const unsigned long long* s = reinterpret_cast<const unsigned long long*>(&v);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 8); ++i) __stcg(d + i, s[i]);
}
__device__ inline int64_t shfl_down_t(int64_t v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ inline I64x2 shfl_down_t(I64x2 v, int d) { return I64x2{shfl_down_t(v.a, d), shfl_down_t(v.b, d)}; }
__device__ inline I64x3 shfl_down_t(I64x3 v, int d) { return I64x3{shfl_down_t(v.a, d), shfl_down_t(v.b, d), shfl_down_t(v.c, d)}; }
__device__ inline int64_t shfl_idx_t(int64_t v, int l) { return __shfl_sync(0xffffffffu, v, l); }
__device__ inline I64x2 shfl_idx_t(I64x2 v, int l) { return I64x2{shfl_idx_t(v.a, l), shfl_idx_t(v.b, l)}; }
__device__ inline I64x3 shfl_idx_t(I64x3 v, int l) { return I64x3{shfl_idx_t(v.a, l), shfl_idx_t(v.b, l), shfl_idx_t(v.c, l)}; }

// tile status: 0 = nothing yet, 1 = aggregate published, 2 = inclusive prefix published
template <class T, class Op, class InF, class OutF>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_1p(InF in, OutF out, int64_t n, T identity, Op op, uint32_t* ticket, uint32_t* status, T* agg, T* incl, T* total_out, int64_t nt) {
    __shared__ T sm[2 * SCAN_THREADS / 32];
    __shared__ uint32_t s_tile;
    __shared__ T s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t base = tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T acc = identity;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        const int64_t i = base + j;
        v[j] = i < n ? in(i) : identity;
        acc = op(acc, v[j]);
    }
    T total;
    const T ex = block_excl_scan(acc, identity, op, total, sm);
    if (tile == 0) {
        if (threadIdx.x == 0) {
            st_cg_t(&incl[0], total);
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(&status[0]) = 2u;
            s_prefix = identity;
        }
    } else {
        if (threadIdx.x == 0) {
            st_cg_t(&agg[tile], total);
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(&status[tile]) = 1u;
        }
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            T prefix = identity;
            int64_t p = tile - 1;
            while (true) {
                const int64_t idx = p - lane;
                uint32_t st = 2u;
                T val = identity;
                if (idx >= 0) {
                    do { st = *reinterpret_cast<volatile uint32_t*>(&status[idx]); } while (st == 0u);
                    __threadfence();
                    val = st == 2u ? ld_cg_t(&incl[idx]) : ld_cg_t(&agg[idx]);
                }
                const uint32_t done = __ballot_sync(0xffffffffu, st == 2u);
                const int first = done ? __ffs(done) - 1 : 32;          // nearest predecessor with an inclusive prefix
                if (lane > first) val = identity;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {                      // lane 0 <- val[31] o ... o val[1] o val[0] (oldest first)
                    const T o = shfl_down_t(val, d);
                    if (lane + d < 32) val = op(o, val);
                }
                prefix = op(shfl_idx_t(val, 0), prefix);
                if (done) break;
                p -= 32;
            }
            if (lane == 0) {
                st_cg_t(&incl[tile], op(prefix, total));
                __threadfence();
                *reinterpret_cast<volatile uint32_t*>(&status[tile]) = 2u;
                s_prefix = prefix;
            }
        }
    }
    __syncthreads();
    const T pre = s_prefix;
    T run = op(pre, ex);
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        const int64_t i = base + j;
        if (i < n) out(i, run, v[j]);
        run = op(run, v[j]);
    }
    if (tile == nt - 1 && threadIdx.x == 0) *total_out = op(pre, total);
}

template <class T> __global__ void k_scan_store(T* dst, T v) { *dst = v; }

// Returns a device pointer to the grand total (valid until tmp is reused).
template <class T, class Op, class InF, class OutF>
inline cudaError_t device_scan(ms_ctx* c, InF in, OutF out, int64_t n, T identity, Op op, DevBuf& tmp, T** d_total) {
    const int64_t nt = ceil_div(n, SCAN_TILE);
    const size_t status_bytes = ((size_t)nt * 4 + 15) & ~(size_t)15;
    cudaError_t e = tmp.ensure(16 + status_bytes + (size_t)(2 * nt + 1) * sizeof(T) + 16);
    if (e != cudaSuccess) return e;
    uint8_t* base = tmp.as<uint8_t>();
    uint32_t* ticket = reinterpret_cast<uint32_t*>(base);
    uint32_t* status = reinterpret_cast<uint32_t*>(base + 16);
    T* agg = reinterpret_cast<T*>(base + 16 + status_bytes);
    T* incl = agg + nt;
    T* total = incl + nt;
    if (nt > 0) {
        e = cudaMemsetAsync(base, 0, 16 + status_bytes, c->stream);
        if (e != cudaSuccess) return e;
        k_scan_1p<T, Op, InF, OutF><<<(unsigned)nt, SCAN_THREADS, 0, c->stream>>>(in, out, n, identity, op, ticket, status, agg, incl, total, nt);
    } else {
        k_scan_store<T><<<1, 1, 0, c->stream>>>(total, identity);
    }
    c->kernel_launches++;
    if (d_total) *d_total = total;
    return cudaGetLastError();
}

}  // namespace ms
