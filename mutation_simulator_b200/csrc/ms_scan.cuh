// ms_scan.cuh — generic, order-preserving device-wide scan (reduce-then-scan).
//
// in(i) -> T produces the i-th input (any fused transform), out(i, excl, val)
// consumes the exclusive prefix, so compaction / scatter fuses into the
// down-sweep.  Three launches: tile reduce, single-CTA scan of tile totals,
// down-sweep.  Works for non-commutative associative operators.
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

template <class T, class Op, class InF>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(InF in, int64_t n, T identity, Op op, T* tile_sums) {
    __shared__ T sm[2 * SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    T acc = identity;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        const int64_t i = base + j;
        if (i < n) acc = op(acc, in(i));
    }
    T total;
    block_excl_scan(acc, identity, op, total, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// In place: tile_sums[0..nt) -> exclusive prefixes, tile_sums[nt] = grand total.
template <class T, class Op>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(T* tile_sums, int64_t nt, T identity, Op op) {
    __shared__ T sm[2 * SCAN_THREADS / 32];
    T carry = identity;
    for (int64_t base = 0; base < nt; base += SCAN_THREADS) {
        const int64_t i = base + threadIdx.x;
        T v = i < nt ? tile_sums[i] : identity;
        T total;
        T ex = block_excl_scan(v, identity, op, total, sm);
        if (i < nt) tile_sums[i] = op(carry, ex);
        carry = op(carry, total);
    }
    if (threadIdx.x == 0) tile_sums[nt] = carry;
}

template <class T, class Op, class InF, class OutF>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_down(InF in, OutF out, int64_t n, T identity, Op op, const T* tile_prefix) {
    __shared__ T sm[2 * SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T acc = identity;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        const int64_t i = base + j;
        v[j] = i < n ? in(i) : identity;
        acc = op(acc, v[j]);
    }
    T total;
    T ex = block_excl_scan(acc, identity, op, total, sm);
    T run = op(tile_prefix[blockIdx.x], ex);
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        const int64_t i = base + j;
        if (i < n) out(i, run, v[j]);
        run = op(run, v[j]);
    }
}

// Returns a device pointer to the grand total (valid until tmp is reused).
template <class T, class Op, class InF, class OutF>
inline cudaError_t device_scan(ms_ctx* c, InF in, OutF out, int64_t n, T identity, Op op, DevBuf& tmp, T** d_total) {
    const int64_t nt = ceil_div(n, SCAN_TILE);
    cudaError_t e = tmp.ensure((size_t)(nt + 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    T* ts = tmp.as<T>();
    if (nt > 0) { k_scan_reduce<T, Op, InF><<<(unsigned)nt, SCAN_THREADS, 0, c->stream>>>(in, n, identity, op, ts); c->kernel_launches++; }
    k_scan_tiles<T, Op><<<1, SCAN_THREADS, 0, c->stream>>>(ts, nt, identity, op); c->kernel_launches++;
    if (nt > 0) { k_scan_down<T, Op, InF, OutF><<<(unsigned)nt, SCAN_THREADS, 0, c->stream>>>(in, out, n, identity, op, ts); c->kernel_launches++; }
    if (d_total) *d_total = ts + nt;
    return cudaGetLastError();
}

}  // namespace ms
