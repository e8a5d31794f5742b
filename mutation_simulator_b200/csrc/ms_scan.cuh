// ms_scan.cuh — generic, order-preserving device-wide scan (reduce-then-scan; a single-pass variant is kept below).
//
// in(i) -> T produces the i-th input (any fused transform), out(i, excl, val)
// consumes the exclusive prefix, so compaction / scatter fuses into the
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

// In place: tile_sums[0..nt) -> exclusive prefixes, tile_sums[nt] = grand total.  One CTA of 1024 threads, each owning a
// contiguous stretch of the tile sums (a 256-thread CTA walking 20 k sums 256 at a time took 80-98 us of an otherwise
// 0.4 ms scan, profiles/r1n_launches.csv).
constexpr int SCAN_MID_THREADS = 1024;
template <class T, class Op>
__global__ void __launch_bounds__(SCAN_MID_THREADS) k_scan_tiles(T* tile_sums, int64_t nt, T identity, Op op) {
    __shared__ T warp_tot[SCAN_MID_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t per = (nt + SCAN_MID_THREADS - 1) / SCAN_MID_THREADS;
    const int64_t lo = (int64_t)tid * per, hi = lo + per < nt ? lo + per : nt;
    T acc = identity;
    for (int64_t i = lo; i < hi; ++i) acc = op(acc, tile_sums[i]);
    T inc = acc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = shfl_up_t(inc, d);
        if (lane >= d) inc = op(o, inc);
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        T w = warp_tot[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            T o = shfl_up_t(w, d);
            if (lane >= d) w = op(o, w);
        }
        warp_tot[lane] = w;       // inclusive over warps
    }
    __syncthreads();
    T ex = shfl_up_t(inc, 1);
    if (lane == 0) ex = identity;
    T run = wid ? op(warp_tot[wid - 1], ex) : ex;
    for (int64_t i = lo; i < hi; ++i) {
        const T v = tile_sums[i];
        tile_sums[i] = run;
        run = op(run, v);
    }
    if (tid == SCAN_MID_THREADS - 1) tile_sums[nt] = op(warp_tot[SCAN_MID_THREADS / 32 - 1], identity);
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

// ---- single pass: chained scan with decoupled look-back ----------------------------------------------------------
// One launch: tiles are numbered by an atomic ticket (so a tile only ever waits for tiles that already run), each
// publishes its aggregate, then its inclusive prefix, in global memory; a tile's first warp looks back over its
// predecessors 32 at a time.  The input is read once (the reduce / scan-of-totals / down-sweep form above reads it
// twice).  Works for non-commutative operators: look-back values are combined
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

template <class T, class Op, class InF, class OutF>
inline cudaError_t device_scan_1p(ms_ctx* c, InF in, OutF out, int64_t n, T identity, Op op, DevBuf& tmp, T** d_total) {
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

// Returns a device pointer to the grand total (valid until tmp is reused).
// Three launches: tile reduce, one-CTA scan of the tile sums, down-sweep (the input is evaluated twice).  The single
// pass above is kept for reference: on this part it LOSES (0.78 vs 0.39 ms for the compaction scan, 1.02 vs 0.55 ms
// for the plan scan, profiles/r2c_scan_metrics.txt) — with ~300 tiles resident every look-back has to walk ~9 windows
// of 32 predecessors before it meets a finished one, each window two dependent L2 round trips plus a fence, while the
// other seven warps of the tile wait at the barrier (stall_barrier 52).
// First two phases: per-tile reduce, then the exclusive scan of the tile sums.  *tile_prefix gets nt prefixes followed
// by the grand total; the caller launches the down-sweep (the generic k_scan_down, or a kernel written for its types).
// middle phase alone: ts[0..nt) tile sums -> exclusive prefixes in place, ts[nt] = grand total
template <class T, class Op>
inline cudaError_t scan_mid_phase(ms_ctx* c, T* ts, int64_t nt, T identity, Op op) {
    if (nt > 4096) {
        // many tile sums: scan them with the single-pass kernel (a handful of tiles: its look-back is one window deep);
        // one CTA walking 20 k sums costs 65-98 us (profiles/r1n_launches.csv, r2e_launches.csv)
        const int64_t nt2 = ceil_div(nt, SCAN_TILE);
        const size_t status_bytes = ((size_t)nt2 * 4 + 15) & ~(size_t)15;
        cudaError_t e2 = c->scan_mid.ensure(16 + status_bytes + (size_t)(2 * nt2 + 1) * sizeof(T) + 16);
        if (e2 != cudaSuccess) return e2;
        uint8_t* base = c->scan_mid.as<uint8_t>();
        uint32_t* status = reinterpret_cast<uint32_t*>(base + 16);
        T* agg = reinterpret_cast<T*>(base + 16 + status_bytes);
        e2 = cudaMemsetAsync(base, 0, 16 + status_bytes, c->stream);
        if (e2 != cudaSuccess) return e2;
        auto in2 = [=] __device__(int64_t i) -> T { return ts[i]; };
        auto out2 = [=] __device__(int64_t i, T ex, T) { ts[i] = ex; };
        k_scan_1p<T, Op, decltype(in2), decltype(out2)><<<(unsigned)nt2, SCAN_THREADS, 0, c->stream>>>(
            in2, out2, nt, identity, op, reinterpret_cast<uint32_t*>(base), status, agg, agg + nt2, ts + nt, nt2);
    } else {
        k_scan_tiles<T, Op><<<1, SCAN_MID_THREADS, 0, c->stream>>>(ts, nt, identity, op);
    }
    c->kernel_launches++;
    return cudaGetLastError();
}

template <class T, class Op, class InF>
inline cudaError_t scan_tile_sums(ms_ctx* c, InF in, int64_t n, T identity, Op op, DevBuf& tmp, T** tile_prefix, int64_t* n_tiles) {
    const int64_t nt = ceil_div(n, SCAN_TILE);
    cudaError_t e = tmp.ensure((size_t)(nt + 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    T* ts = tmp.as<T>();
    if (nt > 0) { k_scan_reduce<T, Op, InF><<<(unsigned)nt, SCAN_THREADS, 0, c->stream>>>(in, n, identity, op, ts); c->kernel_launches++; }
    *tile_prefix = ts;
    *n_tiles = nt;
    return scan_mid_phase<T>(c, ts, nt, identity, op);
}

template <class T, class Op, class InF, class OutF>
inline cudaError_t device_scan(ms_ctx* c, InF in, OutF out, int64_t n, T identity, Op op, DevBuf& tmp, T** d_total) {
    T* ts = nullptr;
    int64_t nt = 0;
    cudaError_t e = scan_tile_sums<T>(c, in, n, identity, op, tmp, &ts, &nt);
    if (e != cudaSuccess) return e;
    if (nt > 0) { k_scan_down<T, Op, InF, OutF><<<(unsigned)nt, SCAN_THREADS, 0, c->stream>>>(in, out, n, identity, op, ts); c->kernel_launches++; }
    if (d_total) *d_total = ts + nt;
    return cudaGetLastError();
}

}  // namespace ms
