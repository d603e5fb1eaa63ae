// csrc/imomd_kernels.cu -- CUDA kernels (sm_100a) and the C ABI of include/imomd_b200.h.
//
// Kernels
//   k_build_vrec     unit vectors + hash cell of every vertex (graph upload)
//   k_lower_bounds   per (query, tree root, cell) haversine lower bounds
//   k_init_trees     free-chunk stacks + tree construction (ImomdRRT ctor :73-78)
//   k_expand         persistent expansion kernel: one CTA per query, `iters` passes
//                    of expandTreeLayers_ (:329-356) without returning to the host
//   k_nearest        parity hook: batched nearest-frontier-vertex lookups
//   k_walk           parent-pointer walk for updatePath_ (:936-954)
//   k_ga_eval        one warp per tour, sequential path cost (eci :442-450)
//
// There is no CPU fallback in this library: every entry point that computes needs a
// CUDA device and fails with IMOMD_ERR_CUDA otherwise.

#include <cuda_runtime.h>
#include <time.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/imomd_b200.h"
#include "imomd_core.cuh"
#include "imomd_setup.h"
#include "imomd_nn_tiles.cuh"

using namespace imomd;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                         \
    do                                                                                   \
    {                                                                                    \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call, __FILE__, __LINE__);          \
    } while (0)

// a failed CUDA call: message, sticky error cleared, device OOM reported as IMOMD_ERR_NOMEM
int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    (void)cudaGetLastError();
    return fail(e == cudaErrorMemoryAllocation ? IMOMD_ERR_NOMEM : IMOMD_ERR_CUDA, "%s failed: %s (%s:%d)", what,
                cudaGetErrorString(e), file, line);
}

// like CU, but runs `cleanup` before returning (constructors: nothing half-built may leak)
#define CU_OR(cleanup, call)                                                             \
    do                                                                                   \
    {                                                                                    \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
        {                                                                                \
            int rc_ = cuda_fail(e_, #call, __FILE__, __LINE__);                          \
            cleanup;                                                                     \
            return rc_;                                                                  \
        }                                                                                \
    } while (0)

constexpr int kThreads = 256;   // CTA size of the expansion kernel (8 warps)
constexpr int kTieBoxes = 4096; // host mailboxes per planner (>= 8 x the largest grid of the batched kernel)

// ------------------------------------------------------------ Cta policy ---

struct CtaDev
{
    __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
    __device__ __forceinline__ int nt() const { return (int)blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ long long clock() const { return clock64(); }
    __device__ __forceinline__ int warp() const { return (int)(threadIdx.x >> 5); }
    __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31); }
    __device__ __forceinline__ int warp_width() const { return 32; }
    __device__ __forceinline__ void warp_sync() const { __syncwarp(); }
    __device__ __forceinline__ uint32_t warp_bcast(uint32_t v) const { return __shfl_sync(0xFFFFFFFFu, v, 0); }
    __device__ __forceinline__ bool warp_all(bool p) const { return __all_sync(0xFFFFFFFFu, p) != 0; }
    // exclusive scan of per-lane counts in [0, maxv]: one ballot per possible count
    __device__ __forceinline__ uint32_t warp_scan_small(uint32_t v, int maxv, uint32_t* total) const
    {
        const unsigned full = 0xFFFFFFFFu;
        unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
        uint32_t off = 0, tot = 0;
        for (int j = 1; j <= maxv; ++j)
        {
            unsigned b = __ballot_sync(full, v >= (uint32_t)j);
            off += __popc(b & lt);
            tot += __popc(b);
        }
        *total = tot;
        return off;
    }
    __device__ __forceinline__ uint32_t warp_scan_exclusive(uint32_t v, uint32_t* total) const
    {
        const unsigned full = 0xFFFFFFFFu;
        int ln = threadIdx.x & 31;
        uint32_t inc = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1)
        {
            uint32_t t = __shfl_up_sync(full, inc, off);
            if (ln >= off) inc += t;
        }
        *total = __shfl_sync(full, inc, 31);
        return inc - v;
    }
    __device__ __forceinline__ uint32_t counter_add(uint32_t* p, uint32_t v) const
    {
        return atomicAdd(p, v);
    }
    __device__ __forceinline__ void counter_or(uint32_t* p, uint32_t v) const { atomicOr(p, v); }
    __device__ __forceinline__ int group() const { return 0; }
    __device__ __forceinline__ int n_groups() const { return 1; }
    __device__ __forceinline__ uint64_t now_ns() const
    {
        uint64_t t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return t;
    }
    // 128-bit accesses that should not displace the resident working sets from L2 (evict-first)
    __device__ __forceinline__ void store_streaming(U4* p, const U4& v) const
    {
        __stcs(reinterpret_cast<uint4*>(p), make_uint4(v.x, v.y, v.z, v.w));
    }
    __device__ __forceinline__ U4 load_streaming(const U4* p) const
    {
        uint4 t = __ldcs(reinterpret_cast<const uint4*>(p));
        U4 r;
        r.x = t.x; r.y = t.y; r.z = t.z; r.w = t.w;
        return r;
    }
    __device__ __forceinline__ uint32_t bcast(uint32_t v) const { return __shfl_sync(0xFFFFFFFFu, v, 0); }
    // batched mode: lexicographic-min update of a record head {parent, frontier slot, cost}: one
    // 128-bit compare-and-swap at the L2 (returns what was there)
    __device__ __forceinline__ U4 cas_head(U4* addr, const U4& expect, const U4& want) const
    {
        unsigned long long elo = ((unsigned long long)expect.y << 32) | expect.x;
        unsigned long long ehi = ((unsigned long long)expect.w << 32) | expect.z;
        unsigned long long wlo = ((unsigned long long)want.y << 32) | want.x;
        unsigned long long whi = ((unsigned long long)want.w << 32) | want.z;
        unsigned long long olo, ohi;
        asm volatile(
            "{\n\t"
            ".reg .b128 e, w, o;\n\t"
            "mov.b128 e, {%2, %3};\n\t"
            "mov.b128 w, {%4, %5};\n\t"
            "atom.relaxed.gpu.global.cas.b128 o, [%6], e, w;\n\t"
            "mov.b128 {%0, %1}, o;\n\t"
            "}"
            : "=l"(olo), "=l"(ohi)
            : "l"(elo), "l"(ehi), "l"(wlo), "l"(whi), "l"(addr)
            : "memory");
        U4 r;
        r.x = (uint32_t)olo; r.y = (uint32_t)(olo >> 32); r.z = (uint32_t)ohi; r.w = (uint32_t)(ohi >> 32);
        return r;
    }
    __device__ __forceinline__ U4 load_cg(const U4* p) const
    {
        uint4 t = __ldcg(reinterpret_cast<const uint4*>(p));
        U4 r;
        r.x = t.x; r.y = t.y; r.z = t.z; r.w = t.w;
        return r;
    }
    __device__ __forceinline__ void atomic_min_u64(unsigned long long* p, unsigned long long v) const { atomicMin(p, v); }
    __device__ __forceinline__ void atomic_min_u32(uint32_t* p, uint32_t v) const { atomicMin(p, v); }
    __device__ __forceinline__ void fence() const { __threadfence(); }
    __device__ __forceinline__ void prefetch_l2(const void* p) const { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
    // Exact values from the host (ONE thread calls): the n coordinate pairs go to this CTA's
    // mailbox in mapped pinned host memory, the host thread waiting for the launch (wait_stream)
    // answers with the reference's haversine evaluated by glibc.  System-scope release / acquire.
    TieBox* box = nullptr;
    __device__ __noinline__ bool certify(int n, double (*in)[4], double* out) const
    {
        if (box == nullptr) return false;
        volatile TieBox* b = box;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < 4; ++j) b->in[i][j] = in[i][j];
        b->n = (uint32_t)n;
        __threadfence_system();
        uint32_t* st = &box->state;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(st), "r"(1u) : "memory");
        uint32_t v = 0;
        for (;;)
        {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(st) : "memory");
            if (v == 2u) break;
            __nanosleep(2000);
        }
        for (int i = 0; i < n; ++i) out[i] = b->out[i];
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(st), "r"(0u) : "memory");
        return true;
    }
    __device__ __forceinline__ uint64_t reduce_sum_u64(uint64_t v, Best* red) const
    {
        const unsigned full = 0xFFFFFFFFu;
        int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(full, v, off);
        uint64_t* scr = reinterpret_cast<uint64_t*>(red);
        if (lane == 0) scr[warp] = v;
        __syncthreads();
        if (warp == 0)
        {
            v = lane < nw ? scr[lane] : 0ull;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(full, v, off);
            if (lane == 0) scr[32] = v;
        }
        __syncthreads();
        v = scr[32];
        __syncthreads();
        return v;
    }
    // block-wide exclusive prefix sum (returns this thread's offset, *total to everyone)
    __device__ __forceinline__ uint32_t scan_exclusive(uint32_t v, uint32_t* total, uint32_t* scr) const
    {
        const unsigned full = 0xFFFFFFFFu;
        int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        uint32_t inc = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1)
        {
            uint32_t t = __shfl_up_sync(full, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) scr[warp] = inc;
        __syncthreads();
        if (warp == 0)
        {
            uint32_t w = lane < nw ? scr[lane] : 0u;
            uint32_t winc = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1)
            {
                uint32_t t = __shfl_up_sync(full, winc, off);
                if (lane >= off) winc += t;
            }
            if (lane < nw) scr[lane] = winc - w;     // exclusive warp offsets
            if (lane == nw - 1) scr[32] = winc;      // grand total
        }
        __syncthreads();
        uint32_t res = scr[warp] + inc - v;
        *total = scr[32];
        __syncthreads();
        return res;
    }
    __device__ __forceinline__ uint32_t reduce_min_u32(uint32_t v, uint32_t* scr) const
    {
        const unsigned full = 0xFFFFFFFFu;
        int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        v = __reduce_min_sync(full, v);
        if (lane == 0) scr[warp] = v;
        __syncthreads();
        if (warp == 0)
        {
            uint32_t w = lane < nw ? scr[lane] : 0xFFFFFFFFu;
            w = __reduce_min_sync(full, w);
            if (lane == 0) scr[32] = w;
        }
        __syncthreads();
        v = scr[32];
        __syncthreads();
        return v;
    }
    __device__ __forceinline__ double reduce_min_f64(double v, Best* red) const
    {
        const unsigned full = 0xFFFFFFFFu;
        int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = fmin(v, __shfl_down_sync(full, v, off));
        if (lane == 0) red[warp].v1 = v;
        __syncthreads();
        if (warp == 0)
        {
            v = lane < nw ? red[lane].v1 : INFINITY;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v = fmin(v, __shfl_down_sync(full, v, off));
            if (lane == 0) red[0].v2 = v;
        }
        __syncthreads();
        v = red[0].v2;
        __syncthreads();
        return v;
    }
    // block-wide merge; every thread gets the result.  Warp shuffles, then one slot
    // per warp in shared memory.
    __device__ __forceinline__ Best reduce(Best b, Best* red) const
    {
        const unsigned full = 0xFFFFFFFFu;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
        {
            Best o;
            o.v1 = __shfl_down_sync(full, b.v1, off);
            o.id1 = __shfl_down_sync(full, b.id1, off);
            o.v2 = __shfl_down_sync(full, b.v2, off);
            b = best_merge(b, o);
        }
        int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        if (lane == 0) red[warp] = b;
        __syncthreads();
        if (warp == 0)
        {
            b = lane < nw ? red[lane] : best_empty();
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
            {
                Best o;
                o.v1 = __shfl_down_sync(full, b.v1, off);
                o.id1 = __shfl_down_sync(full, b.id1, off);
                o.v2 = __shfl_down_sync(full, b.v2, off);
                b = best_merge(b, o);
            }
            if (lane == 0) red[0] = b;
        }
        __syncthreads();
        b = red[0];
        __syncthreads();
        return b;
    }
};

// The cooperating group of the BATCHED mode: one warp (one tree).  Same interface as CtaDev; every
// barrier is a warp barrier, every reduction a shuffle tree, scratch is the warp's own Shared.
struct WarpDev
{
    TieBox* box = nullptr;
    __device__ __forceinline__ int tid() const { return (int)(threadIdx.x & 31); }
    __device__ __forceinline__ int nt() const { return 32; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ long long clock() const { return clock64(); }
    __device__ __forceinline__ int warp() const { return 0; }
    __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31); }
    __device__ __forceinline__ int warp_width() const { return 32; }
    __device__ __forceinline__ void warp_sync() const { __syncwarp(); }
    __device__ __forceinline__ int group() const { return (int)(threadIdx.x >> 5); }
    __device__ __forceinline__ int n_groups() const { return (int)(blockDim.x >> 5); }
    __device__ __forceinline__ bool warp_all(bool p) const { return __all_sync(0xFFFFFFFFu, p) != 0; }
    __device__ __forceinline__ uint32_t bcast(uint32_t v) const { return __shfl_sync(0xFFFFFFFFu, v, 0); }
    __device__ __forceinline__ uint32_t warp_scan_small(uint32_t v, int maxv, uint32_t* total) const
    {
        const unsigned full = 0xFFFFFFFFu;
        unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
        uint32_t off = 0, tot = 0;
        for (int j = 1; j <= maxv; ++j)
        {
            unsigned b = __ballot_sync(full, v >= (uint32_t)j);
            off += __popc(b & lt);
            tot += __popc(b);
        }
        *total = tot;
        return off;
    }
    __device__ __forceinline__ uint32_t counter_add(uint32_t* p, uint32_t v) const { return atomicAdd(p, v); }
    __device__ __forceinline__ void counter_or(uint32_t* p, uint32_t v) const { atomicOr(p, v); }
    // the relaxation of one tree runs on the tree's group (batch_relax_tree): same record-head
    // primitives as the block's
    __device__ __forceinline__ U4 cas_head(U4* addr, const U4& expect, const U4& want) const
    {
        CtaDev c;
        return c.cas_head(addr, expect, want);
    }
    __device__ __forceinline__ U4 load_cg(const U4* p) const
    {
        CtaDev c;
        return c.load_cg(p);
    }
    __device__ __forceinline__ void prefetch_l2(const void* p) const { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
    __device__ __forceinline__ uint64_t now_ns() const
    {
        uint64_t t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return t;
    }
    __device__ __forceinline__ void store_streaming(U4* p, const U4& v) const
    {
        __stcs(reinterpret_cast<uint4*>(p), make_uint4(v.x, v.y, v.z, v.w));
    }
    __device__ __forceinline__ U4 load_streaming(const U4* p) const
    {
        uint4 t = __ldcs(reinterpret_cast<const uint4*>(p));
        U4 r;
        r.x = t.x; r.y = t.y; r.z = t.z; r.w = t.w;
        return r;
    }
    __device__ __noinline__ bool certify(int n, double (*in)[4], double* out) const
    {
        CtaDev c;
        c.box = box;
        return c.certify(n, in, out);
    }
    __device__ __forceinline__ uint32_t scan_exclusive(uint32_t v, uint32_t* total, uint32_t*) const
    {
        const unsigned full = 0xFFFFFFFFu;
        int ln = threadIdx.x & 31;
        uint32_t inc = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1)
        {
            uint32_t t = __shfl_up_sync(full, inc, off);
            if (ln >= off) inc += t;
        }
        *total = __shfl_sync(full, inc, 31);
        return inc - v;
    }
    __device__ __forceinline__ uint32_t reduce_min_u32(uint32_t v, uint32_t*) const
    {
        return __reduce_min_sync(0xFFFFFFFFu, v);
    }
    __device__ __forceinline__ double reduce_min_f64(double v, Best*) const
    {
        const unsigned full = 0xFFFFFFFFu;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(full, v, off));
        return v;
    }
    __device__ __forceinline__ uint64_t reduce_sum_u64(uint64_t v, Best*) const
    {
        const unsigned full = 0xFFFFFFFFu;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(full, v, off);
        return v;
    }
    __device__ __forceinline__ Best reduce(Best b, Best*) const
    {
        const unsigned full = 0xFFFFFFFFu;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
        {
            Best o;
            o.v1 = __shfl_xor_sync(full, b.v1, off);
            o.id1 = __shfl_xor_sync(full, b.id1, off);
            o.v2 = __shfl_xor_sync(full, b.v2, off);
            b = best_merge(b, o);
        }
        return b;
    }
};

// ---------------------------------------------------------------- kernels ---

__global__ void k_build_vrec(uint32_t n, const double* __restrict__ lat,
                             const double* __restrict__ lon, VRec* __restrict__ vrec,
                             double* __restrict__ coslat, int G, double lat0, double lon0, double hlat,
                             double hlon)
{
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    vrec[v] = make_vrec(v, lat[v], lon[v], G, lat0, lon0, hlat, hlon);
    coslat[v] = cos(lat[v] * kDeg2Rad);   // the operand of imomd::haversine's cos() calls
}

// lbd[q][t][c] = lower bound (metres) of haversine(root of tree t, any point of cell c)
__global__ void k_lower_bounds(const __grid_constant__ PlannerDev P)
{
    size_t total = (size_t)P.Q * P.K * P.g.G * P.g.G;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    P.lbd[i] = lbd_entry(P, i);
    P.ubd[i] = ubd_entry(P, i);
}

// Freshly constructed headers (src/imomd_rrt_star.cpp:57-71, :103-108) from the destinations and
// the selection probabilities: one CTA per query; only 4 K + 8 K^2 bytes per query cross the bus.
__global__ void k_fill_headers(const __grid_constant__ PlannerDev P, QueryDev* __restrict__ out,
                               const uint32_t* __restrict__ dest, const double* __restrict__ prob)
{
    const int q = blockIdx.x, K = P.K;
    QueryDev* h = out + q;
    uint32_t* w = reinterpret_cast<uint32_t*>(h);
    for (uint32_t i = threadIdx.x; i < sizeof(QueryDev) / 4; i += blockDim.x) w[i] = 0u;
    __syncthreads();
    if (threadIdx.x == 0) h->K = K;
    for (int i = threadIdx.x; i < K; i += blockDim.x) h->dest[i] = dest[(size_t)q * K + i];
    for (int i = threadIdx.x; i < K * K; i += blockDim.x)
    {
        h->P[i] = prob[(size_t)q * K * K + i];
        h->D[i] = (i / K == i % K) ? 0.0 : INFINITY;
        h->MHv[i] = INFINITY;
        h->MHi[i] = IMOMD_NONE;
        h->CN[i] = IMOMD_NONE;
    }
}

__global__ void k_init_trees(const __grid_constant__ PlannerDev P)
{
    int q = blockIdx.x;
    int K = P.K;
    // free-chunk stacks: pop order 0, 1, 2, ...
    for (int k = 0; k < K; ++k)
    {
        uint32_t* fs = P.free_stack + ((size_t)q * K + k) * P.chunks;
        for (int i = threadIdx.x; i < P.chunks; i += blockDim.x) fs[i] = (uint32_t)(P.chunks - 1 - i);
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        QueryDev* Q = P.query + q;
        for (int k = 0; k < K; ++k) Q->free_top[k] = (uint32_t)P.chunks;
        construct_trees(P, q, q);
    }
}

// The expansion kernel in two register budgets: a latency build (one query per SM, all
// 255 registers, no spills) for planners with few queries, and a throughput build (two
// CTAs per SM, 128 registers) whose point is to hide each query's dependent-load latency behind
// the other resident query.
// The planner descriptor is a __grid_constant__ parameter: the device logic takes it by reference
// (noinline services), and without the qualifier every field read would go through a per-thread
// LOCAL copy of the 200-byte struct (measured: 6.5 % of the batch time, 300 bytes of stack).
template <int THREADS, int MIN_BLOCKS, int KW>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) k_expand(const __grid_constant__ PlannerDev P, int iters)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared& S = *reinterpret_cast<Shared*>(smem_raw);
    double* lb = reinterpret_cast<double*>(smem_raw + ((sizeof(Shared) + 15) / 16) * 16);
    uint32_t* lvl = reinterpret_cast<uint32_t*>(lb + lb_entries(P));
    double* mat = reinterpret_cast<double*>(lvl + 2 * kLevelCap);
    unsigned char* hot = reinterpret_cast<unsigned char*>(mat + (size_t)P.K * P.K * 4);
    CtaDev cta;
    cta.box = P.tiebox ? P.tiebox + blockIdx.x : nullptr;
    if (P.streaming)
    {
        // slot = this CTA; queries come from the work queue in submission order, so a slot that
        // drew a short query simply takes the next one (no wave boundaries, no static tail)
        __shared__ int next_q;
        for (;;)
        {
            if (threadIdx.x == 0) next_q = (int)atomicAdd(P.queue, 1u);
            __syncthreads();
            const int hq = next_q;
            __syncthreads();
            if (hq >= P.Q) break;
            run_streaming_query<KW>(cta, P, (int)blockIdx.x, hq, iters, S, lb, lvl, mat, hot);
        }
        return;
    }
    for (int q = blockIdx.x; q < P.Q; q += gridDim.x)
    {
        run_query<KW>(cta, P, q, q, iters, S, lb, lvl, mat, hot);
        __syncthreads();
    }
}

// imomd_expand_batch: one CTA per query (resident) or per state slot (streaming), one WARP per
// tree inside it; `steps` steps of B expansions per tree (run_query_batched, imomd_core.cuh).
// Shared memory: [Shared x warps | BatchShared | level buffers x warps | K x K matrices | hot header]
template <int THREADS, int MIN_BLOCKS, int KW>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) k_expand_batch(const __grid_constant__ PlannerDev P, int steps, int B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = THREADS / 32;
    constexpr size_t kShared = ((sizeof(Shared) + 15) / 16) * 16;
    Shared* SG = reinterpret_cast<Shared*>(smem_raw);
    BatchShared& BS = *reinterpret_cast<BatchShared*>(smem_raw + NW * kShared);
    double* lb_all = reinterpret_cast<double*>(smem_raw + NW * kShared + ((sizeof(BatchShared) + 15) / 16) * 16);
    double* mat = lb_all + (size_t)NW * kGroupLbCap;
    unsigned char* hot = reinterpret_cast<unsigned char*>(mat + (size_t)P.K * P.K * 4);
    CtaDev blk;
    WarpDev grp;
    // mailboxes of the exact-value service: eight per CTA (the block's = its first warp's)
    blk.box = P.tiebox ? P.tiebox + (size_t)blockIdx.x * 8 : nullptr;
    grp.box = P.tiebox ? P.tiebox + (size_t)blockIdx.x * 8 + (threadIdx.x >> 5) % 8 : nullptr;
    if (P.streaming)
    {
        __shared__ int next_q;
        for (;;)
        {
            if (threadIdx.x == 0) next_q = (int)atomicAdd(P.queue, 1u);
            __syncthreads();
            const int hq = next_q;
            __syncthreads();
            if (hq >= P.Q) break;
            run_streaming_query_batched<KW>(blk, grp, P, (int)blockIdx.x, hq, steps, B, SG, BS, lb_all, mat, hot);
        }
        return;
    }
    for (int q = blockIdx.x; q < P.Q; q += gridDim.x)
    {
        run_query_batched<KW>(blk, grp, P, q, q, steps, B, SG, BS, lb_all, mat, hot);
        __syncthreads();
    }
}

// TREEHASH + checksum of one resident query (imomd_get_treehash)
__global__ void __launch_bounds__(kThreads) k_tree_hash(const __grid_constant__ PlannerDev P, int q)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared& S = *reinterpret_cast<Shared*>(smem_raw);
    CtaDev cta;
    uint64_t h = 0, sum = 0;
    tree_hash(cta, P, q, S, P.arena + (size_t)q * P.arena_entries, P.arena_entries, &h, &sum);
    if (threadIdx.x == 0)
    {
        P.query[q].tree_hash = h;
        P.query[q].tree_sum = sum;
    }
}

// mergeTwoTree_'s per-vertex work (pseudo-tree merge driven by the host facade)
template <int KW>
__global__ void __launch_bounds__(kThreads) k_attach(const __grid_constant__ PlannerDev P, int q, int tree, uint32_t count)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared& S = *reinterpret_cast<Shared*>(smem_raw);
    double* lb = reinterpret_cast<double*>(smem_raw + ((sizeof(Shared) + 15) / 16) * 16);
    uint32_t* lvl = reinterpret_cast<uint32_t*>(lb + lb_entries(P));
    double* mat = reinterpret_cast<double*>(lvl + 2 * kLevelCap);
    unsigned char* hot = reinterpret_cast<unsigned char*>(mat + (size_t)P.K * P.K * 4);
    CtaDev cta;
    cta.box = P.tiebox ? P.tiebox + blockIdx.x : nullptr;
    run_query<KW>(cta, P, q, q, 0, S, lb, lvl, mat, hot, tree, count);
}

// Sparse readback: parent and cost of the listed vertices of one tree (ids in, parents out, in place).
__global__ void k_gather_tree(const __grid_constant__ PlannerDev P, size_t qt, uint32_t* ids, double* cost, uint32_t count)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const unsigned char* rec = P.trec + (qt * P.g.n + ids[i]) * P.trec_stride;
    uint32_t par = *reinterpret_cast<const uint32_t*>(rec);
    ids[i] = par;
    cost[i] = par == IMOMD_NONE ? INFINITY : *reinterpret_cast<const double*>(rec + 8);
}

// Frontier of one tree as a compact id list (order unspecified): out[0] = count, ids from out[1].
__global__ void k_frontier_list(const __grid_constant__ PlannerDev P, size_t qt, uint32_t* out)
{
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    bool in = false;
    if (v < P.g.n)
        in = *reinterpret_cast<const uint32_t*>(P.trec + (qt * P.g.n + v) * P.trec_stride + 4) != IMOMD_NONE;
    unsigned m = __ballot_sync(0xffffffffu, in);
    if (m == 0) return;
    int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(out, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (in) out[1 + base + __popc(m & ((1u << lane) - 1))] = v;
}

// One CTA per lookup slot (grid-stride over the lookups); scratch chunk lists are per CTA.
__global__ void __launch_bounds__(kThreads)
k_nearest(const __grid_constant__ PlannerDev P, int n, const imomd_nn_query* __restrict__ qs, uint32_t* __restrict__ idx,
          double* __restrict__ dist, uint32_t* scratch, int64_t* stats)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared& S = *reinterpret_cast<Shared*>(smem_raw);
    double* lb = reinterpret_cast<double*>(smem_raw + ((sizeof(Shared) + 15) / 16) * 16);
    CtaDev cta;
    cta.box = P.tiebox ? P.tiebox + blockIdx.x : nullptr;
    uint32_t* clist = scratch + (size_t)blockIdx.x * 2 * P.chunks;
    __shared__ int64_t near_ties, unresolved;
    if (threadIdx.x == 0)
    {
        S.tref = nullptr;   // no per-launch table of tree views here
        near_ties = 0;
        unresolved = 0;
    }
    for (int i = blockIdx.x; i < n; i += gridDim.x)
    {
        imomd_nn_query nq = qs[i];
        if (threadIdx.x == 0)
        {
            S.Q = P.query + nq.query;
            S.qlat = nq.lat;
            S.qlon = nq.lon;
        }
        __syncthreads();
        svc_nearest(cta, P, nq.query, nq.tree, S, P.query[nq.query].n_occ[nq.tree] <= lb_entries(P) ? lb : nullptr, clist,
                    &near_ties, &unresolved, false);
        __syncthreads();
        if (threadIdx.x == 0)
        {
            uint32_t v = S.result.id1;
            idx[i] = v;
            dist[i] = v == IMOMD_NONE
                          ? INFINITY
                          : haversine(P.g.lat[v], P.g.lon[v], nq.lat, nq.lon);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && stats)
    {
        atomicAdd((unsigned long long*)&stats[0], (unsigned long long)near_ties);
        atomicAdd((unsigned long long*)&stats[1], (unsigned long long)unresolved);
    }
}

__global__ void k_walk(const unsigned char* __restrict__ trec, uint32_t stride, uint32_t vertex,
                       uint32_t root, uint32_t* out, uint32_t cap, uint32_t* len)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t n = 0;
    uint32_t v = vertex;
    if (n < cap) out[n] = v;
    ++n;
    while (v != root && v != IMOMD_NONE)
    {
        v = *reinterpret_cast<const uint32_t*>(trec + (size_t)v * stride);
        if (n < cap) out[n] = v;
        ++n;
        if (n > 0x7FFFFFF0u) break;
    }
    *len = n;
}

// one warp per tour; lanes gather the leg lengths, the sum is formed left to right
constexpr int kGaMaxLen = 128;
__global__ void __launch_bounds__(256)
k_ga_eval(int K, const double* __restrict__ D, int n_tours, int stride,
          const int32_t* __restrict__ tours, const int32_t* __restrict__ lens,
          double* __restrict__ cost, int* __restrict__ bad)
{
    extern __shared__ double sD[];
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) sD[i] = D[i];
    __syncthreads();
    const unsigned full = 0xFFFFFFFFu;
    int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int t = warp; t < n_tours; t += nwarps)
    {
        const int32_t* s = tours + (size_t)t * stride;
        int len = lens[t];
        bool ok = len >= 1 && len <= stride;
        int legs = ok ? len - 1 : 0;
        double vals[kGaMaxLen / 32];
#pragma unroll
        for (int j = 0; j < kGaMaxLen / 32; ++j)
        {
            int i = j * 32 + lane;
            vals[j] = 0.0;
            if (i < legs)
            {
                // ids are validated here (the host no longer walks every tour before the launch)
                unsigned a = (unsigned)s[i], b = (unsigned)s[i + 1];
                if (a < (unsigned)K && b < (unsigned)K) vals[j] = sD[a * K + b];
                else ok = false;
            }
        }
        if (legs == 0 && ok && lane == 0 && (unsigned)s[0] >= (unsigned)K) ok = false;
        if (!__all_sync(full, ok))
        {
            if (lane == 0)
            {
                atomicExch(bad, t + 1);
                cost[t] = nan("");
            }
            continue;
        }
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < kGaMaxLen / 32; ++j)
        {
            for (int l = 0; l < 32; ++l)
            {
                double v = __shfl_sync(full, vals[j], l);
                if (j * 32 + l < legs) acc += v;
            }
        }
        if (lane == 0) cost[t] = acc;
    }
}

}  // namespace

// ================================================================= C ABI =====

struct imomd_graph
{
    int device;
    GraphDev d;
    uint8_t* deg;
    uint32_t* col;
    double* w;
    double* lat;
    double* lon;
    VRec* vrec;
    double* coslat;
    double* cosrow;
    uint32_t max_degree;
    int sm_count;
};

// device scratch of imomd_nearest_batch: grow-only buffers and the events of its stage timers,
// owned by the planner (a call allocates only when it needs more than any call before it)
struct NnPool
{
    static constexpr int kSlots = 20;
    void* ptr[kSlots] = {};
    size_t bytes[kSlots] = {};
    cudaEvent_t ev[6] = {};
    bool have_events = false;
    template <class T>
    cudaError_t get(int slot, T** x, size_t count)
    {
        const size_t need = std::max<size_t>(count, 1) * sizeof(T);
        if (need > bytes[slot])
        {
            cudaFree(ptr[slot]);
            ptr[slot] = nullptr;
            bytes[slot] = 0;
            const size_t want = need + need / 4;
            cudaError_t rc = cudaMalloc(&ptr[slot], want);
            if (rc != cudaSuccess) return rc;
            bytes[slot] = want;
        }
        *x = static_cast<T*>(ptr[slot]);
        return cudaSuccess;
    }
    cudaError_t events()
    {
        if (have_events) return cudaSuccess;
        for (cudaEvent_t& e : ev)
        {
            cudaError_t rc = cudaEventCreate(&e);
            if (rc != cudaSuccess) return rc;
        }
        have_events = true;
        return cudaSuccess;
    }
    void release()
    {
        for (int i = 0; i < kSlots; ++i)
        {
            cudaFree(ptr[i]);
            ptr[i] = nullptr;
            bytes[i] = 0;
        }
        for (cudaEvent_t& e : ev)
            if (e) cudaEventDestroy(e);
        have_events = false;
    }
};

struct imomd_planner
{
    imomd_graph* g;
    PlannerDev d;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    std::vector<uint32_t> dest_h;   // [Q][K] destinations and [Q][K][K] selection probabilities of the
    std::vector<double> prob_h;     // freshly constructed state (host copies; imomd_planner_reset)
    uint32_t* dest_d = nullptr;
    double* prob_d = nullptr;
    uint32_t* words;
    size_t words_cap, words_fed;
    size_t smem_bytes;
    float last_ms;
    bool pending;
    // lookup scratch (imomd_nearest)
    uint32_t* nn_scratch;
    int nn_ctas;
    int64_t* nn_stats;
    uint32_t* walk_buf;
    uint32_t walk_cap;
    // sparse readback scratch (imomd_gather_tree, imomd_get_frontier): ids / values, n + 1 each
    uint32_t* gather_u32;
    double* gather_f64;
    // device times of the last imomd_nearest_batch call
    float nn_prep_ms = 0.f, nn_tiles_ms = 0.f, nn_merge_ms = 0.f;
    uint32_t nn_segments = 0;
    NnPool nn_pool;
    // streaming planners: device copy of the freshly constructed headers, the work queue
    QueryDev* query_init = nullptr;
    uint32_t* queue = nullptr;
    // host mailboxes of the exact-value service (mapped pinned memory) and what they did
    TieBox* box_host = nullptr;
    uint64_t certified_requests = 0;
    size_t smem_batch = 0;          // dynamic shared memory of k_expand_batch
    bool batch_ready = false;       // event logs / per-tree chunk lists allocated
    bool batch_ran = false;         // resident planner: batched steps ran since the last reset (child rows
                                    // are not maintained there, so the exact mode may not continue)
};

struct imomd_ctx
{
    int device;
    int sm_count;
    cudaStream_t stream;
    double* D;
    int32_t* tours;
    int32_t* lens;
    double* cost;
    int* bad;                 // device flag: a tour held an id outside [0, K) or a bad length
    size_t cap_tours, cap_n;
    // pinned staging: one stream-ordered H2D / kernel / D2H sequence per call
    unsigned char* pin;
    size_t pin_bytes;
};

namespace {

template <class T>
cudaError_t dmalloc(T** p, size_t count)
{
    return cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
}

// Answers the mailboxes of the exact-value service: the reference's haversine
// (include/osm_converter/compute_haversine.hpp:41-58) on the host, i.e. with glibc's sin / cos /
// atan2 / sqrt -- imomd::haversine is the same expression, compiled for the host here.
void service_boxes(imomd_planner* p)
{
    TieBox* boxes = p->box_host;
    if (!boxes) return;
    for (int i = 0; i < kTieBoxes; ++i)
    {
        TieBox& b = boxes[i];
        if (__atomic_load_n(&b.state, __ATOMIC_ACQUIRE) != 1u) continue;
        const uint32_t n = std::min<uint32_t>(b.n, IMOMD_TIE_PAIRS);
        for (uint32_t j = 0; j < n; ++j) b.out[j] = haversine(b.in[j][0], b.in[j][1], b.in[j][2], b.in[j][3]);
        p->certified_requests += 1;
        __atomic_store_n(&b.state, 2u, __ATOMIC_RELEASE);
    }
}

// cudaStreamSynchronize that keeps the exact-value service running: a kernel of this planner may
// be waiting for an answer, so every wait on the planner's stream goes through here.
cudaError_t wait_stream(imomd_planner* p)
{
    if (!p->box_host) return cudaStreamSynchronize(p->stream);
    int spins = 0;
    for (;;)
    {
        cudaError_t rc = cudaStreamQuery(p->stream);
        if (rc != cudaErrorNotReady) return rc;
        service_boxes(p);
        if (++spins > 64)
        {
            struct timespec ts = {0, 20000};   // 20 us between polls once the wait is a long one
            nanosleep(&ts, nullptr);
        }
    }
}

int planner_fill_initial(imomd_planner* p)
{
    PlannerDev& d = p->d;
    p->batch_ran = false;
    const size_t QK = (size_t)d.Q * d.K;
    CU(cudaMemcpyAsync(p->dest_d, p->dest_h.data(), QK * sizeof(uint32_t), cudaMemcpyHostToDevice, p->stream));
    CU(cudaMemcpyAsync(p->prob_d, p->prob_h.data(), QK * d.K * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    if (d.streaming)
    {
        // slots are rebuilt by the kernel when a query takes them; headers are built on the device
        k_fill_headers<<<d.Q, 256, 0, p->stream>>>(d, p->query_init, p->dest_d, p->prob_d);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(d.query, p->query_init, (size_t)d.Q * sizeof(QueryDev), cudaMemcpyDeviceToDevice,
                           p->stream));
        return IMOMD_OK;
    }
    const size_t qk = (size_t)d.Q * d.K;
    const size_t n = d.g.n;
    const size_t cells = (size_t)d.g.G * d.g.G;
    // all-ones = not visited / not in the frontier / no children (the cost field is written
    // before it is ever read)
    CU(cudaMemsetAsync(d.trec, 0xFF, qk * n * d.trec_stride, p->stream));
    CU(cudaMemsetAsync(d.cell_cnt, 0, qk * cells * sizeof(uint32_t), p->stream));
    CU(cudaMemsetAsync(d.cell_head, 0xFF, qk * cells * sizeof(uint32_t), p->stream));
    k_fill_headers<<<d.Q, 256, 0, p->stream>>>(d, d.query, p->dest_d, p->prob_d);
    CU(cudaGetLastError());
    k_init_trees<<<d.Q, 128, 0, p->stream>>>(d);
    CU(cudaGetLastError());
    return IMOMD_OK;
}

}  // namespace

extern "C" {

const char* imomd_last_error(void) { return g_err; }

const char* imomd_version(void) { return "imomd_b200 0.1 (sm_100a)"; }

size_t imomd_sizeof_query_state(void) { return sizeof(QueryDev); }

int imomd_graph_create(int device, uint32_t n, const uint32_t* row_ptr, const uint32_t* col,
                       const double* w, const double* lat, const double* lon, imomd_graph** out)
{
    return imomd_graph_create_ex(device, n, row_ptr, col, w, lat, lon, 0, out);
}

int imomd_graph_create_ex(int device, uint32_t n, const uint32_t* row_ptr, const uint32_t* col,
                          const double* w, const double* lat, const double* lon, int grid_dim,
                          imomd_graph** out)
{
    if (!out || !row_ptr || !col || !w || !lat || !lon || n == 0)
        return fail(IMOMD_ERR_ARG, "imomd_graph_create: null argument or empty graph");
    *out = nullptr;
    const uint32_t arcs = row_ptr[n];
    uint32_t max_deg = 0;
    std::string verr;
    if (!validate_graph(n, row_ptr, col, &max_deg, &verr)) return fail(IMOMD_ERR_GRAPH, "%s", verr.c_str());
    CU(cudaSetDevice(device));
    imomd_graph* g = new (std::nothrow) imomd_graph();
    if (!g) return fail(IMOMD_ERR_NOMEM, "out of host memory");
    g->device = device;
    g->max_degree = max_deg;
#define GCU(call) CU_OR(imomd_graph_destroy(g), call)
    GCU(cudaDeviceGetAttribute(&g->sm_count, cudaDevAttrMultiProcessorCount, device));
    GraphDev& d = g->d;
    d.n = n;
    d.arcs = arcs;
    std::vector<double> cosrow;
    setup_grid(d, n, lat, lon, grid_dim, max_deg, &cosrow);
    const int G = d.G;
    // adjacency as padded rows of d.width slots (GraphDev)
    if ((uint64_t)n * (uint64_t)d.width > 0xFFFFFFF0ull)
    {
        delete g;
        return fail(IMOMD_ERR_GRAPH, "graph too large for 32-bit adjacency slots (n * width)");
    }
    std::vector<uint32_t> pcol;
    std::vector<double> pw;
    std::vector<uint8_t> pdeg;
    build_padded_rows(n, row_ptr, col, w, d.width, &pcol, &pw, &pdeg);
    GCU(dmalloc(&g->deg, (size_t)n));
    GCU(dmalloc(&g->col, pcol.size()));
    GCU(dmalloc(&g->w, pw.size()));
    GCU(dmalloc(&g->lat, n));
    GCU(dmalloc(&g->lon, n));
    GCU(dmalloc(&g->vrec, n));
    GCU(dmalloc(&g->coslat, n));
    GCU(dmalloc(&g->cosrow, 2 * (size_t)G));
    GCU(cudaMemcpy(g->deg, pdeg.data(), (size_t)n, cudaMemcpyHostToDevice));
    GCU(cudaMemcpy(g->col, pcol.data(), pcol.size() * 4, cudaMemcpyHostToDevice));
    GCU(cudaMemcpy(g->w, pw.data(), pw.size() * 8, cudaMemcpyHostToDevice));
    GCU(cudaMemcpy(g->lat, lat, (size_t)n * 8, cudaMemcpyHostToDevice));
    GCU(cudaMemcpy(g->lon, lon, (size_t)n * 8, cudaMemcpyHostToDevice));
    GCU(cudaMemcpy(g->cosrow, cosrow.data(), 2 * (size_t)G * sizeof(double), cudaMemcpyHostToDevice));
    d.deg = g->deg;
    d.col = g->col;
    d.w = g->w;
    d.lat = g->lat;
    d.lon = g->lon;
    d.vrec = g->vrec;
    d.coslat = g->coslat;
    d.cosrow = g->cosrow;
    k_build_vrec<<<(n + 255) / 256, 256>>>(n, g->lat, g->lon, g->vrec, g->coslat, G, d.lat0, d.lon0,
                                           d.hlat, d.hlon);
    GCU(cudaGetLastError());
    GCU(cudaDeviceSynchronize());
#undef GCU
    *out = g;
    return IMOMD_OK;
}

int imomd_graph_destroy(imomd_graph* g)
{
    if (!g) return IMOMD_OK;
    cudaSetDevice(g->device);
    cudaFree(g->deg);
    cudaFree(g->col);
    cudaFree(g->w);
    cudaFree(g->lat);
    cudaFree(g->lon);
    cudaFree(g->vrec);
    cudaFree(g->coslat);
    cudaFree(g->cosrow);
    delete g;
    return IMOMD_OK;
}

int imomd_graph_info(const imomd_graph* g, uint32_t* n, uint32_t* arcs, int32_t* grid_dim,
                     int32_t* device)
{
    if (!g) return fail(IMOMD_ERR_ARG, "null graph");
    if (n) *n = g->d.n;
    if (arcs) *arcs = g->d.arcs;
    if (grid_dim) *grid_dim = g->d.G;
    if (device) *device = g->device;
    return IMOMD_OK;
}

int imomd_planner_create(imomd_graph* g, int n_queries, int K, const uint32_t* dest,
                         const double* prob, const imomd_params* params, imomd_planner** out)
{
    if (!g || !dest || !prob || !params || !out)
        return fail(IMOMD_ERR_ARG, "imomd_planner_create: null argument");
    if (n_queries < 1) return fail(IMOMD_ERR_ARG, "n_queries must be positive");
    if (K < 2 || K > (int)IMOMD_KMAX) return fail(IMOMD_ERR_ARG, "K must be in [2, %u]", IMOMD_KMAX);
    for (size_t i = 0; i < (size_t)n_queries * K; ++i)
    {
        if (dest[i] >= g->d.n) return fail(IMOMD_ERR_ARG, "destination %u out of range", dest[i]);
    }
    if (params->pseudo_mode && params->max_resident_queries != 0 && params->max_resident_queries < n_queries)
        return fail(IMOMD_ERR_ARG, "pseudo mode needs a resident planner (the merge reads the trees back)");
    *out = nullptr;
    CU(cudaSetDevice(g->device));
    imomd_planner* p = new (std::nothrow) imomd_planner();
    if (!p) return fail(IMOMD_ERR_NOMEM, "out of host memory");
#define PCU(call) CU_OR(imomd_planner_destroy(p), call)
    p->g = g;
    p->words = nullptr;
    p->words_cap = p->words_fed = 0;
    p->pending = false;
    p->last_ms = 0.f;
    p->nn_scratch = nullptr;
    p->nn_ctas = 0;
    p->nn_stats = nullptr;
    p->walk_buf = nullptr;
    p->gather_u32 = nullptr;
    p->gather_f64 = nullptr;
    p->walk_cap = 0;
    p->stream = nullptr;
    p->ev0 = p->ev1 = nullptr;
    PlannerDev& d = p->d;
    std::memset(&d, 0, sizeof(d));
    d.g = g->d;
    d.Q = n_queries;
    d.K = K;
    d.goal_bias = params->goal_bias;
    d.tie_rel = params->tie_band > 0.0 ? params->tie_band : kTieRel;
    d.pseudo = params->pseudo_mode ? 1 : 0;
    d.profile = getenv("IMOMD_PROFILE") ? 1u : 0u;
    const size_t n = g->d.n;
    const size_t cells = (size_t)g->d.G * g->d.G;
    default_sizes(d, params->frontier_chunks, params->arena_entries, params->log_entries);
    const size_t chunks = (size_t)d.chunks;
    d.trec_stride = 16u + 4u * (uint32_t)g->d.width;
    // bytes of device state one resident query needs (trees, frontier hashes, scratch)
    const size_t per_slot = (size_t)K * (n * d.trec_stride + n / 8 + cells * (4 + 4 + 8 + 8 + 8) +
                                         chunks * (IMOMD_CHUNK * sizeof(VRec) + 8)) +
                            (size_t)d.arena_entries * 4 + chunks * 8 + (size_t)d.log_entries * 8 +
                            (d.pseudo ? (size_t)K * n * 4 : 0);
    // resident planner: one slot per query.  Streaming planner (max_resident_queries < n_queries):
    // queries share `slots` sets of state through a work queue.  -1 = as many as run concurrently
    // (two CTAs per SM) and fit in 85 % of the free device memory.
    int slots = n_queries;
    if (params->max_resident_queries > 0) slots = std::min(n_queries, params->max_resident_queries);
    else if (params->max_resident_queries < 0)
    {
        size_t free_b = 0, total_b = 0;
        PCU(cudaMemGetInfo(&free_b, &total_b));
        size_t fit = (size_t)((double)free_b * 0.85) / std::max<size_t>(per_slot, 1);
        slots = (int)std::min<size_t>({(size_t)n_queries, (size_t)2 * g->sm_count, std::max<size_t>(fit, 1)});
    }
    d.slots = slots;
    d.streaming = slots < n_queries ? 1 : 0;
    const size_t sk = (size_t)slots * K;
    PCU(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    PCU(cudaEventCreate(&p->ev0));
    PCU(cudaEventCreate(&p->ev1));
    PCU(dmalloc(&d.query, (size_t)n_queries));
    PCU(dmalloc(&d.trec, sk * n * d.trec_stride));
    PCU(dmalloc(&d.cell_head, sk * cells));
    PCU(dmalloc(&d.cell_cnt, sk * cells));
    PCU(dmalloc(&d.rec, sk * chunks * IMOMD_CHUNK));
    PCU(dmalloc(&d.chunk_next, sk * chunks));
    PCU(dmalloc(&d.free_stack, sk * chunks));
    PCU(dmalloc(&d.lbd, sk * cells));
    PCU(dmalloc(&d.ubd, sk * cells));
    // streaming planners: a dirty bitmap of every tree's records behind its occ lists, and the
    // records all-ones ONCE -- from here on a slot is recycled by rewriting the marked records
    d.dirty_words = d.streaming ? (uint32_t)((n + 31) / 32) : 0u;
    PCU(dmalloc(&d.occ, sk * (2 * cells + d.dirty_words)));
    if (d.streaming)
    {
        PCU(cudaMemsetAsync(d.occ, 0, sk * (2 * cells + d.dirty_words) * sizeof(uint32_t), p->stream));
        PCU(cudaMemsetAsync(d.trec, 0xFF, sk * n * d.trec_stride, p->stream));
    }
    PCU(dmalloc(&d.arena, (size_t)slots * d.arena_entries));
    PCU(dmalloc(&d.clist, (size_t)slots * 2 * chunks));
    PCU(dmalloc(&d.log, (size_t)slots * 2 * d.log_entries));
    PCU(dmalloc(&d.reports, (size_t)n_queries));
    PCU(cudaMemset(d.reports, 0, (size_t)n_queries * sizeof(ReportDev)));
    d.visit_log = nullptr;
    d.attach_list = nullptr;
    if (d.pseudo) PCU(dmalloc(&d.visit_log, sk * n));
    if (d.pseudo) PCU(dmalloc(&d.attach_list, n));
    if (d.streaming)
    {
        PCU(dmalloc(&p->query_init, (size_t)n_queries));
        PCU(dmalloc(&p->queue, 2));
        d.query_init = p->query_init;
        d.queue = p->queue;
    }
    // exact-value service: mailboxes in mapped pinned host memory (IMOMD_NO_HOST_CERTIFY=1 runs
    // without it: undecidable comparisons then keep the device's choice and are counted)
    if (!getenv("IMOMD_NO_HOST_CERTIFY"))
    {
        PCU(cudaHostAlloc((void**)&p->box_host, sizeof(TieBox) * kTieBoxes, cudaHostAllocMapped));
        std::memset(p->box_host, 0, sizeof(TieBox) * kTieBoxes);
        PCU(cudaHostGetDevicePointer((void**)&d.tiebox, p->box_host, 0));
    }
    d.words = nullptr;
    d.words_avail = 0;
    // initial headers
    p->dest_h.assign(dest, dest + (size_t)n_queries * K);
    p->prob_h.assign(prob, prob + (size_t)n_queries * K * K);
    PCU(dmalloc(&p->dest_d, (size_t)n_queries * K));
    PCU(dmalloc(&p->prob_d, (size_t)n_queries * K * K));
    p->smem_bytes = ((sizeof(Shared) + 15) / 16) * 16 + std::min<size_t>(cells, kCtaLbCap) * sizeof(double) +
                    2 * kLevelCap * sizeof(uint32_t) + (size_t)K * K * 32 +
                    ((IMOMD_QUERY_HOT_BYTES + 15) / 16) * 16 + (size_t)K * sizeof(TreeRef) +
                    (size_t)g->d.width * K * sizeof(double);
    {
        cudaError_t e = cudaSuccess;
        auto set = [&](const void* f) {
            cudaError_t r = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)p->smem_bytes);
            if (r != cudaSuccess) e = r;
        };
#define IMOMD_SET_ALL(W) set((const void*)k_expand<256, 1, W>); set((const void*)k_expand<256, 2, W>); set((const void*)k_attach<W>);
        IMOMD_SET_ALL(4) IMOMD_SET_ALL(8) IMOMD_SET_ALL(16)
#undef IMOMD_SET_ALL
        // Single-CTA kernels (the merge's k_attach, the one-query-per-SM latency build of k_expand)
        // want the rest of the SM's unified L1 / shared memory as L1: their chains of dependent
        // record loads hit it or wait for the L2.  Carveout = what one CTA needs, in percent.
        if (!getenv("IMOMD_NO_CARVEOUT"))
        {
            int smem_sm = 0;
            cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, g->device);
            const int pct = smem_sm > 0 ? (int)std::min<size_t>(100, (100 * (p->smem_bytes + 2048) + smem_sm - 1) / smem_sm) : 50;
            auto carve = [&](const void* f) {
                cudaError_t r = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
                if (r != cudaSuccess) e = r;
            };
            carve((const void*)k_attach<4>); carve((const void*)k_attach<8>); carve((const void*)k_attach<16>);
            carve((const void*)k_expand<256, 1, 4>); carve((const void*)k_expand<256, 1, 8>); carve((const void*)k_expand<256, 1, 16>);
        }
        PCU(e);
    }
    PCU(cudaFuncSetAttribute(k_nearest, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)p->smem_bytes));
    int rc = planner_fill_initial(p);
    if (rc != IMOMD_OK)
    {
        imomd_planner_destroy(p);
        return rc;
    }
    if (!d.streaming)
    {
        size_t total = sk * cells;
        k_lower_bounds<<<(unsigned)((total + 255) / 256), 256, 0, p->stream>>>(d);
        PCU(cudaGetLastError());
    }
    PCU(wait_stream(p));
#undef PCU
    *out = p;
    return IMOMD_OK;
}

int imomd_planner_destroy(imomd_planner* p)
{
    if (!p) return IMOMD_OK;
    cudaSetDevice(p->g->device);
    if (p->stream) wait_stream(p);
    PlannerDev& d = p->d;
    cudaFree(d.query);
    cudaFree(d.trec);
    cudaFree(d.cell_head);
    cudaFree(d.cell_cnt);
    cudaFree(d.rec);
    cudaFree(d.chunk_next);
    cudaFree(d.free_stack);
    cudaFree(d.lbd);
    cudaFree(d.ubd);
    cudaFree(d.occ);
    cudaFree(d.arena);
    cudaFree(d.clist);
    cudaFree(d.log);
    cudaFree(d.reports);
    cudaFree(d.visit_log);
    cudaFree(d.attach_list);
    cudaFree(p->query_init);
    cudaFree(p->queue);
    cudaFree(p->dest_d);
    cudaFree(p->prob_d);
    cudaFree(d.events);
    cudaFree(d.clist_tree);
    if (p->box_host) cudaFreeHost(p->box_host);
    cudaFree(p->words);
    cudaFree(p->nn_scratch);
    cudaFree(p->nn_stats);
    cudaFree(p->walk_buf);
    cudaFree(p->gather_u32);
    cudaFree(p->gather_f64);
    p->nn_pool.release();
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return IMOMD_OK;
}

int imomd_planner_set_queries(imomd_planner* p, const uint32_t* dest, const double* prob)
{
    if (!p || !dest || !prob) return fail(IMOMD_ERR_ARG, "null argument");
    const int Q = p->d.Q, K = p->d.K;
    for (size_t i = 0; i < (size_t)Q * K; ++i)
    {
        if (dest[i] >= p->d.g.n) return fail(IMOMD_ERR_ARG, "destination %u out of range", dest[i]);
    }
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));      // the previous upload may still be reading the host copies
    p->dest_h.assign(dest, dest + (size_t)Q * K);
    p->prob_h.assign(prob, prob + (size_t)Q * K * K);
    int rc = planner_fill_initial(p);
    if (rc != IMOMD_OK) return rc;
    if (!p->d.streaming)
    {
        size_t total = (size_t)Q * K * p->d.g.G * p->d.g.G;
        k_lower_bounds<<<(unsigned)((total + 255) / 256), 256, 0, p->stream>>>(p->d);
        CU(cudaGetLastError());
    }
    CU(wait_stream(p));
    return IMOMD_OK;
}

int imomd_attach_list(imomd_planner* p, int query, int tree, const uint32_t* vertices, uint32_t count,
                      uint32_t* parents_out, imomd_query_report* report)
{
    if (!p || query < 0 || query >= p->d.Q || tree < 0 || tree >= p->d.K || !vertices)
        return fail(IMOMD_ERR_ARG, "bad planner/query/tree/list");
    if (p->d.streaming) return fail(IMOMD_ERR_ARG, "%s: the per-query trees of a streaming planner are not resident", __func__);
    if (!p->d.attach_list) return fail(IMOMD_ERR_ARG, "imomd_attach needs a pseudo-mode planner");
    if (count == 0) return IMOMD_OK;
    if (count > p->d.g.n) return fail(IMOMD_ERR_ARG, "attach list longer than the graph");
    for (uint32_t i = 0; i < count; ++i)
        if (vertices[i] >= p->d.g.n) return fail(IMOMD_ERR_ARG, "attach list: vertex id out of range");
    CU(cudaSetDevice(p->g->device));
    CU(cudaMemcpyAsync(p->d.attach_list, vertices, (size_t)count * 4, cudaMemcpyHostToDevice, p->stream));
    // The whole CTA: a merge re-parents large subtrees, so its time goes to the cost propagation and
    // the examination of the updated vertices' rewire calls, which use every thread (measured on the
    // bug-trap merge, profiles/README.md: 136 ms with 256 threads, 148 / 155 / 178 ms with 128 / 64 /
    // 32); the second warp prefetches what the list's next entries touch (warm_attach).
    // IMOMD_ATTACH_THREADS overrides (development aid).
    static const int attach_threads = [] {
        const char* e = getenv("IMOMD_ATTACH_THREADS");
        int v = e ? atoi(e) : kThreads;
        return (v >= 32 && v <= kThreads && v % 32 == 0) ? v : kThreads;
    }();
    const int w = p->d.g.width;
    if (w == 4) k_attach<4><<<1, attach_threads, p->smem_bytes, p->stream>>>(p->d, query, tree, count);
    else if (w == 8) k_attach<8><<<1, attach_threads, p->smem_bytes, p->stream>>>(p->d, query, tree, count);
    else k_attach<16><<<1, attach_threads, p->smem_bytes, p->stream>>>(p->d, query, tree, count);
    CU(cudaGetLastError());
    // (a device->host copy into pageable memory blocks the host until the stream reaches it: the
    // kernel may be waiting for the exact-value service, so the wait -- which runs the service --
    // always comes BEFORE the copies)
    CU(wait_stream(p));
    if (report)
        CU(cudaMemcpyAsync(report, p->d.reports + query, sizeof(ReportDev), cudaMemcpyDeviceToHost,
                           p->stream));
    if (parents_out)
        CU(cudaMemcpyAsync(parents_out, p->d.attach_list, (size_t)count * 4, cudaMemcpyDeviceToHost,
                           p->stream));
    CU(wait_stream(p));
    return IMOMD_OK;
}

int imomd_attach(imomd_planner* p, int query, int tree, uint32_t vertex, uint32_t* parent_out,
                 imomd_query_report* report)
{
    return imomd_attach_list(p, query, tree, &vertex, 1, parent_out, report);
}

int imomd_get_visit_order(imomd_planner* p, int query, int tree, uint32_t* out, uint32_t cap,
                          uint32_t* count)
{
    if (!p || query < 0 || query >= p->d.Q || tree < 0 || tree >= p->d.K || !count)
        return fail(IMOMD_ERR_ARG, "bad planner/query/tree");
    if (p->d.streaming) return fail(IMOMD_ERR_ARG, "%s: the per-query trees of a streaming planner are not resident", __func__);
    if (!p->d.visit_log) return fail(IMOMD_ERR_ARG, "the visit order is only recorded in pseudo mode");
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));
    QueryDev h;
    CU(cudaMemcpy(&h, p->d.query + query, IMOMD_QUERY_HOT_BYTES, cudaMemcpyDeviceToHost));
    uint32_t n = (uint32_t)h.tree_size[tree];
    *count = n;
    uint32_t take = std::min(n, cap);
    const size_t qt = (size_t)query * p->d.K + tree;
    if (take && out)
        CU(cudaMemcpy(out, p->d.visit_log + qt * p->d.g.n, (size_t)take * 4, cudaMemcpyDeviceToHost));
    return IMOMD_OK;
}

// host-side edits of a query's state the pseudo-tree merge needs (mergePseudoTrees_ :797-816,
// mergeTwoTree_ :905-915): distance / connection-node entry, is_done flag, "set non-empty" flag
int imomd_set_distance(imomd_planner* p, int query, int i, int j, double d, uint32_t conn_node)
{
    if (!p || query < 0 || query >= p->d.Q || i < 0 || j < 0 || i >= p->d.K || j >= p->d.K)
        return fail(IMOMD_ERR_ARG, "bad argument");
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));
    QueryDev* qd = p->d.query + query;
    const int K = p->d.K;
    int32_t one = 1;
    CU(cudaMemcpy(&qd->D[i * K + j], &d, 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(&qd->D[j * K + i], &d, 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(&qd->CN[i * K + j], &conn_node, 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(&qd->CN[j * K + i], &conn_node, 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(&qd->dm_updated, &one, 4, cudaMemcpyHostToDevice));
    return IMOMD_OK;
}

int imomd_set_tree_done(imomd_planner* p, int query, int tree, int done)
{
    if (!p || query < 0 || query >= p->d.Q || tree < 0 || tree >= p->d.K)
        return fail(IMOMD_ERR_ARG, "bad argument");
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));
    int32_t v = done ? 1 : 0;
    CU(cudaMemcpy(&p->d.query[query].is_done[tree], &v, 4, cudaMemcpyHostToDevice));
    return IMOMD_OK;
}

int imomd_set_contact(imomd_planner* p, int query, int set_idx)
{
    if (!p || query < 0 || query >= p->d.Q || set_idx < 0 || set_idx >= p->d.K * (p->d.K - 1) / 2)
        return fail(IMOMD_ERR_ARG, "bad argument");
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));
    uint8_t one = 1;
    CU(cudaMemcpy(&p->d.query[query].contact[set_idx], &one, 1, cudaMemcpyHostToDevice));
    return IMOMD_OK;
}

int imomd_get_tie_info(imomd_planner* p, int query, double* out)
{
    if (!p || query < 0 || query >= p->d.Q || !out) return fail(IMOMD_ERR_ARG, "bad argument");
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));
    CU(cudaMemcpy(out, p->d.query[query].tie_info, 8 * sizeof(double), cudaMemcpyDeviceToHost));
    return IMOMD_OK;
}

int imomd_get_counters(imomd_planner* p, uint64_t* out)
{
    if (!p || !out) return fail(IMOMD_ERR_ARG, "null argument");
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));
    CU(cudaMemcpy2D(out, 8 * sizeof(uint64_t), p->d.query[0].stat, sizeof(QueryDev),
                    8 * sizeof(uint64_t), (size_t)p->d.Q, cudaMemcpyDeviceToHost));
    return IMOMD_OK;
}

int imomd_planner_clear_words(imomd_planner* p)
{
    if (!p) return fail(IMOMD_ERR_ARG, "null planner");
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));
    p->words_fed = 0;
    p->d.words_avail = 0;
    return IMOMD_OK;
}

int imomd_planner_reset(imomd_planner* p)
{
    if (!p) return fail(IMOMD_ERR_ARG, "null planner");
    CU(cudaSetDevice(p->g->device));
    int rc = planner_fill_initial(p);
    if (rc != IMOMD_OK) return rc;
    CU(wait_stream(p));
    return IMOMD_OK;
}

int imomd_planner_feed_words(imomd_planner* p, const uint32_t* words, size_t n)
{
    if (!p || (!words && n)) return fail(IMOMD_ERR_ARG, "null argument");
    CU(cudaSetDevice(p->g->device));
    if (p->words_fed + n > p->words_cap)
    {
        size_t cap = std::max<size_t>(p->words_cap * 2, p->words_fed + n);
        cap = std::max<size_t>(cap, 1 << 16);
        uint32_t* nw = nullptr;
        CU(dmalloc(&nw, cap));
        CU(wait_stream(p));
        if (p->words_fed)
            CU(cudaMemcpy(nw, p->words, p->words_fed * 4, cudaMemcpyDeviceToDevice));
        cudaFree(p->words);
        p->words = nw;
        p->words_cap = cap;
    }
    CU(cudaMemcpyAsync(p->words + p->words_fed, words, n * 4, cudaMemcpyHostToDevice, p->stream));
    p->words_fed += n;
    p->d.words = p->words;
    p->d.words_avail = p->words_fed;
    return IMOMD_OK;
}

int imomd_planner_words_available(const imomd_planner* p, uint64_t* total_fed)
{
    if (!p || !total_fed) return fail(IMOMD_ERR_ARG, "null argument");
    *total_fed = p->words_fed;
    return IMOMD_OK;
}

int imomd_expand_async(imomd_planner* p, int iters)
{
    if (!p) return fail(IMOMD_ERR_ARG, "null planner");
    if (iters < 0) return fail(IMOMD_ERR_ARG, "negative iteration count");
    if (p->batch_ran && !p->d.streaming)
        return fail(IMOMD_ERR_ARG, "batched steps ran on this planner: reset it before the exact mode (the batched "
                                   "mode does not maintain the child rows the exact rewiring walks)");
    CU(cudaSetDevice(p->g->device));
    CU(cudaEventRecord(p->ev0, p->stream));
    {
        // Register budget by the number of queries: up to one query per SM the latency build
        // (255 registers, no spills); beyond that two CTAs per SM, whose dependent-load
        // latencies overlap.  Measured on B200 (round 1): four or eight smaller CTAs per SM
        // (128 threads and/or 64 registers) are 5-20 % slower than two -- the CTA-wide
        // services lose more than the extra residency returns -- so larger query counts simply
        // run in waves.  Adjacency width by the graph.
        // A streaming planner launches one CTA per slot; its CTAs pull queries from the queue.
        const int ctas = p->d.streaming ? p->d.slots : p->d.Q;
        if (p->d.streaming) CU(cudaMemsetAsync(p->queue, 0, 2 * sizeof(uint32_t), p->stream));
        const bool two_per_sm = ctas > p->g->sm_count;
        const int w = p->d.g.width;
        void (*kern)(PlannerDev, int) = nullptr;
        if (w == 4) kern = two_per_sm ? k_expand<256, 2, 4> : k_expand<256, 1, 4>;
        else if (w == 8) kern = two_per_sm ? k_expand<256, 2, 8> : k_expand<256, 1, 8>;
        else kern = two_per_sm ? k_expand<256, 2, 16> : k_expand<256, 1, 16>;
        kern<<<ctas, kThreads, p->smem_bytes, p->stream>>>(p->d, iters);
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(p->ev1, p->stream));
    p->pending = true;
    return IMOMD_OK;
}

// Batched mode: `steps` steps of B expansions per tree, the K trees of a query expanded
// concurrently (one warp each), cross-tree effects applied after every step.
int imomd_expand_batch_async(imomd_planner* p, int steps, int B)
{
    if (!p) return fail(IMOMD_ERR_ARG, "null planner");
    if (steps < 0 || B < 1 || B > 4096) return fail(IMOMD_ERR_ARG, "steps must be >= 0 and B in [1, 4096]");
    if (p->d.pseudo) return fail(IMOMD_ERR_ARG, "the batched mode does not cover pseudo mode (the merge needs the exact visit order)");
    CU(cudaSetDevice(p->g->device));
    PlannerDev& d = p->d;
    constexpr int kBatchThreads = 256;
    if (!p->batch_ready)
    {
        CU(wait_stream(p));
        const size_t sk = (size_t)d.slots * d.K;
        d.ev_cap = 16384;
        CU(dmalloc(&d.events, sk * d.ev_cap));
        CU(dmalloc(&d.clist_tree, sk * 2 * (size_t)d.chunks));
        const size_t cells_unused = 0;
        (void)cells_unused;
        p->smem_batch = (kBatchThreads / 32) * (((sizeof(Shared) + 15) / 16) * 16) + ((sizeof(BatchShared) + 15) / 16) * 16 +
                        (size_t)(kBatchThreads / 32) * kGroupLbCap * sizeof(double) + (size_t)d.K * d.K * 32 +
                        ((IMOMD_QUERY_HOT_BYTES + 15) / 16) * 16 + (size_t)d.K * sizeof(TreeRef) +
                        (size_t)(kBatchThreads / 32) * d.g.width * d.K * sizeof(double) + 64;
        cudaError_t e = cudaSuccess;
        auto set = [&](const void* f) {
            cudaError_t r = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_batch);
            if (r != cudaSuccess) e = r;
        };
        set((const void*)k_expand_batch<kBatchThreads, 2, 4>);
        set((const void*)k_expand_batch<kBatchThreads, 2, 8>);
        set((const void*)k_expand_batch<kBatchThreads, 2, 16>);
        CU(e);
        p->batch_ready = true;
    }
    p->batch_ran = true;
    CU(cudaEventRecord(p->ev0, p->stream));
    const int ctas = d.streaming ? d.slots : d.Q;
    if (d.streaming) CU(cudaMemsetAsync(p->queue, 0, 2 * sizeof(uint32_t), p->stream));
    const int w = d.g.width;
    if (w == 4) k_expand_batch<kBatchThreads, 2, 4><<<ctas, kBatchThreads, p->smem_batch, p->stream>>>(d, steps, B);
    else if (w == 8) k_expand_batch<kBatchThreads, 2, 8><<<ctas, kBatchThreads, p->smem_batch, p->stream>>>(d, steps, B);
    else k_expand_batch<kBatchThreads, 2, 16><<<ctas, kBatchThreads, p->smem_batch, p->stream>>>(d, steps, B);
    CU(cudaGetLastError());
    CU(cudaEventRecord(p->ev1, p->stream));
    p->pending = true;
    return IMOMD_OK;
}

int imomd_expand_batch(imomd_planner* p, int steps, int B, imomd_query_report* reports)
{
    int rc = imomd_expand_batch_async(p, steps, B);
    if (rc != IMOMD_OK) return rc;
    return imomd_sync(p, reports);
}

// max_time of the reference's loop (:258-262): the following imomd_expand calls stop after the
// first pass that ends more than `seconds` after the launch started (reports say how many passes
// ran); 0 = no budget
int imomd_set_time_budget(imomd_planner* p, double seconds)
{
    if (!p || !(seconds >= 0.0)) return fail(IMOMD_ERR_ARG, "bad argument");
    p->d.budget_ns = seconds > 0.0 ? std::max<uint64_t>(1, (uint64_t)(seconds * 1e9)) : 0;
    return IMOMD_OK;
}

int imomd_sync(imomd_planner* p, imomd_query_report* reports)
{
    if (!p) return fail(IMOMD_ERR_ARG, "null planner");
    CU(cudaSetDevice(p->g->device));
    static_assert(sizeof(imomd_query_report) == sizeof(ReportDev), "report layouts must agree");
    CU(wait_stream(p));      // first: the kernel may be asking the exact-value service
    if (reports)
        CU(cudaMemcpy(reports, p->d.reports, (size_t)p->d.Q * sizeof(ReportDev), cudaMemcpyDeviceToHost));
    if (p->pending)
    {
        CU(cudaEventElapsedTime(&p->last_ms, p->ev0, p->ev1));
        p->pending = false;
    }
    return IMOMD_OK;
}

int imomd_expand(imomd_planner* p, int iters, imomd_query_report* reports)
{
    int rc = imomd_expand_async(p, iters);
    if (rc != IMOMD_OK) return rc;
    return imomd_sync(p, reports);
}

int imomd_last_kernel_ms(imomd_planner* p, float* ms)
{
    if (!p || !ms) return fail(IMOMD_ERR_ARG, "null argument");
    *ms = p->last_ms;
    return IMOMD_OK;
}

int imomd_get_state(imomd_planner* p, int query, double* D, uint32_t* conn_node, double* prob,
                    uint32_t* mh_id, double* mh_val, int32_t* is_done, int32_t* ds_parent)
{
    if (!p || query < 0 || query >= p->d.Q) return fail(IMOMD_ERR_ARG, "bad planner/query");
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));
    QueryDev h;
    CU(cudaMemcpy(&h, p->d.query + query, sizeof(h), cudaMemcpyDeviceToHost));
    int K = p->d.K;
    for (int i = 0; i < K * K; ++i)
    {
        if (D) D[i] = h.D[i];
        if (conn_node) conn_node[i] = h.CN[i];
        if (prob) prob[i] = h.P[i];
        if (mh_id) mh_id[i] = h.MHi[i];
        if (mh_val) mh_val[i] = h.MHv[i];
    }
    for (int k = 0; k < K; ++k)
    {
        if (is_done) is_done[k] = h.is_done[k];
        if (ds_parent) ds_parent[k] = h.ds_parent[k];
    }
    return IMOMD_OK;
}

// K x K matrices of queries [first, first + count) in ONE device->host copy: each output holds
// count * K * K entries (count * K for is_done), any of them may be NULL
int imomd_get_states(imomd_planner* p, int first, int count, double* D, uint32_t* conn_node, double* prob,
                     int32_t* is_done)
{
    if (!p || first < 0 || count < 0 || first + count > p->d.Q) return fail(IMOMD_ERR_ARG, "bad planner/query range");
    if (count == 0) return IMOMD_OK;
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));
    std::vector<QueryDev> h((size_t)count);
    CU(cudaMemcpy(h.data(), p->d.query + first, (size_t)count * sizeof(QueryDev), cudaMemcpyDeviceToHost));
    const int K = p->d.K;
    for (int q = 0; q < count; ++q)
    {
        const size_t o = (size_t)q * K * K;
        for (int i = 0; i < K * K; ++i)
        {
            if (D) D[o + i] = h[q].D[i];
            if (conn_node) conn_node[o + i] = h[q].CN[i];
            if (prob) prob[o + i] = h[q].P[i];
        }
        if (is_done)
            for (int k = 0; k < K; ++k) is_done[(size_t)q * K + k] = h[q].is_done[k];
    }
    return IMOMD_OK;
}

// TREEHASH of SURVEY.md Appendix A (trees in id order, visited vertices ascending, FNV fold of
// vertex / parent / cost bits) and an order-free 64-bit checksum of the same entries, computed on
// the device.  Streaming planners: the values the query left when it finished; resident planners:
// computed now from the trees in HBM.
int imomd_get_treehash(imomd_planner* p, int query, uint64_t* tree_hash, uint64_t* tree_sum)
{
    if (!p || query < 0 || query >= p->d.Q) return fail(IMOMD_ERR_ARG, "bad planner/query");
    CU(cudaSetDevice(p->g->device));
    if (!p->d.streaming)
    {
        CU(cudaFuncSetAttribute(k_tree_hash, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Shared) + 16));
        k_tree_hash<<<1, kThreads, sizeof(Shared) + 16, p->stream>>>(p->d, query);
        CU(cudaGetLastError());
    }
    CU(wait_stream(p));
    uint64_t v[2];
    CU(cudaMemcpy(v, &p->d.query[query].tree_hash, sizeof(v), cudaMemcpyDeviceToHost));   // tree_hash, tree_sum
    if (tree_hash) *tree_hash = v[0];
    if (tree_sum) *tree_sum = v[1];
    return IMOMD_OK;
}

// how the planner is laid out: queries, sets of per-query device state, 1 when queries stream
// through the slots, bytes of device memory per slot, requests the exact-value service answered
int imomd_planner_info(imomd_planner* p, int32_t* n_queries, int32_t* slots, int32_t* streaming,
                       uint64_t* certified_requests)
{
    if (!p) return fail(IMOMD_ERR_ARG, "null planner");
    if (n_queries) *n_queries = p->d.Q;
    if (slots) *slots = p->d.slots;
    if (streaming) *streaming = p->d.streaming;
    if (certified_requests) *certified_requests = p->certified_requests;
    return IMOMD_OK;
}

int imomd_get_flags(imomd_planner* p, int query, int32_t* connected, int32_t* dm_updated)
{
    if (!p || query < 0 || query >= p->d.Q) return fail(IMOMD_ERR_ARG, "bad planner/query");
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));
    int32_t v[2];
    QueryDev* qd = p->d.query + query;
    CU(cudaMemcpy(v, &qd->connected, sizeof(v), cudaMemcpyDeviceToHost));   // connected, dm_updated
    if (connected) *connected = v[0];
    if (dm_updated) *dm_updated = v[1];
    return IMOMD_OK;
}

int imomd_set_profiling(imomd_planner* p, int on)
{
    if (!p) return fail(IMOMD_ERR_ARG, "bad argument");
    p->d.profile = on ? 1u : 0u;
    return IMOMD_OK;
}

int imomd_get_profile(imomd_planner* p, int query, uint64_t* cycles, uint64_t* counts)
{
    if (!p || query < 0 || query >= p->d.Q || !cycles || !counts)
        return fail(IMOMD_ERR_ARG, "bad argument");
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));
    QueryDev* qd = p->d.query + query;
    CU(cudaMemcpy(cycles, qd->prof_cycles, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(counts, qd->prof_count, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (getenv("IMOMD_PHASE_PROFILE"))
    {
        uint64_t pc[16];
        CU(cudaMemcpy(pc, qd->phase_cycles, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        const char* names[8] = {"next_tree", "after_nn", "jps", "insert", "finish", "rewire", "last_conn", "-"};
        for (int i = 0; i < 7; ++i)
            fprintf(stderr, "  phase %-10s steps %10llu  cycles/step %8.0f\n", names[i],
                    (unsigned long long)pc[8 + i], pc[8 + i] ? (double)pc[i] / (double)pc[8 + i] : 0.0);
    }
    return IMOMD_OK;
}

int imomd_get_phase_cycles(imomd_planner* p, int query, uint64_t* cycles, uint64_t* counts)
{
    if (!p || query < 0 || query >= p->d.Q || !cycles || !counts)
        return fail(IMOMD_ERR_ARG, "bad argument");
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));
    QueryDev* qd = p->d.query + query;
    CU(cudaMemcpy(cycles, qd->phase_cycles, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(counts, qd->phase_count, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return IMOMD_OK;
}

int imomd_clear_dm_updated(imomd_planner* p, int query)
{
    if (!p || query < 0 || query >= p->d.Q) return fail(IMOMD_ERR_ARG, "bad planner/query");
    CU(cudaSetDevice(p->g->device));
    int32_t zero = 0;
    QueryDev* qd = p->d.query + query;
    CU(cudaMemcpyAsync(&qd->dm_updated, &zero, sizeof(zero), cudaMemcpyHostToDevice, p->stream));
    CU(wait_stream(p));
    return IMOMD_OK;
}

int imomd_get_tree(imomd_planner* p, int query, int tree, uint32_t* parent, double* cost)
{
    if (!p || query < 0 || query >= p->d.Q || tree < 0 || tree >= p->d.K || !parent)
        return fail(IMOMD_ERR_ARG, "bad planner/query/tree");
    if (p->d.streaming) return fail(IMOMD_ERR_ARG, "%s: the per-query trees of a streaming planner are not resident", __func__);
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));
    size_t n = p->d.g.n;
    size_t qt = (size_t)query * p->d.K + tree;
    const unsigned char* base = p->d.trec + qt * n * p->d.trec_stride;
    CU(cudaMemcpy2D(parent, 4, base, p->d.trec_stride, 4, n, cudaMemcpyDeviceToHost));
    if (cost)
    {
        CU(cudaMemcpy2D(cost, 8, base + 8, p->d.trec_stride, 8, n, cudaMemcpyDeviceToHost));
        for (size_t v = 0; v < n; ++v)
        {
            if (parent[v] == IMOMD_NONE) cost[v] = INFINITY;
        }
    }
    return IMOMD_OK;
}

static int ensure_gather(imomd_planner* p)
{
    if (!p->gather_u32) CU(dmalloc(&p->gather_u32, (size_t)p->d.g.n + 1));
    if (!p->gather_f64) CU(dmalloc(&p->gather_f64, (size_t)p->d.g.n + 1));
    return IMOMD_OK;
}

int imomd_get_frontier(imomd_planner* p, int query, int tree, uint32_t* out, uint32_t cap,
                       uint32_t* count)
{
    if (!p || query < 0 || query >= p->d.Q || tree < 0 || tree >= p->d.K || !count)
        return fail(IMOMD_ERR_ARG, "bad planner/query/tree");
    if (p->d.streaming) return fail(IMOMD_ERR_ARG, "%s: the per-query trees of a streaming planner are not resident", __func__);
    CU(cudaSetDevice(p->g->device));
    if (int rc = ensure_gather(p)) return rc;
    const uint32_t n = p->d.g.n;
    size_t qt = (size_t)query * p->d.K + tree;
    CU(cudaMemsetAsync(p->gather_u32, 0, 4, p->stream));
    k_frontier_list<<<(n + 255) / 256, 256, 0, p->stream>>>(p->d, qt, p->gather_u32);
    CU(cudaGetLastError());
    uint32_t c = 0;
    CU(wait_stream(p));
    CU(cudaMemcpyAsync(&c, p->gather_u32, 4, cudaMemcpyDeviceToHost, p->stream));
    CU(wait_stream(p));
    *count = c;
    uint32_t take = std::min(c, cap);
    if (take && out)
    {
        CU(cudaMemcpy(out, p->gather_u32 + 1, (size_t)take * 4, cudaMemcpyDeviceToHost));
        std::sort(out, out + take);   // ascending ids, as the dense scan used to return them
    }
    return IMOMD_OK;
}

int imomd_gather_tree(imomd_planner* p, int query, int tree, const uint32_t* vertices, uint32_t count,
                      uint32_t* parent, double* cost)
{
    if (!p || query < 0 || query >= p->d.Q || tree < 0 || tree >= p->d.K || (count && !vertices))
        return fail(IMOMD_ERR_ARG, "bad planner/query/tree/list");
    if (count > p->d.g.n) return fail(IMOMD_ERR_ARG, "gather list longer than the graph");
    if (count == 0) return IMOMD_OK;
    for (uint32_t i = 0; i < count; ++i)
        if (vertices[i] >= p->d.g.n) return fail(IMOMD_ERR_ARG, "gather list: vertex id out of range");
    if (p->d.streaming) return fail(IMOMD_ERR_ARG, "%s: the per-query trees of a streaming planner are not resident", __func__);
    CU(cudaSetDevice(p->g->device));
    if (int rc = ensure_gather(p)) return rc;
    size_t qt = (size_t)query * p->d.K + tree;
    CU(cudaMemcpyAsync(p->gather_u32, vertices, (size_t)count * 4, cudaMemcpyHostToDevice, p->stream));
    k_gather_tree<<<(count + 255) / 256, 256, 0, p->stream>>>(p->d, qt, p->gather_u32, p->gather_f64, count);
    CU(cudaGetLastError());
    CU(wait_stream(p));
    if (parent)
        CU(cudaMemcpyAsync(parent, p->gather_u32, (size_t)count * 4, cudaMemcpyDeviceToHost, p->stream));
    if (cost)
        CU(cudaMemcpyAsync(cost, p->gather_f64, (size_t)count * 8, cudaMemcpyDeviceToHost, p->stream));
    CU(wait_stream(p));
    return IMOMD_OK;
}

int imomd_walk_to_root(imomd_planner* p, int query, int tree, uint32_t vertex, uint32_t* out,
                       uint32_t cap, uint32_t* len)
{
    if (!p || query < 0 || query >= p->d.Q || tree < 0 || tree >= p->d.K || !out || !len)
        return fail(IMOMD_ERR_ARG, "bad planner/query/tree");
    if (p->d.streaming) return fail(IMOMD_ERR_ARG, "%s: the per-query trees of a streaming planner are not resident", __func__);
    if (vertex >= p->d.g.n) return fail(IMOMD_ERR_ARG, "vertex out of range");
    CU(cudaSetDevice(p->g->device));
    if (p->walk_cap < cap + 1)
    {
        cudaFree(p->walk_buf);
        p->walk_buf = nullptr;
        CU(dmalloc(&p->walk_buf, (size_t)cap + 1));
        p->walk_cap = cap + 1;
    }
    size_t n = p->d.g.n;
    size_t qt = (size_t)query * p->d.K + tree;
    uint32_t root = p->dest_h[(size_t)query * p->d.K + tree];
    k_walk<<<1, 32, 0, p->stream>>>(p->d.trec + qt * n * p->d.trec_stride, p->d.trec_stride, vertex,
                                    root, p->walk_buf + 1, cap, p->walk_buf);
    CU(cudaGetLastError());
    uint32_t l = 0;
    CU(wait_stream(p));
    CU(cudaMemcpyAsync(&l, p->walk_buf, 4, cudaMemcpyDeviceToHost, p->stream));
    CU(wait_stream(p));
    *len = l;
    uint32_t take = std::min(l, cap);
    if (take) CU(cudaMemcpy(out, p->walk_buf + 1, (size_t)take * 4, cudaMemcpyDeviceToHost));
    return IMOMD_OK;
}

int imomd_get_connection_log_from(imomd_planner* p, int query, uint32_t first, uint32_t* set_idx,
                                  uint32_t* vertex, uint32_t cap, uint32_t* count)
{
    if (!p || query < 0 || query >= p->d.Q || !count) return fail(IMOMD_ERR_ARG, "bad argument");
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));
    uint32_t total = 0;
    CU(cudaMemcpy(&total, &p->d.query[query].log_count, 4, cudaMemcpyDeviceToHost));
    *count = total;
    if (total > p->d.log_entries) return fail(IMOMD_ERR_ARG, "connection log overflow: raise log_entries");
    if (first >= total) return IMOMD_OK;
    uint32_t take = std::min(total - first, cap);
    if (take && set_idx && vertex)
    {
        std::vector<uint32_t> buf((size_t)take * 2);
        CU(cudaMemcpy(buf.data(), p->d.log + (size_t)query * 2 * p->d.log_entries + 2 * (size_t)first,
                      (size_t)take * 8, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < take; ++i)
        {
            set_idx[i] = buf[2 * i];
            vertex[i] = buf[2 * i + 1];
        }
    }
    return IMOMD_OK;
}

int imomd_clear_connection_log(imomd_planner* p, int query)
{
    if (!p || query < 0 || query >= p->d.Q) return fail(IMOMD_ERR_ARG, "bad planner/query");
    CU(cudaSetDevice(p->g->device));
    uint32_t zero = 0;
    CU(cudaMemcpyAsync(&p->d.query[query].log_count, &zero, 4, cudaMemcpyHostToDevice, p->stream));
    CU(wait_stream(p));
    return IMOMD_OK;
}

int imomd_get_connection_log(imomd_planner* p, int query, uint32_t* set_idx, uint32_t* vertex,
                             uint32_t cap, uint32_t* count)
{
    if (!p || query < 0 || query >= p->d.Q || !count) return fail(IMOMD_ERR_ARG, "bad argument");
    CU(cudaSetDevice(p->g->device));
    CU(wait_stream(p));
    QueryDev h;
    CU(cudaMemcpy(&h, p->d.query + query, sizeof(h), cudaMemcpyDeviceToHost));
    *count = h.log_count;
    uint32_t take = std::min(std::min(h.log_count, p->d.log_entries), cap);
    if (take && set_idx && vertex)
    {
        std::vector<uint32_t> buf((size_t)take * 2);
        CU(cudaMemcpy(buf.data(), p->d.log + (size_t)query * 2 * p->d.log_entries,
                      (size_t)take * 8, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < take; ++i)
        {
            set_idx[i] = buf[2 * i];
            vertex[i] = buf[2 * i + 1];
        }
    }
    return IMOMD_OK;
}

int imomd_nearest(imomd_planner* p, int n, const imomd_nn_query* q, uint32_t* idx, double* dist)
{
    if (!p || !q || !idx || !dist || n < 0) return fail(IMOMD_ERR_ARG, "bad argument");
    if (n == 0) return IMOMD_OK;
    for (int i = 0; i < n; ++i)
    {
        if (q[i].query < 0 || q[i].query >= p->d.Q || q[i].tree < 0 || q[i].tree >= p->d.K)
            return fail(IMOMD_ERR_ARG, "lookup %d: query/tree out of range", i);
    }
    if (p->d.streaming) return fail(IMOMD_ERR_ARG, "%s: the per-query trees of a streaming planner are not resident", __func__);
    CU(cudaSetDevice(p->g->device));
    int ctas = std::min(n, p->g->sm_count * 4);
    if (p->nn_ctas < ctas)
    {
        cudaFree(p->nn_scratch);
        p->nn_scratch = nullptr;
        CU(dmalloc(&p->nn_scratch, (size_t)ctas * 2 * p->d.chunks));
        p->nn_ctas = ctas;
    }
    if (!p->nn_stats) CU(dmalloc(&p->nn_stats, 2));
    imomd_nn_query* dq = nullptr;
    uint32_t* didx = nullptr;
    double* ddist = nullptr;
    CU(p->nn_pool.get(16, &dq, (size_t)n));     // grow-only scratch owned by the planner
    CU(p->nn_pool.get(17, &didx, (size_t)n));
    CU(p->nn_pool.get(18, &ddist, (size_t)n));
    CU(cudaMemcpyAsync(dq, q, (size_t)n * sizeof(imomd_nn_query), cudaMemcpyHostToDevice, p->stream));
    CU(cudaMemsetAsync(p->nn_stats, 0, 16, p->stream));
    k_nearest<<<ctas, kThreads, p->smem_bytes, p->stream>>>(p->d, n, dq, didx, ddist, p->nn_scratch,
                                                            p->nn_stats);
    CU(cudaGetLastError());
    CU(wait_stream(p));      // runs the exact-value service while k_nearest may be asking
    CU(cudaMemcpyAsync(idx, didx, (size_t)n * 4, cudaMemcpyDeviceToHost, p->stream));
    CU(cudaMemcpyAsync(dist, ddist, (size_t)n * 8, cudaMemcpyDeviceToHost, p->stream));
    CU(wait_stream(p));
    return IMOMD_OK;
}

int imomd_nearest_batch(imomd_planner* p, int n, const imomd_nn_query* q, uint32_t* idx, double* dist,
                        float* kernel_ms, uint64_t* bytes_streamed, uint32_t* n_refined)
{
    if (!p || !q || !idx || !dist || n < 0) return fail(IMOMD_ERR_ARG, "bad argument");
    if (n == 0) return IMOMD_OK;
    const int QK = p->d.Q * p->d.K;
    for (int i = 0; i < n; ++i)
    {
        if (q[i].query < 0 || q[i].query >= p->d.Q || q[i].tree < 0 || q[i].tree >= p->d.K)
            return fail(IMOMD_ERR_ARG, "lookup %d: query/tree out of range", i);
    }
    if (p->d.streaming) return fail(IMOMD_ERR_ARG, "%s: the per-query trees of a streaming planner are not resident", __func__);
    CU(cudaSetDevice(p->g->device));
    // tiles: lookups grouped by (query, tree) in their order of appearance (counting sort by tree),
    // at most kTileLookups per tile
    std::vector<uint32_t> first((size_t)QK + 1, 0u);
    for (int i = 0; i < n; ++i) ++first[(size_t)(q[i].query * p->d.K + q[i].tree) + 1];
    for (int t = 0; t < QK; ++t) first[(size_t)t + 1] += first[t];
    std::vector<int> order(n);
    {
        std::vector<uint32_t> at(first.begin(), first.end() - 1);
        for (int i = 0; i < n; ++i) order[at[(size_t)(q[i].query * p->d.K + q[i].tree)]++] = i;
    }
    std::vector<NnTile> tiles;
    tiles.reserve((size_t)n / kTileLookups + (size_t)std::min(n, QK));
    for (int t = 0; t < QK; ++t)
    {
        for (uint32_t i = first[t]; i < first[(size_t)t + 1];)
        {
            NnTile x{};
            x.qt = t;
            while (i < first[(size_t)t + 1] && x.n < kTileLookups) x.lookup[x.n++] = order[i++];
            tiles.push_back(x);
        }
    }
    NnPool& sc = p->nn_pool;
    const int n_tiles = (int)tiles.size();
    // tiles are sorted by tree: [tile_begin[qt], tile_begin[qt + 1]) are the tiles of tree qt
    std::vector<uint32_t> tile_begin((size_t)QK + 1, 0u);
    for (const NnTile& t : tiles) ++tile_begin[(size_t)t.qt + 1];
    for (int i = 0; i < QK; ++i) tile_begin[(size_t)i + 1] += tile_begin[i];
    uint32_t *lists = nullptr, *counts = nullptr, *didx = nullptr, *damb = nullptr, *tile_valid = nullptr,
             *seg_start = nullptr, *next_work = nullptr, *n_work_dev = nullptr, *dbegin = nullptr;
    double* ddist = nullptr;
    imomd_nn_query* dq = nullptr;
    NnTile* dtiles = nullptr;
    NnSeed* seed = nullptr;
    NnWork* work = nullptr;
    Best* part = nullptr;
    int* derr = nullptr;
    CU(sc.events());
    CU(sc.get(0, &lists, (size_t)QK * 2 * p->d.chunks));
    CU(sc.get(1, &counts, (size_t)QK));
    CU(sc.get(2, &didx, (size_t)n));
    CU(sc.get(3, &damb, (size_t)n));
    CU(sc.get(4, &ddist, (size_t)n));
    CU(sc.get(5, &dq, (size_t)n));
    CU(sc.get(6, &dtiles, (size_t)n_tiles));
    CU(sc.get(7, &dbegin, (size_t)QK + 1));
    CU(sc.get(8, &seed, (size_t)n_tiles));
    CU(sc.get(9, &tile_valid, (size_t)n_tiles));
    CU(sc.get(10, &seg_start, (size_t)n_tiles));
    CU(sc.get(11, &next_work, 2));
    n_work_dev = next_work + 1;
    CU(sc.get(12, &derr, 1));
    CU(cudaMemsetAsync(derr, 0, sizeof(int), p->stream));
    CU(cudaMemsetAsync(next_work, 0, 2 * sizeof(uint32_t), p->stream));
    CU(cudaMemsetAsync(counts, 0, (size_t)QK * sizeof(uint32_t), p->stream));
    CU(cudaMemcpyAsync(dq, q, (size_t)n * sizeof(imomd_nn_query), cudaMemcpyHostToDevice, p->stream));
    CU(cudaMemcpyAsync(dtiles, tiles.data(), tiles.size() * sizeof(NnTile), cudaMemcpyHostToDevice,
                       p->stream));
    CU(cudaMemcpyAsync(dbegin, tile_begin.data(), tile_begin.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                       p->stream));
    // preparation: chunk lists of the trees, seeds and segment ranges of the tiles in one launch
    CU(cudaEventRecord(sc.ev[0], p->stream));
    k_nn_prepare<<<QK, kPrepThreads, 0, p->stream>>>(p->d, dbegin, dtiles, dq, lists, counts, seed, tile_valid,
                                                     seg_start, n_work_dev);
    CU(cudaGetLastError());
    CU(cudaEventRecord(sc.ev[1], p->stream));
    uint32_t n_work = 0;
    CU(wait_stream(p));
    CU(cudaMemcpyAsync(&n_work, n_work_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, p->stream));
    CU(wait_stream(p));
    CU(sc.get(13, &work, (size_t)n_work));
    CU(sc.get(14, &part, (size_t)n_work * kConsumerWarps * kTileLookups));
    CU(cudaEventRecord(sc.ev[5], p->stream));
    k_nn_fill_work<<<(n_tiles + 3) / 4, 128, 0, p->stream>>>(p->d, n_tiles, dtiles, lists, counts, seg_start, work);
    CU(cudaGetLastError());
    // the tile kernel: persistent CTAs over the segments
    const size_t smem = sizeof(NnShared) + 128;
    CU(cudaFuncSetAttribute(k_nearest_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(cudaEventRecord(sc.ev[2], p->stream));
    if (n_work > 0)
    {
        const int grid = (int)std::min<size_t>(n_work, (size_t)p->g->sm_count * kTileCtasPerSm);
        k_nearest_tiles<<<grid, kTileThreads, smem, p->stream>>>(p->d, seed, work, n_work_dev, next_work,
                                                                 lists, part, derr);
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(sc.ev[3], p->stream));
    k_nn_merge<<<(n_tiles * kTileLookups + 255) / 256, 256, 0, p->stream>>>(p->d, n_tiles, dtiles, dq, seg_start,
                                                                            counts, part, didx, ddist, damb);
    CU(cudaGetLastError());
    CU(cudaEventRecord(sc.ev[4], p->stream));
    std::vector<uint32_t> amb(n), hcounts(QK), hvalid(n_tiles);
    int herr = 0;
    CU(wait_stream(p));
    CU(cudaMemcpyAsync(idx, didx, (size_t)n * 4, cudaMemcpyDeviceToHost, p->stream));
    CU(cudaMemcpyAsync(dist, ddist, (size_t)n * 8, cudaMemcpyDeviceToHost, p->stream));
    CU(cudaMemcpyAsync(amb.data(), damb, (size_t)n * 4, cudaMemcpyDeviceToHost, p->stream));
    CU(cudaMemcpyAsync(hcounts.data(), counts, (size_t)QK * 4, cudaMemcpyDeviceToHost, p->stream));
    CU(cudaMemcpyAsync(hvalid.data(), tile_valid, (size_t)n_tiles * 4, cudaMemcpyDeviceToHost, p->stream));
    CU(cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CU(wait_stream(p));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, sc.ev[2], sc.ev[3]));
    float fill_ms = 0.f;
    CU(cudaEventElapsedTime(&p->nn_prep_ms, sc.ev[0], sc.ev[1]));
    CU(cudaEventElapsedTime(&fill_ms, sc.ev[5], sc.ev[2]));
    p->nn_prep_ms += fill_ms;
    CU(cudaEventElapsedTime(&p->nn_merge_ms, sc.ev[3], sc.ev[4]));
    p->nn_tiles_ms = ms;
    p->nn_segments = n_work;
    if (herr) return fail(IMOMD_ERR_CUDA, "imomd_nearest_batch: pipeline barrier timed out");
    if (kernel_ms) *kernel_ms = ms;
    if (bytes_streamed)
    {
        uint64_t b = 0;
        for (int t = 0; t < n_tiles; ++t)
            b += (uint64_t)hvalid[t] * sizeof(VRec) + (uint64_t)hcounts[tiles[t].qt] * 8 +
                 (uint64_t)tiles[t].n * 36;
        *bytes_streamed = b;
    }
    // lookups the chord ranking could not decide go through the exact path
    std::vector<int> redo;
    for (int i = 0; i < n; ++i)
        if (amb[i]) redo.push_back(i);
    if (n_refined) *n_refined = (uint32_t)redo.size();
    if (!redo.empty())
    {
        std::vector<imomd_nn_query> rq(redo.size());
        std::vector<uint32_t> ri(redo.size());
        std::vector<double> rd(redo.size());
        for (size_t j = 0; j < redo.size(); ++j) rq[j] = q[redo[j]];
        int rc = imomd_nearest(p, (int)redo.size(), rq.data(), ri.data(), rd.data());
        if (rc != IMOMD_OK) return rc;
        for (size_t j = 0; j < redo.size(); ++j)
        {
            idx[redo[j]] = ri[j];
            dist[redo[j]] = rd[j];
        }
    }
    return IMOMD_OK;
}

int imomd_nearest_batch_last_timing(imomd_planner* p, float* prep_ms, float* tiles_ms, float* merge_ms,
                                    uint32_t* n_segments)
{
    if (!p) return fail(IMOMD_ERR_ARG, "null argument");
    if (prep_ms) *prep_ms = p->nn_prep_ms;
    if (tiles_ms) *tiles_ms = p->nn_tiles_ms;
    if (merge_ms) *merge_ms = p->nn_merge_ms;
    if (n_segments) *n_segments = p->nn_segments;
    return IMOMD_OK;
}

int imomd_ctx_create(int device, imomd_ctx** out)
{
    if (!out) return fail(IMOMD_ERR_ARG, "null argument");
    *out = nullptr;
    CU(cudaSetDevice(device));
    imomd_ctx* c = new (std::nothrow) imomd_ctx();
    if (!c) return fail(IMOMD_ERR_NOMEM, "out of host memory");
    c->device = device;
#define XCU(call) CU_OR(imomd_ctx_destroy(c), call)
    XCU(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    XCU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    XCU(dmalloc(&c->D, (size_t)IMOMD_KMAX * IMOMD_KMAX));
    XCU(dmalloc(&c->bad, 1));
#undef XCU
    *out = c;
    return IMOMD_OK;
}

int imomd_ctx_destroy(imomd_ctx* c)
{
    if (!c) return IMOMD_OK;
    cudaSetDevice(c->device);
    cudaFree(c->D);
    cudaFree(c->tours);
    cudaFree(c->lens);
    cudaFree(c->cost);
    cudaFree(c->bad);
    if (c->pin) cudaFreeHost(c->pin);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return IMOMD_OK;
}

int imomd_ga_eval(imomd_ctx* c, int K, const double* D, int n_tours, int stride,
                  const int32_t* tours, const int32_t* lens, double* cost_out)
{
    if (!c || !D || !tours || !lens || !cost_out) return fail(IMOMD_ERR_ARG, "null argument");
    if (K < 1 || K > (int)IMOMD_KMAX) return fail(IMOMD_ERR_ARG, "K out of range");
    if (stride < 1 || stride > kGaMaxLen) return fail(IMOMD_ERR_ARG, "stride must be in [1, %d]", kGaMaxLen);
    if (n_tours <= 0) return n_tours == 0 ? IMOMD_OK : fail(IMOMD_ERR_ARG, "negative tour count");
    CU(cudaSetDevice(c->device));
    const size_t need = (size_t)n_tours * stride;
    if (need > c->cap_tours)
    {
        cudaFree(c->tours);
        c->tours = nullptr;
        c->cap_tours = 0;
        CU(dmalloc(&c->tours, need + need / 4));
        c->cap_tours = need + need / 4;
    }
    if ((size_t)n_tours > c->cap_n)
    {
        cudaFree(c->lens);
        cudaFree(c->cost);
        c->lens = nullptr;
        c->cost = nullptr;
        c->cap_n = 0;
        const size_t cap = (size_t)n_tours + (size_t)n_tours / 4;
        CU(dmalloc(&c->lens, cap));
        CU(dmalloc(&c->cost, cap));
        c->cap_n = cap;
    }
    // pinned staging area: [D | tours | lens | cost out | bad flag]
    const size_t bD = (size_t)K * K * 8, bT = need * 4, bL = (size_t)n_tours * 4, bC = (size_t)n_tours * 8;
    const size_t oT = (bD + 15) / 16 * 16, oL = oT + (bT + 15) / 16 * 16, oC = oL + (bL + 15) / 16 * 16,
                 oB = oC + (bC + 15) / 16 * 16, total = oB + 16;
    if (total > c->pin_bytes)
    {
        if (c->pin) cudaFreeHost(c->pin);
        c->pin = nullptr;
        c->pin_bytes = 0;
        CU(cudaHostAlloc((void**)&c->pin, total + total / 4, cudaHostAllocDefault));
        c->pin_bytes = total + total / 4;
    }
    std::memcpy(c->pin, D, bD);
    std::memcpy(c->pin + oT, tours, bT);
    std::memcpy(c->pin + oL, lens, bL);
    CU(cudaMemsetAsync(c->bad, 0, sizeof(int), c->stream));
    CU(cudaMemcpyAsync(c->D, c->pin, bD, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->tours, c->pin + oT, bT, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->lens, c->pin + oL, bL, cudaMemcpyHostToDevice, c->stream));
    int warps_per_block = 8;
    int blocks = std::min((n_tours + warps_per_block - 1) / warps_per_block, c->sm_count * 8);
    k_ga_eval<<<blocks, 256, (size_t)K * K * 8, c->stream>>>(K, c->D, n_tours, stride, c->tours,
                                                             c->lens, c->cost, c->bad);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->pin + oC, c->cost, bC, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->pin + oB, c->bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    int bad = 0;
    std::memcpy(&bad, c->pin + oB, sizeof(int));
    if (bad) return fail(IMOMD_ERR_ARG, "tour %d: bad length or id out of range", bad - 1);
    std::memcpy(cost_out, c->pin + oC, bC);
    return IMOMD_OK;
}

}  // extern "C"
