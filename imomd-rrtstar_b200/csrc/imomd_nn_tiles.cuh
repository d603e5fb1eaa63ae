// csrc/imomd_nn_tiles.cuh -- brute-force tile kernel for LARGE batches of nearest-frontier
// lookups (imomd_nearest_batch): the bandwidth-shaped variant of steerNewNode_'s arg-min
// (src/imomd_rrt_star.cpp:406-418 upstream).
//
// A tile = one (query, tree) frontier and up to kTileLookups (4) samples against it.  The
// frontier lives in 1 KB chunks of 32 records (csrc/imomd_layout.h); a producer warp streams
// the tree's chunks into a 4-stage shared-memory ring with bulk asynchronous copies
// (cp.async.bulk ... mbarrier::complete_tx, the TMA engine; one 1 KB copy per chunk because
// chunks of a cell chain are not contiguous), four consumer warps rank the staged records
// against the tile's samples by squared chord (every shared-memory read is a broadcast or
// conflict free) and keep (best, lowest id on ties, runner-up) per sample.  Full/empty
// mbarriers hand the stages back and forth; the kernel is persistent over tiles.
//
// Algorithmic bytes: 32 B per frontier record per tile + 24 B per lookup + 12 B per answer
// (brute-force variant).  With kTileLookups = 4 the FP64 work per staged byte
// stays below the machine balance, so the kernel is HBM-bound when the frontiers do not fit
// in L2.  Lookups whose two best candidates are closer than the rounding band are flagged
// and re-done by the exact path (k_nearest).
#ifndef IMOMD_NN_TILES_CUH
#define IMOMD_NN_TILES_CUH

#include "imomd_core.cuh"

namespace imomd {

constexpr int kTileLookups = 4;        // samples per tile
constexpr int kStageChunks = 8;        // 1 KB chunks per pipeline stage
constexpr int kStages = 4;
constexpr int kConsumerWarps = 4;
constexpr int kTileThreads = 32 * (kConsumerWarps + 1);   // warp 0 = producer

struct NnTile
{
    int32_t qt;                // query * K + tree
    int32_t n;                 // lookups in this tile (<= kTileLookups)
    int32_t lookup[kTileLookups];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: returns false if the barrier never flipped (reported, never hangs the GPU)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok = 0;
    for (uint32_t spin = 0; spin < (1u << 28); ++spin)
    {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// flat (chunk id, valid records) list of every tree: one warp per tree walks the occupied
// cells and their chunk chains
__global__ void k_tree_chunk_lists(const __grid_constant__ PlannerDev P, uint32_t* lists, uint32_t* counts)
{
    int qt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (qt >= P.Q * P.K) return;
    int q = qt / P.K, k = qt % P.K;
    TreeRef T = tree_ref(P, q, k);
    uint32_t n_occ = P.query[q].n_occ[k];
    uint32_t* out = lists + (size_t)qt * 2 * P.chunks;
    uint32_t total = 0;
    for (uint32_t base = 0; base < n_occ; base += 32)
    {
        uint32_t i = base + lane;
        uint32_t cnt = 0, head = IMOMD_NONE, nch = 0;
        if (i < n_occ)
        {
            uint32_t c = T.occ[i];
            cnt = T.cell_cnt[c];
            head = T.cell_head[c];
            nch = (cnt + IMOMD_CHUNK - 1) / IMOMD_CHUNK;
        }
        uint32_t inc = nch;
        for (int off = 1; off < 32; off <<= 1)
        {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, off);
            if (lane >= off) inc += t;
        }
        uint32_t at = total + inc - nch;
        uint32_t n = cnt ? (cnt - 1) % IMOMD_CHUNK + 1 : 0;
        uint32_t chunk = head;
        for (uint32_t j = 0; j < nch; ++j)
        {
            out[2 * (at + j)] = chunk;
            out[2 * (at + j) + 1] = n;
            n = IMOMD_CHUNK;
            chunk = T.chunk_next[chunk];
        }
        total += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
    if (lane == 0) counts[qt] = total;
}

struct __align__(16) NnShared
{
    VRec stage[kStages][kStageChunks * IMOMD_CHUNK];   // 4 x 8 KB
    uint32_t valid[kStages][kStageChunks];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    double qx[kTileLookups], qy[kTileLookups], qz[kTileLookups];
    Best part[kConsumerWarps * 32];
    int error;
};

__global__ void __launch_bounds__(kTileThreads, 4)
k_nearest_tiles(const __grid_constant__ PlannerDev P, int n_tiles, const NnTile* __restrict__ tiles,
                const imomd_nn_query* __restrict__ qs, const uint32_t* __restrict__ lists,
                const uint32_t* __restrict__ counts, uint32_t* __restrict__ idx,
                double* __restrict__ dist, uint32_t* __restrict__ ambiguous, int* __restrict__ error)
{
    extern __shared__ __align__(128) unsigned char nn_smem_raw[];
    NnShared& sh = *reinterpret_cast<NnShared*>(nn_smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0)
    {
        for (int s = 0; s < kStages; ++s)
        {
            mbar_init(&sh.full[s], 1);
            mbar_init(&sh.empty[s], kConsumerWarps);
        }
        sh.error = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t it = 0;                 // stages produced / consumed so far (same on both sides)
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
    {
        const NnTile t = tiles[tile];
        const uint32_t n_chunks = counts[t.qt];
        const uint32_t n_stages = (n_chunks + kStageChunks - 1) / kStageChunks;
        const uint32_t* list = lists + (size_t)t.qt * 2 * P.chunks;
        const VRec* rec = P.rec + (size_t)t.qt * (size_t)P.chunks * IMOMD_CHUNK;
        if (threadIdx.x < (unsigned)t.n)
        {
            imomd_nn_query nq = qs[t.lookup[threadIdx.x]];
            double la = nq.lat * kDeg2Rad, lo = nq.lon * kDeg2Rad;
            double cl = cos(la);
            sh.qx[threadIdx.x] = cl * cos(lo);
            sh.qy[threadIdx.x] = cl * sin(lo);
            sh.qz[threadIdx.x] = sin(la);
        }
        __syncthreads();
        if (warp == 0)
        {
            // ---------------- producer: bulk copies chunk -> stage ----------------
            // lane j of the producer warp owns chunk j of every stage: it fetches the chunk id
            // (prefetched one stage ahead) and issues its own 1 KB bulk copy
            uint32_t my_chunk = IMOMD_NONE, my_valid = 0;
            if (n_stages > 0 && lane < kStageChunks && (uint32_t)lane < n_chunks)
            {
                my_chunk = list[2 * lane];
                my_valid = list[2 * lane + 1];
            }
            for (uint32_t s = 0; s < n_stages; ++s)
            {
                const uint32_t slot = (it + s) % kStages, round = (it + s) / kStages;
                const uint32_t c0 = s * kStageChunks;
                const uint32_t nc = min((uint32_t)kStageChunks, n_chunks - c0);
                const uint32_t chunk = my_chunk, valid = my_valid;
                // next stage's ids are requested before this stage's barrier wait
                const uint32_t nxt = c0 + kStageChunks + lane;
                if (lane < kStageChunks && nxt < n_chunks)
                {
                    my_chunk = list[2 * nxt];
                    my_valid = list[2 * nxt + 1];
                }
                if (lane == 0)
                {
                    if (round > 0 && !mbar_wait(&sh.empty[slot], (round - 1) & 1)) sh.error = 1;
                }
                __syncwarp();
                if (lane < kStageChunks) sh.valid[slot][lane] = (uint32_t)lane < nc ? valid : 0u;
                __syncwarp();
                if (lane == 0) mbar_expect_tx(&sh.full[slot], nc * (uint32_t)(IMOMD_CHUNK * sizeof(VRec)));
                __syncwarp();
                if ((uint32_t)lane < nc)
                    bulk_g2s(&sh.stage[slot][lane * IMOMD_CHUNK], rec + (size_t)chunk * IMOMD_CHUNK,
                             (uint32_t)(IMOMD_CHUNK * sizeof(VRec)), &sh.full[slot]);
                __syncwarp();
            }
        }
        else
        {
            // ---------------- consumers: rank the staged records ----------------
            // each thread takes records c, c+128, ... of a stage and ranks them against ALL
            // samples of the tile (one record load feeds kTileLookups chord evaluations)
            const int c = threadIdx.x - 32;                 // 0 .. 127
            double qx[kTileLookups], qy[kTileLookups], qz[kTileLookups];
            Best b[kTileLookups];
#pragma unroll
            for (int l = 0; l < kTileLookups; ++l)
            {
                qx[l] = sh.qx[l];
                qy[l] = sh.qy[l];
                qz[l] = sh.qz[l];
                b[l] = best_empty();
            }
            for (uint32_t s = 0; s < n_stages; ++s)
            {
                const uint32_t slot = (it + s) % kStages, round = (it + s) / kStages;
                if (!mbar_wait(&sh.full[slot], round & 1)) sh.error = 1;
                const VRec* st = sh.stage[slot];
#pragma unroll
                for (int r = c; r < kStageChunks * IMOMD_CHUNK; r += kConsumerWarps * 32)
                {
                    if ((uint32_t)(r % IMOMD_CHUNK) < sh.valid[slot][r / IMOMD_CHUNK])
                    {
                        const VRec v = st[r];
#pragma unroll
                        for (int l = 0; l < kTileLookups; ++l)
                        {
                            double dx = v.x - qx[l], dy = v.y - qy[l], dz = v.z - qz[l];
                            double c2 = fma(dz, dz, fma(dy, dy, dx * dx));
                            // common case first: not better than the incumbent -> only the
                            // runner-up can move
                            if (c2 > b[l].v1)
                                b[l].v2 = fmin(b[l].v2, c2);
                            else
                                best_add(b[l], c2, v.id);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.empty[slot]);
            }
            // per-sample merge over the warp, one partial per (warp, sample)
#pragma unroll
            for (int l = 0; l < kTileLookups; ++l)
            {
                Best x = b[l];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1)
                {
                    Best o;
                    o.v1 = __shfl_down_sync(0xFFFFFFFFu, x.v1, off);
                    o.id1 = __shfl_down_sync(0xFFFFFFFFu, x.id1, off);
                    o.v2 = __shfl_down_sync(0xFFFFFFFFu, x.v2, off);
                    x = best_merge(x, o);
                }
                if (lane == 0) sh.part[(warp - 1) * kTileLookups + l] = x;
            }
        }
        it += n_stages;
        __syncthreads();
        if (threadIdx.x < (unsigned)t.n)
        {
            Best b = best_empty();
            for (int w = 0; w < kConsumerWarps; ++w) b = best_merge(b, sh.part[w * kTileLookups + threadIdx.x]);
            const int li = t.lookup[threadIdx.x];
            const imomd_nn_query nq = qs[li];
            idx[li] = b.id1;
            dist[li] = b.id1 == IMOMD_NONE
                           ? INFINITY
                           : haversine(P.g.lat[b.id1], P.g.lon[b.id1], nq.lat, nq.lon);
            const double band = 2e-12 * b.v1 + 4e-15 * sqrt(b.v1);
            ambiguous[li] = (b.id1 != IMOMD_NONE && b.v2 - b.v1 <= band) ? 1u : 0u;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && sh.error) *error = 1;
}

}  // namespace imomd
#endif
