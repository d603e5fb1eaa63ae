// csrc/imomd_nn_tiles.cuh -- brute-force tile kernel for LARGE batches of nearest-frontier
// lookups (imomd_nearest_batch): the bandwidth-shaped variant of steerNewNode_'s arg-min
// (src/imomd_rrt_star.cpp:406-418 upstream).
//
// A tile = one (query, tree) frontier and up to kTileLookups (4) samples against it.  The
// frontier lives in 1 KB chunks of 32 records (csrc/imomd_layout.h) of which only the first
// `valid` are occupied.  The launch is cut into SEGMENTS of kSegChunks chunks of one tile
// (k_nn_prepare / k_nn_fill_work); persistent CTAs take segments from an atomic counter, so
// the launch ends within one segment (~40 KB) of perfect balance whatever the frontier sizes
// (the counter's answer, the segment record and the first chunk ids of the next segment are
// requested a segment / a stage ahead: no dependent global load sits between two stages).
// Per CTA a producer warp streams the segment into a 3-stage shared-memory ring (8 KB = 256
// records per stage) with bulk asynchronous copies (cp.async.bulk ... mbarrier::complete_tx,
// the TMA engine): one copy per chunk, of its OCCUPIED records only, packed back to back in
// the stage -- no padding crosses the memory system and no consumer lane idles on an empty
// slot.  Two consumer warps rank the staged records against the tile's samples and keep (best,
// lowest id on ties, runner-up) per sample.  The CTA is small on purpose (96 threads, 8 CTAs
// per SM): a bulk copy is issued through the uniform datapath, one lane at a time (~10
// instructions per copy), so the producer warp is the scarce resource -- one producer per two
// consumer warps measured best (0.098 ms per launch against 0.120 ms with 1 : 4; ring depth
// beyond 24 KB per CTA did not matter).  A record is first screened by its dot product with the sample (3 FP64 operations):
// chord^2 = |v|^2 + |q|^2 - 2 v.q, so v.q < 1 - cap/2 - margin proves chord^2 > cap, where cap
// is an upper bound of the runner-up (seeded by k_nn_prepare from the first record of every
// chunk, tightened by what the thread has seen); only survivors get the exact squared chord
// of the pruned path.  The screen never changes an answer: every record whose chord^2 does
// not exceed the final runner-up passes it.  Full/empty mbarriers hand the stages back and
// forth; k_nn_merge folds the per-segment partials and writes the answers.
//
// Algorithmic bytes: 32 B per occupied frontier record per tile + 8 B per chunk-list entry +
// 24 B per lookup + 12 B per answer (brute-force variant).  Lookups whose two
// best candidates are closer than the rounding band are flagged and re-done by the exact path
// (k_nearest).
#ifndef IMOMD_NN_TILES_CUH
#define IMOMD_NN_TILES_CUH

#include "imomd_core.cuh"

namespace imomd {

constexpr int kTileLookups = 4;        // samples per tile
constexpr int kStageChunks = 16;       // at most this many chunks per pipeline stage (one producer lane each)
constexpr int kStageRecs = 256;        // records per pipeline stage (8 KB)
#ifndef IMOMD_NN_STAGES
#define IMOMD_NN_STAGES 3
#endif
#ifndef IMOMD_NN_CW
#define IMOMD_NN_CW 2
#endif
#ifndef IMOMD_NN_CTAS
#define IMOMD_NN_CTAS 8
#endif
constexpr int kStages = IMOMD_NN_STAGES;
constexpr int kSegChunks = 64;         // chunks per segment (work item)
constexpr int kSegStagesMax = kSegChunks * IMOMD_CHUNK / kStageRecs;   // 8: a stage short of the end holds > 256 - 32 records or 16 chunks
constexpr int kConsumerWarps = IMOMD_NN_CW;
constexpr int kTileThreads = 32 * (kConsumerWarps + 1);   // warp 0 = producer
constexpr int kTileCtasPerSm = IMOMD_NN_CTAS;
constexpr double kDotMargin = 1e-14;   // >= 4x the rounding of |v|^2, |q|^2 and the dot product

struct NnTile
{
    int32_t qt;                // query * K + tree
    int32_t n;                 // lookups in this tile (<= kTileLookups)
    int32_t lookup[kTileLookups];
};

// per tile, written by k_nn_prepare: unit vectors of the samples and an upper bound of each
// sample's runner-up chord^2 (unused sample slots: zero vector, bound 1 -> always screened out)
struct __align__(16) NnSeed
{
    double qx[kTileLookups], qy[kTileLookups], qz[kTileLookups], ub[kTileLookups];
};

// one segment: kSegChunks consecutive entries of a tile's chunk list, packed greedily into stages
// of <= kStageRecs occupied records and <= kStageChunks chunks (k_nn_fill_work).  64 bytes: the
// producer warp fetches a segment with one load per lane.
struct __align__(16) NnWork
{
    int32_t tile;
    int32_t qt;                // query * K + tree
    uint32_t chunk0;           // first entry of the tile's chunk list
    uint32_t n_stages;
    uint32_t stage[12];        // [0, n_stages): first entry (bits 0-6, relative to chunk0) | chunks << 7
};
static_assert(sizeof(NnWork) == 64, "NnWork is fetched as 16 words");
static_assert(kSegStagesMax <= 12, "stage table of a segment");

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

__device__ __forceinline__ double chord2(const VRec& v, double qx, double qy, double qz)
{
    double dx = v.x - qx, dy = v.y - qy, dz = v.z - qz;
    return fma(dz, dz, fma(dy, dy, dx * dx));
}

__device__ __forceinline__ Best warp_best(Best x)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
    {
        Best o;
        o.v1 = __shfl_down_sync(0xFFFFFFFFu, x.v1, off);
        o.id1 = __shfl_down_sync(0xFFFFFFFFu, x.id1, off);
        o.v2 = __shfl_down_sync(0xFFFFFFFFu, x.v2, off);
        x = best_merge(x, o);
    }
    return x;
}

// ---- preparation, one CTA per tree that has lookups ------------------------------------------
// (1) flat (chunk id, occupied records) list of the tree: one thread per occupied cell walks the
//     cell's chunk chain (the dependent loads of 128 chains overlap);
// (2) per tile of the tree: unit vectors of the samples and an upper bound of each sample's
//     runner-up chord^2 from the FIRST record of every chunk (a spatially uniform subsample of
//     the frontier: the chunks follow the hash cells; the tile kernel finds these sectors in L2
//     afterwards).  For the tree's first tile the records are requested inside the chain walk,
//     right when a chunk id is known; further tiles (more than 4 lookups on one frontier) re-read
//     the heads through the list;
// (3) the tile's segments are reserved with one atomicAdd (their order does not matter).
#ifndef IMOMD_NN_PREP
#define IMOMD_NN_PREP 64
#endif
constexpr int kPrepThreads = IMOMD_NN_PREP;

struct NnPrepShared
{
    double q[3][kTileLookups];
    Best part[kPrepThreads / 32][kTileLookups];
    uint32_t warp_sum[kPrepThreads / 32];
};

__device__ __forceinline__ void nn_prep_samples(NnPrepShared& sh, const NnTile& t,
                                                const imomd_nn_query* __restrict__ qs)
{
    if (threadIdx.x < kTileLookups)
    {
        double mx = 0, my = 0, mz = 0;
        if ((int)threadIdx.x < t.n)
        {
            imomd_nn_query nq = qs[t.lookup[threadIdx.x]];
            double la = nq.lat * kDeg2Rad, lo = nq.lon * kDeg2Rad;
            double cl = cos(la);
            mx = cl * cos(lo);
            my = cl * sin(lo);
            mz = sin(la);
        }
        sh.q[0][threadIdx.x] = mx;
        sh.q[1][threadIdx.x] = my;
        sh.q[2][threadIdx.x] = mz;
    }
}

// block reduction of the per-thread bounds, seed record and segment reservation of one tile
__device__ __forceinline__ void nn_prep_finish(NnPrepShared& sh, int tile, Best* b, uint32_t n_chunks,
                                               uint32_t n_valid, NnSeed* __restrict__ seed,
                                               uint32_t* __restrict__ tile_valid, uint32_t* __restrict__ seg_start,
                                               uint32_t* __restrict__ n_work)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int l = 0; l < kTileLookups; ++l)
    {
        Best x = warp_best(b[l]);
        if (lane == 0) sh.part[warp][l] = x;
    }
    __syncthreads();
    if (threadIdx.x < kTileLookups)
    {
        Best x = sh.part[0][threadIdx.x];
        for (int w = 1; w < kPrepThreads / 32; ++w) x = best_merge(x, sh.part[w][threadIdx.x]);
        seed[tile].qx[threadIdx.x] = sh.q[0][threadIdx.x];
        seed[tile].qy[threadIdx.x] = sh.q[1][threadIdx.x];
        seed[tile].qz[threadIdx.x] = sh.q[2][threadIdx.x];
        seed[tile].ub[threadIdx.x] = x.v2;
    }
    if (threadIdx.x == 0)
    {
        tile_valid[tile] = n_valid;
        seg_start[tile] = atomicAdd(n_work, (n_chunks + kSegChunks - 1) / kSegChunks);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kPrepThreads)
k_nn_prepare(const __grid_constant__ PlannerDev P, const uint32_t* __restrict__ tile_begin,
             const NnTile* __restrict__ tiles, const imomd_nn_query* __restrict__ qs, uint32_t* lists,
             uint32_t* counts, NnSeed* __restrict__ seed, uint32_t* __restrict__ tile_valid,
             uint32_t* __restrict__ seg_start, uint32_t* __restrict__ n_work)
{
    __shared__ NnPrepShared sh;
    const int qt = blockIdx.x;
    const uint32_t t0 = tile_begin[qt], t1 = tile_begin[qt + 1];
    if (t0 == t1) return;                               // no lookup on this tree
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = qt / P.K, k = qt % P.K;
    TreeRef T = tree_ref(P, q, k);
    const uint32_t n_occ = P.query[q].n_occ[k];
    uint32_t* out = lists + (size_t)qt * 2 * P.chunks;
    const VRec* rec = P.rec + (size_t)qt * (size_t)P.chunks * IMOMD_CHUNK;
    // Software pipeline over the batches of cells (three dependent loads per cell: occupied-cell
    // id -> header -> first record): the ids are requested two batches ahead, the headers one
    // batch ahead, so that a batch only waits for its records.  The first batch is requested
    // while the samples are being prepared.
    uint32_t cnt = 0, head = IMOMD_NONE, c_next = IMOMD_NONE;
    {
        uint32_t c_first = IMOMD_NONE;
        if (threadIdx.x < n_occ) c_first = T.occ[threadIdx.x];
        if (threadIdx.x + kPrepThreads < n_occ) c_next = T.occ[threadIdx.x + kPrepThreads];
        if (c_first != IMOMD_NONE)
        {
            cnt = T.cell_cnt[c_first];
            head = T.cell_head[c_first];
        }
    }
    nn_prep_samples(sh, tiles[t0], qs);
    __syncthreads();
    Best b[kTileLookups];
#pragma unroll
    for (int l = 0; l < kTileLookups; ++l) b[l] = best_empty();
    uint32_t total = 0, n_valid = 0;
    for (uint32_t base = 0; base < n_occ; base += kPrepThreads)
    {
        uint32_t c_next2 = IMOMD_NONE, cnt_next = 0, head_next = IMOMD_NONE;
        if (base + 2 * kPrepThreads + threadIdx.x < n_occ) c_next2 = T.occ[base + 2 * kPrepThreads + threadIdx.x];
        if (c_next != IMOMD_NONE)
        {
            cnt_next = T.cell_cnt[c_next];
            head_next = T.cell_head[c_next];
        }
        const uint32_t nch = (cnt + IMOMD_CHUNK - 1) / IMOMD_CHUNK;
        uint32_t inc = nch, vinc = cnt;
        for (int off = 1; off < 32; off <<= 1)
        {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, off);
            if (lane >= off) inc += t;
        }
        for (int off = 16; off > 0; off >>= 1) vinc += __shfl_down_sync(0xFFFFFFFFu, vinc, off);
        if (lane == 31) sh.warp_sum[warp] = inc;
        __syncthreads();
        uint32_t before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < kPrepThreads / 32; ++w)
        {
            const uint32_t x = sh.warp_sum[w];
            if (w < warp) before += x;
            all += x;
        }
        __syncthreads();
        if (lane == 0) n_valid += vinc;                 // per-warp partial, folded below
        const uint32_t at = total + before + inc - nch;
        uint32_t n = cnt ? (cnt - 1) % IMOMD_CHUNK + 1 : 0;
        uint32_t chunk = head;
        for (uint32_t j = 0; j < nch; ++j)
        {
            const VRec v = rec[(size_t)chunk * IMOMD_CHUNK];
            const uint32_t nxt = j + 1 < nch ? T.chunk_next[chunk] : IMOMD_NONE;
            out[2 * (at + j)] = chunk;
            out[2 * (at + j) + 1] = n;
#pragma unroll
            for (int l = 0; l < kTileLookups; ++l)
                best_add(b[l], chord2(v, sh.q[0][l], sh.q[1][l], sh.q[2][l]), v.id);
            n = IMOMD_CHUNK;
            chunk = nxt;
        }
        total += all;
        cnt = cnt_next;
        head = head_next;
        c_next = c_next2;
    }
    // occupied records of the tree
    if (lane == 0) sh.warp_sum[warp] = n_valid;
    __syncthreads();
    n_valid = 0;
#pragma unroll
    for (int w = 0; w < kPrepThreads / 32; ++w) n_valid += sh.warp_sum[w];
    if (threadIdx.x == 0) counts[qt] = total;
    nn_prep_finish(sh, (int)t0, b, total, n_valid, seed, tile_valid, seg_start, n_work);
    // further tiles of the same frontier: the heads again, through the list written above
    for (uint32_t tile = t0 + 1; tile < t1; ++tile)
    {
        nn_prep_samples(sh, tiles[tile], qs);
        __syncthreads();
#pragma unroll
        for (int l = 0; l < kTileLookups; ++l) b[l] = best_empty();
        for (uint32_t i = threadIdx.x; i < total; i += kPrepThreads)
        {
            const VRec v = rec[(size_t)out[2 * i] * IMOMD_CHUNK];
#pragma unroll
            for (int l = 0; l < kTileLookups; ++l)
                best_add(b[l], chord2(v, sh.q[0][l], sh.q[1][l], sh.q[2][l]), v.id);
        }
        nn_prep_finish(sh, (int)tile, b, total, n_valid, seed, tile_valid, seg_start, n_work);
    }
}

// one warp per tile: the tile's segments, each packed greedily into stages of <= kStageRecs
// occupied records and <= kStageChunks chunks (prefix sums + ballots)
__global__ void k_nn_fill_work(const __grid_constant__ PlannerDev P, int n_tiles, const NnTile* __restrict__ tiles,
                               const uint32_t* __restrict__ lists, const uint32_t* __restrict__ counts,
                               const uint32_t* __restrict__ seg_start, NnWork* __restrict__ work)
{
    const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (tile >= n_tiles) return;
    const int32_t qt = tiles[tile].qt;
    const uint32_t n_chunks = counts[qt];
    const uint32_t w0 = seg_start[tile];
    const uint32_t* list = lists + (size_t)qt * 2 * P.chunks;
    for (uint32_t chunk0 = 0, w = w0; chunk0 < n_chunks; chunk0 += kSegChunks, ++w)
    {
        const uint32_t n = min((uint32_t)kSegChunks, n_chunks - chunk0);
        uint32_t p0 = (uint32_t)lane < n ? list[2 * (chunk0 + lane) + 1] : 0u;
        uint32_t p1 = (uint32_t)lane + 32 < n ? list[2 * (chunk0 + 32 + lane) + 1] : 0u;
        for (int off = 1; off < 32; off <<= 1)
        {
            uint32_t a = __shfl_up_sync(0xFFFFFFFFu, p0, off), c = __shfl_up_sync(0xFFFFFFFFu, p1, off);
            if (lane >= off)
            {
                p0 += a;
                p1 += c;
            }
        }
        p1 += __shfl_sync(0xFFFFFFFFu, p0, 31);          // inclusive prefix over the 64 entries
        uint32_t word = lane == 0 ? (uint32_t)tile : lane == 1 ? (uint32_t)qt : lane == 2 ? chunk0 : 0u;
        uint32_t first = 0, base = 0, ns = 0;
        while (first < n)
        {
            const uint32_t i0 = lane, i1 = lane + 32;
            const bool in0 = i0 >= first && i0 < first + kStageChunks && i0 < n && p0 - base <= (uint32_t)kStageRecs;
            const bool in1 = i1 >= first && i1 < first + kStageChunks && i1 < n && p1 - base <= (uint32_t)kStageRecs;
            const uint32_t cnt = __popc(__ballot_sync(0xFFFFFFFFu, in0)) + __popc(__ballot_sync(0xFFFFFFFFu, in1));
            if (lane == 4 + (int)ns) word = first | (cnt << 7);
            const uint32_t last = first + cnt - 1;
            const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, p0, last & 31), b1 = __shfl_sync(0xFFFFFFFFu, p1, last & 31);
            base = last < 32 ? b0 : b1;
            first += cnt;
            ++ns;
        }
        if (lane == 3) word = ns;
        if (lane < 16) reinterpret_cast<uint32_t*>(work + w)[lane] = word;
    }
}

// Rare path of the consumers (a few percent of the records): exact squared chord of a record
// that passed the screen of at least one sample.  Out of line on purpose, with the running
// results in local memory: the screening loop then carries nothing but the thresholds.
struct Thr4
{
    double t[kTileLookups];
};
__device__ __noinline__ Thr4 nn_rank_exact(const VRec v, const double* q /*[3][kTileLookups], shared*/, Thr4 thr,
                                           Best* b)
{
#pragma unroll
    for (int l = 0; l < kTileLookups; ++l)
    {
        const double qx = q[l], qy = q[kTileLookups + l], qz = q[2 * kTileLookups + l];
        const double dot = fma(v.z, qz, fma(v.y, qy, v.x * qx));      // the screening loop's own value
        if (!(dot < thr.t[l]))
        {
            best_add(b[l], chord2(v, qx, qy, qz), v.id);
            thr.t[l] = fmax(thr.t[l], 1.0 - 0.5 * b[l].v2 - kDotMargin);
        }
    }
    return thr;
}

struct NnStageMeta
{
    int32_t work;              // segment staged here, -1 = no more work
    int32_t flags;             // 1 = first stage of its segment, 2 = last
    uint32_t n_rec;            // occupied records packed in the stage
    int32_t pad;
};

struct __align__(16) NnShared
{
    VRec stage[kStages][kStageRecs];                   // 3 x 8 KB
    NnSeed seed[kStages];                              // meaningful in a segment's first stage
    double cur[kConsumerWarps][3 * kTileLookups];      // samples of the segment a consumer warp is in
    NnStageMeta meta[kStages];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    int error;
};

__global__ void __launch_bounds__(kTileThreads, kTileCtasPerSm)
k_nearest_tiles(const __grid_constant__ PlannerDev P, const NnSeed* __restrict__ seed,
                const NnWork* __restrict__ work,
                const uint32_t* __restrict__ n_work_ptr, uint32_t* __restrict__ next_work,
                const uint32_t* __restrict__ lists, Best* __restrict__ part, int* __restrict__ error)
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
    if (warp == 0)
    {
        // ---------------- producer: bulk copies chunk -> stage ----------------
        // Segments come from an atomic counter two ahead (the counter's answer, the segment
        // record and the first chunk ids of the next segment are all requested one segment /
        // one stage before they are needed: no dependent global load sits between two stages).
        // Lane j < 16 owns chunk j of a stage and issues the copy of its occupied records,
        // packed behind lane j-1's.
        const uint32_t n_work = *n_work_ptr;
        uint32_t w0 = 0, w1 = 0;
        if (lane == 0)
        {
            w0 = atomicAdd(next_work, 1u);
            w1 = atomicAdd(next_work, 1u);
        }
        w0 = __shfl_sync(0xFFFFFFFFu, w0, 0);
        w1 = __shfl_sync(0xFFFFFFFFu, w1, 0);
        // lane l < 16 holds word l of the segment record
        auto fetch_work = [&](uint32_t w) -> uint32_t {
            return (w < n_work && lane < 16) ? reinterpret_cast<const uint32_t*>(work + w)[lane] : 0u;
        };
        auto fetch_entry = [&](uint32_t qt, uint32_t chunk0, uint32_t desc) -> uint2 {
            const uint32_t off = desc & 127u, nc = desc >> 7;
            if ((uint32_t)lane >= nc) return make_uint2(IMOMD_NONE, 0u);
            return *reinterpret_cast<const uint2*>(lists + (size_t)qt * 2 * P.chunks +
                                                   2 * ((size_t)chunk0 + off + lane));
        };
        uint32_t X0 = fetch_work(w0), X1 = fetch_work(w1);
        uint2 mine = make_uint2(IMOMD_NONE, 0u);
        double seed_word = 0;
        if (w0 < n_work)
        {
            const uint32_t tile = __shfl_sync(0xFFFFFFFFu, X0, 0), qt = __shfl_sync(0xFFFFFFFFu, X0, 1);
            mine = fetch_entry(qt, __shfl_sync(0xFFFFFFFFu, X0, 2), __shfl_sync(0xFFFFFFFFu, X0, 4));
            if (lane < 16) seed_word = reinterpret_cast<const double*>(seed + tile)[lane];
        }
        while (w0 < n_work)
        {
            uint32_t w2 = 0;
            if (lane == 0) w2 = atomicAdd(next_work, 1u);       // consumed after this segment
            const uint32_t qt = __shfl_sync(0xFFFFFFFFu, X0, 1);
            const uint32_t chunk0 = __shfl_sync(0xFFFFFFFFu, X0, 2);
            const uint32_t n_stages = __shfl_sync(0xFFFFFFFFu, X0, 3);
            const VRec* rec = P.rec + (size_t)qt * (size_t)P.chunks * IMOMD_CHUNK;
            for (uint32_t s = 0; s < n_stages; ++s)
            {
                const uint32_t slot = (it + s) % kStages, round = (it + s) / kStages;
                const uint32_t chunk = mine.x, valid = mine.y;
                const double seed_now = seed_word;
                // what the next stage needs is requested before this stage's barrier wait
                if (s + 1 < n_stages)
                    mine = fetch_entry(qt, chunk0, __shfl_sync(0xFFFFFFFFu, X0, 4 + s + 1));
                else if (w1 < n_work)
                {
                    const uint32_t tile1 = __shfl_sync(0xFFFFFFFFu, X1, 0);
                    mine = fetch_entry(__shfl_sync(0xFFFFFFFFu, X1, 1), __shfl_sync(0xFFFFFFFFu, X1, 2),
                                       __shfl_sync(0xFFFFFFFFu, X1, 4));
                    if (lane < 16) seed_word = reinterpret_cast<const double*>(seed + tile1)[lane];
                }
                uint32_t inc = valid;                    // packed position of this lane's records
                for (int off = 1; off < kStageChunks; off <<= 1)
                {
                    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, off);
                    if (lane >= off) inc += t;
                }
                const uint32_t n_rec = __shfl_sync(0xFFFFFFFFu, inc, kStageChunks - 1);
                if (lane == 0)
                {
                    if (round > 0 && !mbar_wait(&sh.empty[slot], (round - 1) & 1)) sh.error = 1;
                }
                __syncwarp();
                if (s == 0 && lane < 16) reinterpret_cast<double*>(&sh.seed[slot])[lane] = seed_now;
                if (lane == 0)
                {
                    NnStageMeta m;
                    m.work = (int32_t)w0;
                    m.flags = (s == 0 ? 1 : 0) | (s + 1 == n_stages ? 2 : 0);
                    m.n_rec = n_rec;
                    m.pad = 0;
                    sh.meta[slot] = m;
                }
                __syncwarp();
                if (lane == 0) mbar_expect_tx(&sh.full[slot], n_rec * (uint32_t)sizeof(VRec));
                __syncwarp();
                if (valid)
                    bulk_g2s(&sh.stage[slot][inc - valid], rec + (size_t)chunk * IMOMD_CHUNK,
                             valid * (uint32_t)sizeof(VRec), &sh.full[slot]);
                __syncwarp();
            }
            it += n_stages;
            w0 = w1;
            X0 = X1;
            w1 = __shfl_sync(0xFFFFFFFFu, w2, 0);
            X1 = fetch_work(w1);
        }
        // end marker for the consumers
        {
            const uint32_t slot = it % kStages, round = it / kStages;
            if (lane == 0)
            {
                if (round > 0 && !mbar_wait(&sh.empty[slot], (round - 1) & 1)) sh.error = 1;
                NnStageMeta m;
                m.work = -1; m.flags = 0; m.n_rec = 0; m.pad = 0;
                sh.meta[slot] = m;
                mbar_arrive(&sh.full[slot]);
            }
        }
    }
    else
    {
        // ---------------- consumers: screen and rank the staged records ----------------
        const int c = threadIdx.x - 32;                 // 0 .. 127
        double qx[kTileLookups], qy[kTileLookups], qz[kTileLookups];
        Thr4 thr;
        Best b[kTileLookups];                           // local memory: touched by the rare path only
        bool any = false;                               // this thread ranked something in this segment
        const double* qsh = sh.seed[0].qx;
#pragma unroll
        for (int l = 0; l < kTileLookups; ++l)
        {
            qx[l] = qy[l] = qz[l] = 0;
            thr.t[l] = -INFINITY;
            b[l] = best_empty();
        }
        for (;; ++it)
        {
            const uint32_t slot = it % kStages, round = it / kStages;
            if (!mbar_wait(&sh.full[slot], round & 1))
            {
                sh.error = 1;
                break;
            }
            const NnStageMeta m = sh.meta[slot];
            if (m.work < 0) break;
            if (m.flags & 1)
            {
                // the rare path reads the samples from the warp's own copy in shared memory
#pragma unroll
                for (int l = 0; l < kTileLookups; ++l)
                {
                    qx[l] = sh.seed[slot].qx[l];
                    qy[l] = sh.seed[slot].qy[l];
                    qz[l] = sh.seed[slot].qz[l];
                    thr.t[l] = 1.0 - 0.5 * sh.seed[slot].ub[l] - kDotMargin;
                    b[l] = best_empty();
                }
                any = false;
                if (lane < 3 * kTileLookups)
                    sh.cur[warp - 1][lane] = reinterpret_cast<const double*>(&sh.seed[slot])[lane];
                __syncwarp();
                qsh = sh.cur[warp - 1];
            }
            const VRec* st = sh.stage[slot];
#pragma unroll 1
            for (uint32_t r = c; r < m.n_rec; r += kConsumerWarps * 32)
            {
                const VRec v = st[r];
                bool pass = false;
#pragma unroll
                for (int l = 0; l < kTileLookups; ++l)
                {
                    const double dot = fma(v.z, qz[l], fma(v.y, qy[l], v.x * qx[l]));
                    pass |= !(dot < thr.t[l]);
                }
                if (pass)
                {
                    thr = nn_rank_exact(v, qsh, thr, b);
                    any = true;
                }
            }
            if (m.flags & 2)
            {
                // most segments are far from most samples: nothing passed the screen in the whole warp
                const bool warp_any = __any_sync(0xFFFFFFFFu, any);
#pragma unroll
                for (int l = 0; l < kTileLookups; ++l)
                {
                    Best x = best_empty();
                    if (warp_any) x = warp_best(b[l]);
                    if (lane == 0) part[((size_t)m.work * kConsumerWarps + (warp - 1)) * kTileLookups + l] = x;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.empty[slot]);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && sh.error) *error = 1;
}

// one thread per (tile, sample): fold the segment partials, write the answer
__global__ void k_nn_merge(const __grid_constant__ PlannerDev P, int n_tiles, const NnTile* __restrict__ tiles,
                           const imomd_nn_query* __restrict__ qs, const uint32_t* __restrict__ seg_start,
                           const uint32_t* __restrict__ counts, const Best* __restrict__ part,
                           uint32_t* __restrict__ idx,
                           double* __restrict__ dist, uint32_t* __restrict__ ambiguous)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int tile = i / kTileLookups, l = i % kTileLookups;
    if (tile >= n_tiles) return;
    const NnTile t = tiles[tile];
    if (l >= t.n) return;
    Best b = best_empty();
    const uint32_t w0 = seg_start[tile], w1 = w0 + (counts[t.qt] + kSegChunks - 1) / kSegChunks;
    for (uint32_t w = w0; w < w1; ++w)
        for (int cw = 0; cw < kConsumerWarps; ++cw)
            b = best_merge(b, part[((size_t)w * kConsumerWarps + cw) * kTileLookups + l]);
    const int li = t.lookup[l];
    const imomd_nn_query nq = qs[li];
    idx[li] = b.id1;
    dist[li] = b.id1 == IMOMD_NONE ? INFINITY : haversine(P.g.lat[b.id1], P.g.lon[b.id1], nq.lat, nq.lon);
    const double band = 2e-12 * b.v1 + 4e-15 * sqrt(b.v1);
    ambiguous[li] = (b.id1 != IMOMD_NONE && b.v2 - b.v1 <= band) ? 1u : 0u;
}

}  // namespace imomd
#endif
