// csrc/imomd_core.cuh -- device logic of the multi-tree expansion path.
//
// One CTA drives one query.  The reference's expansion is a strict sequence
// ("sequential dependence"), so the CTA runs a small state
// machine: thread 0 executes the inherently serial bookkeeping (JPS hops,
// choose-parent, frontier hash maintenance, the rewire cascade, the tree-to-tree
// connection check) and posts *service requests* that the whole CTA executes in
// parallel:
//
//   REQ_NN      nearest frontier vertex to the sample          (steerNewNode_ :406-418)
//   REQ_RESCAN  arg-min of the two-root heuristic over the frontier (updateExpandables :510-524)
//   REQ_MHINS   heuristics of freshly inserted frontier vertices    (updateExpandables :541-556)
//
// Both searches are the same *pruned arg-min over the uniform-grid frontier hash*:
// a conservative per-cell lower bound discards cells, the surviving 32-record
// chunks are evaluated record-per-thread, and a shuffle/shared-memory reduction
// returns (min, lowest id on exact ties, runner-up).  NN ranks by squared chord
// between unit vectors (== 4 * haversine `a`, monotone in the reference's
// distance); decisions closer than the rounding band are re-done with the full
// double-precision haversine and counted (near_ties / unresolved_ties).
//
// Everything is written against a tiny `Cta` policy (thread id, barrier, block
// reductions, shared counter) so that tests/hostsim can compile the very same
// logic for the host with a one-thread policy and check it against the oracle
// without a GPU.  That build is test infrastructure; the shipped library has no
// CPU path.
//
// Reference lines cited below are under /root/reference.
#ifndef IMOMD_CORE_CUH
#define IMOMD_CORE_CUH

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "imomd_layout.h"

#if defined(__CUDACC__)
#define IMOMD_HD __host__ __device__
#define IMOMD_NOINLINE __noinline__
#else
#define IMOMD_HD
#define IMOMD_NOINLINE
#endif

// Address-space facts the compiler cannot see through structs and references: with them the
// generic LD / ST of the device logic become LDG / LDS (no address-window resolution on the
// serial thread's critical path).  No-ops in the host simulation.
#ifdef __CUDA_ARCH__
#define IMOMD_ASSUME_GLOBAL(p) __builtin_assume(__isGlobal(p))
#define IMOMD_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#else
#define IMOMD_ASSUME_GLOBAL(p) ((void)0)
#define IMOMD_ASSUME_SHARED(p) ((void)0)
#endif


namespace imomd {

constexpr double kDeg2Rad = 3.14159265358979323846 / 180;
constexpr double kEarthRadius = 6371000.0;   // `int EARTH_RADIUS = 6371e3` upstream
constexpr double kCellPadDeg = 1e-9;         // cell rectangles are grown by this much
constexpr double kPruneSlack = 1e-9;         // relative slack of every pruning test
constexpr double kTieRel = 1e-12;            // default of PlannerDev::tie_rel: "cannot be decided in device arithmetic"

enum ReqType
{
    REQ_EXIT = 0,
    REQ_NN = 1,
    REQ_RESCAN = 2,
    REQ_MHINS = 3,
    REQ_CONN = 4,
    REQ_BFS = 5,
    REQ_CHECK = 6
};
enum Phase
{
    PH_NEXT_TREE = 0,
    PH_AFTER_NN = 1,
    PH_JPS = 2,
    PH_INSERT = 3,
    PH_FINISH = 4,
    PH_REWIRE = 5,
    PH_LAST_CONN = 6
};
constexpr uint32_t kCheckBatch = 64;   // rewire calls examined per speculative check
// Cell lower bounds a CTA keeps in shared memory between the two passes of a pruned search: one per
// OCCUPIED cell, at most kCtaLbCap (a frontier that occupies more cells recomputes its bounds in the
// second pass).  What the G x G table of round 1 took beyond that is L1 now.
constexpr uint32_t kCtaLbCap = 256;

// include/osm_converter/compute_haversine.hpp:41-58, same expression order.
IMOMD_HD IMOMD_NOINLINE double haversine(double lat1, double lon1, double lat2, double lon2)
{
    double lat1_rad = lat1 * kDeg2Rad;
    double lat2_rad = lat2 * kDeg2Rad;
    double diff_lat_rad = (lat2 - lat1) * kDeg2Rad;
    double diff_long_rad = (lon2 - lon1) * kDeg2Rad;
    double s1 = sin(diff_lat_rad / 2);
    double s2 = sin(diff_long_rad / 2);
    double a = s1 * s1 + cos(lat1_rad) * cos(lat2_rad) * (s2 * s2);
    double c = 2 * atan2(sqrt(a), sqrt(1 - a));
    return kEarthRadius * c;
}

// sin(h) and 2 asin(sqrt(a)) for the small arguments of a road map (half-differences below 2^-6
// rad = 1.8 degrees of latitude / longitude, a <= 2^-12 = distances below 200 km): Taylor
// polynomials whose remainders are below 2^-80 relative, libm beyond.  The heuristic VALUES are
// CUDA arithmetic either way (compared to 1e-12 relative with the reference's; comparisons closer
// than that are decided on the host's glibc values): what changes is a few
// dozen instructions instead of the ~400 of two sin() and an atan2() per call -- a tenth of all
// instructions the batched kernel executed, and 13 KB less hot code in the instruction cache.
IMOMD_HD inline double sin_small(double h)
{
    const double t = h * h;
    if (t > 1.0 / 4096.0) return sin(h);
    return h * (1.0 + t * (-1.0 / 6.0 + t * (1.0 / 120.0 + t * (-1.0 / 5040.0 + t * (1.0 / 362880.0)))));
}
IMOMD_HD inline double central_angle_of_a(double a)
{
    if (a > 1.0 / 4096.0) return 2 * atan2(sqrt(a), sqrt(1 - a));
    // 2 asin(x), x = sqrt(a): x (1 + a/6 + 3 a^2/40 + 15 a^3/336 + 105 a^4/3456 + ...)
    const double x = sqrt(a);
    return 2 * (x * (1.0 + a * (1.0 / 6.0 + a * (3.0 / 40.0 + a * (15.0 / 336.0 + a * (105.0 / 3456.0))))));
}

// The heuristics' haversine from the per-vertex cosines precomputed at upload (GraphDev::coslat holds
// cos(lat * kDeg2Rad), the very operand of the two cos() calls above).
IMOMD_HD IMOMD_NOINLINE double haversine_pc(double lat1, double lon1, double cos1, double lat2, double lon2,
                                            double cos2)
{
    double diff_lat_rad = (lat2 - lat1) * kDeg2Rad;
    double diff_long_rad = (lon2 - lon1) * kDeg2Rad;
    double s1 = sin_small(diff_lat_rad / 2);
    double s2 = sin_small(diff_long_rad / 2);
    double a = s1 * s1 + cos1 * cos2 * (s2 * s2);
    return kEarthRadius * central_angle_of_a(a);
}

// distance for a haversine `a` value (used for the cell lower bounds)
IMOMD_HD inline double dist_of_a(double a)
{
    if (a > 1.0) a = 1.0;
    return kEarthRadius * (2 * atan2(sqrt(a), sqrt(1 - a)));
}

// Polynomial bounds of sin on [0, pi/2] (alternating Taylor remainders): the cell bounds run
// for every occupied cell of every search, and a bound only has to be valid, not tight to the
// last bit -- the pruning slack absorbs the rounding of these few operations.
IMOMD_HD inline double sin_lower(double h)
{
    double v = h - h * h * h * (1.0 / 6.0);
    return v > 0.0 ? v : 0.0;
}
IMOMD_HD inline double sin_upper(double h)
{
    double h2 = h * h;
    double v = h - h * h2 * (1.0 / 6.0) + h * h2 * h2 * (1.0 / 120.0);
    return v < 1.0 ? v : 1.0;
}

// Lower bound of haversine `a` between (qlat, qlon) and any point of hash cell c.
// a = sin^2(dphi/2) + cos(phi_q) cos(phi_v) sin^2(dlam/2) is increasing in |dphi| and
// |dlam| (both below pi here) and cos(phi_v) >= cosrow[row].
IMOMD_HD IMOMD_NOINLINE double cell_lower_a(const GraphDev& g, uint32_t c, double qlat, double qlon,
                                    double cos_qlat)
{
    if (!g.prune) return 0.0;
    int row = (int)(c / (uint32_t)g.G), colm = (int)(c % (uint32_t)g.G);
    double lo = g.lat0 + row * g.hlat - kCellPadDeg;
    double hi = g.lat0 + (row + 1) * g.hlat + kCellPadDeg;
    double dphi = 0.0;
    if (qlat < lo) dphi = lo - qlat;
    else if (qlat > hi) dphi = qlat - hi;
    double lo2 = g.lon0 + colm * g.hlon - kCellPadDeg;
    double hi2 = g.lon0 + (colm + 1) * g.hlon + kCellPadDeg;
    double dlam = 0.0;
    if (qlon < lo2) dlam = lo2 - qlon;
    else if (qlon > hi2) dlam = qlon - hi2;
    // sin(h) >= h - h^3/6 for every h >= 0 (clamped at 0; h <= pi/2 here): a bound needs no libm
    double s1 = sin_lower(dphi * kDeg2Rad / 2);
    double s2 = sin_lower(dlam * kDeg2Rad / 2);
    double a = s1 * s1 + cos_qlat * g.cosrow[row] * (s2 * s2);
    return a * (1.0 - kPruneSlack);
}

// Upper bound of haversine `a` between (qlat, qlon) and any point of hash cell c (the far
// corner of the padded cell rectangle with the largest cos(lat) of the row).
IMOMD_HD IMOMD_NOINLINE double cell_upper_a(const GraphDev& g, uint32_t c, double qlat, double qlon,
                                    double cos_qlat)
{
    if (!g.prune) return 1.0;
    int row = (int)(c / (uint32_t)g.G), colm = (int)(c % (uint32_t)g.G);
    double lo = g.lat0 + row * g.hlat - kCellPadDeg;
    double hi = g.lat0 + (row + 1) * g.hlat + kCellPadDeg;
    double d1 = fabs(qlat - lo), d2 = fabs(qlat - hi);
    double dphi = d1 > d2 ? d1 : d2;
    double lo2 = g.lon0 + colm * g.hlon - kCellPadDeg;
    double hi2 = g.lon0 + (colm + 1) * g.hlon + kCellPadDeg;
    double e1 = fabs(qlon - lo2), e2 = fabs(qlon - hi2);
    double dlam = e1 > e2 ? e1 : e2;
    // sin(h) <= h - h^3/6 + h^5/120 for every h >= 0
    double s1 = sin_upper(dphi * kDeg2Rad / 2);
    double s2 = sin_upper(dlam * kDeg2Rad / 2);
    double a = s1 * s1 + cos_qlat * g.cosrow[g.G + row] * (s2 * s2);
    a *= (1.0 + kPruneSlack);
    return a > 1.0 ? 1.0 : a;
}

// Both bounds of one cell at once (the per-search path: the row / column decoding and the cell
// rectangle are shared).
IMOMD_HD IMOMD_NOINLINE void cell_bounds_a(const GraphDev& g, uint32_t c, double qlat, double qlon,
                                           double cos_qlat, double* lower, double* upper)
{
    if (!g.prune)
    {
        *lower = 0.0;
        *upper = 1.0;
        return;
    }
    int row = (int)(c / (uint32_t)g.G), colm = (int)(c % (uint32_t)g.G);
    double lo = g.lat0 + row * g.hlat - kCellPadDeg;
    double hi = g.lat0 + (row + 1) * g.hlat + kCellPadDeg;
    double lo2 = g.lon0 + colm * g.hlon - kCellPadDeg;
    double hi2 = g.lon0 + (colm + 1) * g.hlon + kCellPadDeg;
    double dphi = 0.0;
    if (qlat < lo) dphi = lo - qlat;
    else if (qlat > hi) dphi = qlat - hi;
    double dlam = 0.0;
    if (qlon < lo2) dlam = lo2 - qlon;
    else if (qlon > hi2) dlam = qlon - hi2;
    double s1 = sin_lower(dphi * kDeg2Rad / 2);
    double s2 = sin_lower(dlam * kDeg2Rad / 2);
    double a = s1 * s1 + cos_qlat * g.cosrow[row] * (s2 * s2);
    *lower = a * (1.0 - kPruneSlack);
    double d1 = fabs(qlat - lo), d2 = fabs(qlat - hi);
    double fphi = d1 > d2 ? d1 : d2;
    double e1 = fabs(qlon - lo2), e2 = fabs(qlon - hi2);
    double flam = e1 > e2 ? e1 : e2;
    double t1 = sin_upper(fphi * kDeg2Rad / 2);
    double t2 = sin_upper(flam * kDeg2Rad / 2);
    double b = t1 * t1 + cos_qlat * g.cosrow[g.G + row] * (t2 * t2);
    b *= (1.0 + kPruneSlack);
    *upper = b > 1.0 ? 1.0 : b;
}

// unit vector + hash cell of one vertex (graph upload)
IMOMD_HD inline VRec make_vrec(uint32_t v, double lat, double lon, int G, double lat0, double lon0,
                               double hlat, double hlon)
{
    double la = lat * kDeg2Rad, lo = lon * kDeg2Rad;
    double cl = cos(la), sl = sin(la), co = cos(lo), so = sin(lo);
    VRec r;
    r.x = cl * co;
    r.y = cl * so;
    r.z = sl;
    r.id = v;
    int row = (int)((lat - lat0) / hlat);
    int colm = (int)((lon - lon0) / hlon);
    row = row < 0 ? 0 : (row >= G ? G - 1 : row);
    colm = colm < 0 ? 0 : (colm >= G ? G - 1 : colm);
    r.cell = (uint32_t)(row * G + colm);
    return r;
}

// lower / upper bound (metres) of haversine(root, any point of cell c)
IMOMD_HD inline double lbd_of(const GraphDev& g, uint32_t root, uint32_t c)
{
    double rlat = g.lat[root], rlon = g.lon[root];
    double a = cell_lower_a(g, c, rlat, rlon, cos(rlat * kDeg2Rad));
    return dist_of_a(a) * (1.0 - kPruneSlack);
}

IMOMD_HD inline double ubd_of(const GraphDev& g, uint32_t root, uint32_t c)
{
    double rlat = g.lat[root], rlon = g.lon[root];
    double a = cell_upper_a(g, c, rlat, rlon, cos(rlat * kDeg2Rad));
    return dist_of_a(a) * (1.0 + kPruneSlack);
}

// lbd[slot][t][c] of a RESIDENT planner (slot == query): entry i of the flat table
IMOMD_HD inline double lbd_entry(const PlannerDev& P, size_t i)
{
    uint32_t cells = (uint32_t)P.g.G * (uint32_t)P.g.G;
    uint32_t c = (uint32_t)(i % cells);
    int t = (int)((i / cells) % P.K);
    int q = (int)(i / ((size_t)cells * P.K));
    return lbd_of(P.g, P.query[q].dest[t], c);
}

IMOMD_HD inline double ubd_entry(const PlannerDev& P, size_t i)
{
    uint32_t cells = (uint32_t)P.g.G * (uint32_t)P.g.G;
    uint32_t c = (uint32_t)(i % cells);
    int t = (int)((i / cells) % P.K);
    int q = (int)(i / ((size_t)cells * P.K));
    return ubd_of(P.g, P.query[q].dest[t], c);
}

// --------------------------------------------------------- per-tree views ---

// One record per (tree, vertex): {parent, frontier slot, cost} in the first 16 bytes (one
// 128-bit load answers "visited? at what cost? in the frontier?"), then the child row.
struct __attribute__((aligned(16))) U4
{
    uint32_t x, y, z, w;
};

struct THead
{
    uint32_t parent;   // NONE = not visited
    uint32_t fslot;    // frontier slot, NONE = not in the frontier
    double cost;
};

IMOMD_HD inline THead head_from(const U4& t)
{
    THead h;
    h.parent = t.x;
    h.fslot = t.y;
    unsigned long long bits = ((unsigned long long)t.w << 32) | (unsigned long long)t.z;
    memcpy(&h.cost, &bits, 8);
    return h;
}

// padded adjacency rows (see GraphDev)
IMOMD_HD inline uint32_t lb_entries(const PlannerDev& P)
{
    const uint32_t cells = (uint32_t)P.g.G * (uint32_t)P.g.G;
    return cells < kCtaLbCap ? cells : kCtaLbCap;
}
IMOMD_HD inline uint32_t row_begin(const GraphDev& g, uint32_t v) { return v * (uint32_t)g.width; }
IMOMD_HD inline uint32_t row_end(const GraphDev& g, uint32_t v) { return v * (uint32_t)g.width + g.deg[v]; }

struct TreeRef
{
    unsigned char* base;
    uint32_t stride;          // bytes per vertex record: 16 + 4 * width
    uint32_t* cell_head;
    uint32_t* cell_cnt;
    uint32_t* occ;        // [cells] occupied cells, then [cells] cell -> position
    VRec* rec;
    uint32_t* chunk_next;
    uint32_t* free_stack;

    IMOMD_HD THead* hp(uint32_t v) const
    {
        return reinterpret_cast<THead*>(base + (size_t)v * stride);
    }
    IMOMD_HD THead head(uint32_t v) const   // one 128-bit load
    {
        U4 t = *reinterpret_cast<const U4*>(base + (size_t)v * stride);
        THead h;
        h.parent = t.x;
        h.fslot = t.y;
        unsigned long long bits = ((unsigned long long)t.w << 32) | (unsigned long long)t.z;
        memcpy(&h.cost, &bits, 8);
        return h;
    }
    IMOMD_HD uint32_t parent(uint32_t v) const { return hp(v)->parent; }
    IMOMD_HD uint32_t fslot(uint32_t v) const { return hp(v)->fslot; }
    IMOMD_HD double cost(uint32_t v) const { return hp(v)->cost; }
    IMOMD_HD void set_parent(uint32_t v, uint32_t p) const { hp(v)->parent = p; }
    IMOMD_HD void set_fslot(uint32_t v, uint32_t f) const { hp(v)->fslot = f; }
    IMOMD_HD void set_cost(uint32_t v, double c) const { hp(v)->cost = c; }
    IMOMD_HD uint32_t* chl(uint32_t v) const
    {
        return reinterpret_cast<uint32_t*>(base + (size_t)v * stride + 16);
    }
};

IMOMD_HD inline TreeRef tree_ref(const PlannerDev& P, int q, int k)
{
    size_t qt = (size_t)q * P.K + k;
    size_t cells = (size_t)P.g.G * P.g.G;
    TreeRef t;
    t.stride = P.trec_stride;
    t.base = P.trec + qt * (size_t)P.g.n * P.trec_stride;
    t.cell_head = P.cell_head + qt * cells;
    t.cell_cnt = P.cell_cnt + qt * cells;
    t.occ = P.occ + qt * (2 * cells + P.dirty_words);
    t.rec = P.rec + qt * (size_t)P.chunks * IMOMD_CHUNK;
    t.chunk_next = P.chunk_next + qt * P.chunks;
    t.free_stack = P.free_stack + qt * P.chunks;
    IMOMD_ASSUME_GLOBAL(t.base);
    IMOMD_ASSUME_GLOBAL(t.cell_head);
    IMOMD_ASSUME_GLOBAL(t.cell_cnt);
    IMOMD_ASSUME_GLOBAL(t.occ);
    IMOMD_ASSUME_GLOBAL(t.rec);
    IMOMD_ASSUME_GLOBAL(t.chunk_next);
    IMOMD_ASSUME_GLOBAL(t.free_stack);
    return t;
}

// The view of tree k inside run_query: read back from the per-launch table in shared memory
// (64 bytes, four 128-bit loads) instead of eight 64-bit multiply-adds per call.
struct Shared;
IMOMD_HD inline TreeRef tree_ref_of(const PlannerDev& P, int q, int k, const Shared& S);

// every pointer of the planner descriptor is device-global memory
IMOMD_HD inline void assume_planner(const PlannerDev& P)
{
    IMOMD_ASSUME_GLOBAL(P.g.deg);
    IMOMD_ASSUME_GLOBAL(P.g.col);
    IMOMD_ASSUME_GLOBAL(P.g.w);
    IMOMD_ASSUME_GLOBAL(P.g.lat);
    IMOMD_ASSUME_GLOBAL(P.g.lon);
    IMOMD_ASSUME_GLOBAL(P.g.vrec);
    IMOMD_ASSUME_GLOBAL(P.g.coslat);
    IMOMD_ASSUME_GLOBAL(P.g.cosrow);
    IMOMD_ASSUME_GLOBAL(P.query);
    IMOMD_ASSUME_GLOBAL(P.trec);
    IMOMD_ASSUME_GLOBAL(P.lbd);
    IMOMD_ASSUME_GLOBAL(P.ubd);
    IMOMD_ASSUME_GLOBAL(P.arena);
    IMOMD_ASSUME_GLOBAL(P.clist);
    IMOMD_ASSUME_GLOBAL(P.log);
    IMOMD_ASSUME_GLOBAL(P.reports);
    IMOMD_ASSUME_GLOBAL(P.words);
    (void)P;
}

struct Best
{
    double v1;       // minimum
    uint32_t id1;    // its vertex (lowest id among exact ties)
    double v2;       // runner-up value (== v1 when the minimum is attained twice)
};

IMOMD_HD inline Best best_empty()
{
    Best b;
    b.v1 = INFINITY;
    b.id1 = IMOMD_NONE;
    b.v2 = INFINITY;
    return b;
}

IMOMD_HD inline void best_add(Best& b, double v, uint32_t id)
{
    if (v < b.v1 || (v == b.v1 && id < b.id1))
    {
        b.v2 = b.v1;
        b.v1 = v;
        b.id1 = id;
    }
    else if (v < b.v2)
    {
        b.v2 = v;
    }
}

IMOMD_HD inline Best best_merge(const Best& a, const Best& b)
{
    Best r;
    if (b.v1 < a.v1 || (b.v1 == a.v1 && b.id1 < a.id1))
    {
        r.v1 = b.v1;
        r.id1 = b.id1;
        r.v2 = a.v1 < b.v2 ? a.v1 : b.v2;
    }
    else
    {
        r.v1 = a.v1;
        r.id1 = a.id1;
        r.v2 = a.v2 < b.v1 ? a.v2 : b.v1;
    }
    return r;
}

// State shared by the threads of the CTA (shared memory on the device).
struct Shared
{
    // request posted by the sequential thread
    int32_t req;
    int32_t phase;
    int32_t it;            // iterations completed in this launch
    int32_t k;             // tree being expanded
    int32_t stop;          // the launch's time budget is spent: no further pass starts
    int32_t pad_stop;
    uint64_t t_start_ns;   // global timer at the start of the launch (time budget)
    // sample (selectRandomVertex_)
    uint64_t rand_id;
    int32_t sample_is_vertex;   // the sample's coordinates are exactly those of vertex rand_id
    int32_t pad_sample;
    double qlat, qlon;
    // search results
    Best result;
    int32_t ambiguous;
    uint32_t x_new;
    // rescan / heuristic-insert requests
    int32_t n_rescan;
    int32_t rescan_o[IMOMD_KMAX];
    int32_t n_ys;
    uint32_t ys[IMOMD_MAXDEG];
    double* h;             // [width * K] haversine of the inserted vertices to the K roots (behind the tree views)
    // work list of the pruned search
    uint32_t n_list;
    uint32_t n_cells, n_recs;   // statistics of the current search
    // connection gather: the other trees' (parent, cost) at conn_x (REQ_CONN / REQ_BFS)
    uint32_t conn_x;
    int32_t conn_pending;
    double conn_cx;        // cost of conn_x in the tree being expanded
    uint32_t conn_par[IMOMD_KMAX];
    double conn_cost[IMOMD_KMAX];
    // jump-point search in flight
    uint32_t jps_next;
    int32_t attach_mode;   // 1: imomd_attach (pseudo-tree merge): no sampling, JPS, pruning
    uint32_t attach_i, attach_n;   // position in / length of P.attach_list
    // rewire machine: explicit LIFO in the arena
    uint32_t rw_fp, rw_top;
    // cost propagation request / result (tree_t::updateCost)
    uint32_t bfs_node, bfs_lo, bfs_hi;
    int32_t bfs_valid;
    double bfs_cost;
    uint32_t lv_cur, lv_start, lv_n, lv_tail;
    int32_t lv_overflow;
    // speculative no-op check of the next rewire calls
    uint32_t chk_hi, chk_lo, chk_skip;
    uint32_t chk_noop, chk_slot;
    uint32_t chk_u, chk_e, chk_e1, chk_y, chk_py;   // the first call that re-parents
    double chk_nc;
    int32_t chk_valid;
    // hot header fields of the query (a shared-memory copy while a launch runs)
    QueryDev* Q;
    // the K tree views of the query, computed once per launch (null outside run_query)
    TreeRef* tref;
    // the query's K x K matrices, staged in shared memory for the duration of a launch
    double* mP;
    double* mD;
    double* mMHv;
    uint32_t* mMHi;
    uint32_t* mCN;
    // block reduction / scan scratch
    Best red[32];
    uint32_t scan[40];
    // scratch of the cooperating group in HBM: rewire arena (frames + update lists), chunk work
    // list of the pruned searches.  One group (the CTA) per query in the exact mode; one group (a
    // warp) per TREE in the batched mode, each with its own share.
    uint32_t* arena;
    uint32_t arena_cap;
    int32_t batch;              // 1: batched mode (cross-tree effects are logged, not applied)
    uint32_t* clist;
    BEvent* ev;                 // batched mode: selection-probability tests of the tree being expanded
    uint32_t ev_n;              //   entries written
    uint32_t exp_n;             //   expansions performed in this batch
    uint32_t* dirty;            // batched mode: vertices queued for the rewiring relaxation of the step
    uint32_t dirty_n, dirty_cap;
    uint32_t* chk;              // batched mode: vertices whose connection check is due after the step
    uint32_t chk_n, chk_cap;
    // counters of this group, folded into the query header when the launch ends
    uint64_t stat[8];
    uint64_t gprof[8];          // batched mode: cycles of this group by request type (7 = serial steps)
    int64_t near_ties, unresolved_ties, certified_ties;
    // exact certification of a comparison the device arithmetic cannot decide (Cta::certify):
    // candidate vertices, the coordinate pairs handed to the host, glibc's answers
    uint32_t tie_n;
    uint32_t mh_tie_mask;       // heuristic inserts: bit o = the comparison for tree o needs exact values
    uint32_t tie_id[IMOMD_TIE_IDS];
    double tie_in[IMOMD_TIE_PAIRS][4];
    double tie_out[IMOMD_TIE_PAIRS];
};

// Thread 0: decide an arg-min among the n candidate vertices S.tie_id[0..n) with the REFERENCE's
// arithmetic: the objective hav(v, a) (+ hav(v, b) when `two`) is evaluated by the host with glibc
// (Cta::certify), the minimum is taken with the reference's strict `<` over ascending vertex ids
// (lowest id among exact ties, north_star).  Returns the position of the winner, or -1 when no
// host service is attached / the candidates do not fit (the caller keeps the device decision and
// counts it as unresolved).  *exact_tie: the minimum is attained by two vertices bit for bit.
template <class Cta>
IMOMD_HD inline int certify_argmin(Cta& cta, const GraphDev& g, Shared& S, uint32_t n, bool two, double alat,
                                   double alon, double blat, double blon, double* vals, bool* exact_tie)
{
    *exact_tie = false;
    const uint32_t per = two ? 2u : 1u;
    if (n == 0 || n > IMOMD_TIE_IDS) return -1;
    for (uint32_t i = 1; i < n; ++i)     // ascending ids
    {
        uint32_t x = S.tie_id[i];
        uint32_t j = i;
        while (j > 0 && S.tie_id[j - 1] > x)
        {
            S.tie_id[j] = S.tie_id[j - 1];
            --j;
        }
        S.tie_id[j] = x;
    }
    // the mailbox takes IMOMD_TIE_PAIRS coordinate pairs per request: long candidate lists (vertices
    // along the segment between two roots all have the same two-root heuristic) go in rounds
    const uint32_t round = IMOMD_TIE_PAIRS / per;
    for (uint32_t i0 = 0; i0 < n; i0 += round)
    {
        const uint32_t m = n - i0 < round ? n - i0 : round;
        for (uint32_t i = 0; i < m; ++i)
        {
            const uint32_t v = S.tie_id[i0 + i];
            double* a = S.tie_in[per * i];
            a[0] = g.lat[v]; a[1] = g.lon[v]; a[2] = alat; a[3] = alon;
            if (two)
            {
                double* b = S.tie_in[per * i + 1];
                b[0] = g.lat[v]; b[1] = g.lon[v]; b[2] = blat; b[3] = blon;
            }
        }
        if (!cta.certify((int)(m * per), S.tie_in, S.tie_out)) return -1;
        for (uint32_t i = 0; i < m; ++i)
            vals[i0 + i] = two ? S.tie_out[2 * i] + S.tie_out[2 * i + 1] : S.tie_out[i];
    }
    int best = 0;
    for (uint32_t i = 1; i < n; ++i)
        if (vals[i] < vals[best]) best = (int)i;
    for (uint32_t i = 0; i < n; ++i)
        if ((int)i != best && vals[i] == vals[best]) *exact_tie = true;
    return best;
}

// --------------------------------------------------- frontier hash (serial) ---

// Streaming planners keep one bit per vertex record of a tree, set when the record is first
// written (the vertex enters the frontier or the tree): recycling a slot and hashing its trees
// then walk n / 32 words per tree instead of n records (reset_slot, tree_hash).  The bitmap sits
// behind the tree's occ lists.  Different trees never share a word; a tree's records are written by
// one thread at a time in every mode, the atomic only keeps the update a single memory operation.
IMOMD_HD inline void dirty_mark(const PlannerDev& P, const TreeRef& T, uint32_t v)
{
    if (P.dirty_words == 0) return;
    uint32_t* w = T.occ + 2 * (size_t)P.g.G * P.g.G + (v >> 5);
#ifdef __CUDA_ARCH__
    atomicOr(w, 1u << (v & 31));
#else
    *w |= 1u << (v & 31);
#endif
}

IMOMD_HD inline bool fr_insert(const PlannerDev& P, QueryDev* Q, const TreeRef& T, int k,
                               uint32_t v)
{
    if (T.fslot(v) != IMOMD_NONE) return true;
    VRec r = P.g.vrec[v];
    uint32_t c = r.cell;
    uint32_t cnt = T.cell_cnt[c];
    uint32_t off = cnt % IMOMD_CHUNK;
    uint32_t head;
    if (off == 0)
    {
        if (Q->free_top[k] == 0)
        {
            Q->status = 2;   // IMOMD_RUN_FRONTIER_FULL
            return false;
        }
        head = T.free_stack[--Q->free_top[k]];
        T.chunk_next[head] = cnt ? T.cell_head[c] : IMOMD_NONE;
        T.cell_head[c] = head;
        if (cnt == 0)
        {
            uint32_t pos = Q->n_occ[k]++;
            T.occ[pos] = c;
            T.occ[(size_t)P.g.G * P.g.G + c] = pos;
        }
    }
    else
    {
        head = T.cell_head[c];
    }
    uint32_t slot = head * IMOMD_CHUNK + off;
    T.rec[slot] = r;
    T.set_fslot(v, slot);
    dirty_mark(P, T, v);
    T.cell_cnt[c] = cnt + 1;
    Q->frontier_count[k] += 1;
    return true;
}

IMOMD_HD inline void fr_erase(const PlannerDev& P, QueryDev* Q, const TreeRef& T, int k,
                              uint32_t v)
{
    uint32_t s = T.fslot(v);
    if (s == IMOMD_NONE) return;
    uint32_t c = P.g.vrec[v].cell;
    uint32_t cnt = T.cell_cnt[c];
    uint32_t head = T.cell_head[c];
    uint32_t off = (cnt - 1) % IMOMD_CHUNK;
    uint32_t last = head * IMOMD_CHUNK + off;
    if (last != s)
    {
        VRec m = T.rec[last];
        T.rec[s] = m;
        T.set_fslot(m.id, s);
    }
    T.set_fslot(v, IMOMD_NONE);
    T.cell_cnt[c] = cnt - 1;
    Q->frontier_count[k] -= 1;
    if (off == 0)
    {
        T.cell_head[c] = T.chunk_next[head];
        T.free_stack[Q->free_top[k]++] = head;
        if (cnt == 1)
        {
            size_t cells = (size_t)P.g.G * P.g.G;
            uint32_t pos = T.occ[cells + c];
            uint32_t last = T.occ[--Q->n_occ[k]];
            T.occ[pos] = last;
            T.occ[cells + last] = pos;
        }
    }
}

// ---------------------------------------- children lists (container order) ---
// unordered_set<size_t> with <= 13 elements: a new key goes to the head of its
// bucket's run (bucket = key % 13) or, for an empty bucket, to the list head
// (SURVEY.md 8a-7).  Each vertex owns a fixed row of `width` slots (width = power of
// two >= max degree, children of p are neighbours of p), NONE-terminated, so a row is
// read or written with one or a few 128-bit accesses and needs no row_ptr lookup.

template <int W>
struct ChlRow
{
    uint32_t v[W];
};

template <int W>
IMOMD_HD inline ChlRow<W> chl_load(const uint32_t* rowp)
{
    ChlRow<W> r;
    const U4* src = reinterpret_cast<const U4*>(rowp);
#pragma unroll
    for (int i = 0; i < W / 4; ++i)
    {
        U4 t = src[i];
        r.v[4 * i] = t.x;
        r.v[4 * i + 1] = t.y;
        r.v[4 * i + 2] = t.z;
        r.v[4 * i + 3] = t.w;
    }
    return r;
}

template <int W>
IMOMD_HD inline void chl_store(uint32_t* rowp, const ChlRow<W>& r)
{
    U4* dst = reinterpret_cast<U4*>(rowp);
#pragma unroll
    for (int i = 0; i < W / 4; ++i)
    {
        U4 t;
        t.x = r.v[4 * i];
        t.y = r.v[4 * i + 1];
        t.z = r.v[4 * i + 2];
        t.w = r.v[4 * i + 3];
        dst[i] = t;
    }
}

template <int W>
IMOMD_HD inline void chl_row_insert(ChlRow<W>& r, uint32_t x)
{
    int n = 0, at = -1;
#pragma unroll
    for (int i = 0; i < W; ++i)
    {
        if (r.v[i] != IMOMD_NONE)
        {
            if (r.v[i] == x) return;
            if (at < 0 && r.v[i] % 13 == x % 13) at = i;
            n = i + 1;
        }
    }
    if (at < 0) at = 0;
    if (n >= W) return;
#pragma unroll
    for (int i = W - 1; i > 0; --i)
    {
        if (i > at) r.v[i] = r.v[i - 1];
    }
#pragma unroll
    for (int i = 0; i < W; ++i)
    {
        if (i == at) r.v[i] = x;
    }
}

template <int W>
IMOMD_HD inline void chl_row_erase(ChlRow<W>& r, uint32_t x)
{
    int at = -1;
#pragma unroll
    for (int i = 0; i < W; ++i)
    {
        if (r.v[i] == x) at = i;
    }
    if (at < 0) return;
#pragma unroll
    for (int i = 0; i < W - 1; ++i)
    {
        if (i >= at) r.v[i] = r.v[i + 1];
    }
    r.v[W - 1] = IMOMD_NONE;
}

template <int W>
IMOMD_HD inline void chl_insert_w(uint32_t* rowp, uint32_t x)
{
    ChlRow<W> r = chl_load<W>(rowp);
    int n = 0, at = -1;
#pragma unroll
    for (int i = 0; i < W; ++i)
    {
        if (r.v[i] != IMOMD_NONE)
        {
            if (r.v[i] == x) return;
            if (at < 0 && r.v[i] % 13 == x % 13) at = i;
            n = i + 1;
        }
    }
    if (at < 0) at = 0;
    if (n >= W) return;   // cannot happen: children are distinct neighbours, W >= degree
#pragma unroll
    for (int i = W - 1; i > 0; --i)
    {
        if (i > at) r.v[i] = r.v[i - 1];
    }
#pragma unroll
    for (int i = 0; i < W; ++i)
    {
        if (i == at) r.v[i] = x;
    }
    chl_store<W>(rowp, r);
}

template <int W>
IMOMD_HD inline void chl_erase_w(uint32_t* rowp, uint32_t x)
{
    ChlRow<W> r = chl_load<W>(rowp);
    int at = -1;
#pragma unroll
    for (int i = 0; i < W; ++i)
    {
        if (r.v[i] == x) at = i;
    }
    if (at < 0) return;
#pragma unroll
    for (int i = 0; i < W - 1; ++i)
    {
        if (i >= at) r.v[i] = r.v[i + 1];
    }
    r.v[W - 1] = IMOMD_NONE;
    chl_store<W>(rowp, r);
}

template <int W>
IMOMD_HD inline void chl_assign_single_w(uint32_t* rowp, uint32_t x)
{
    ChlRow<W> r;
#pragma unroll
    for (int i = 0; i < W; ++i) r.v[i] = IMOMD_NONE;
    r.v[0] = x;
    chl_store<W>(rowp, r);
}

// (*connection_ptr_)[a][b]; the graph was validated symmetric at upload
IMOMD_HD inline double edge_weight(const GraphDev& g, uint32_t a, uint32_t b)
{
    uint32_t e0 = row_begin(g, a), e1 = row_end(g, a);
    for (uint32_t e = e0; e < e1; ++e)
    {
        if (g.col[e] == b) return g.w[e];
    }
    return 0.0;
}

// ------------------------------------------------- random sample stream ----
// libstdc++ on top of raw mt19937 words.

struct WordStream
{
    const uint32_t* words;
    uint64_t cursor;
    IMOMD_HD uint32_t next() { return words[cursor++]; }
};

// std::generate_canonical<double,53>: (w0 + w1 * 2^32) / 2^64, clamped below 1
IMOMD_HD inline double ws_canonical(WordStream& ws)
{
    double sum = 0.0;
    double tmp = 1.0;
    sum += (double)ws.next() * tmp;
    tmp *= 4294967296.0;
    sum += (double)ws.next() * tmp;
    tmp *= 4294967296.0;
    double ret = sum / tmp;
    if (ret >= 1.0) ret = 0.99999999999999988897769753748;   // nextafter(1.0, 0.0)
    return ret;
}

IMOMD_HD inline double ws_uniform_real(WordStream& ws, double a, double b)
{
    return ws_canonical(ws) * (b - a) + a;
}

// std::uniform_int_distribution over [0, urange], urange < 2^32 - 1 (Lemire)
IMOMD_HD inline uint32_t ws_uniform_u32(WordStream& ws, uint32_t urange)
{
    if (urange == 0xFFFFFFFFu) return ws.next();
    uint32_t range = urange + 1u;
    uint64_t product = (uint64_t)ws.next() * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range)
    {
        uint32_t threshold = (0u - range) % range;
        while (low < threshold)
        {
            product = (uint64_t)ws.next() * (uint64_t)range;
            low = (uint32_t)product;
        }
    }
    return (uint32_t)(product >> 32);
}

// ===================================================== sequential stages =====

IMOMD_HD inline TreeRef tree_ref_of(const PlannerDev& P, int q, int k, const Shared& S)
{
    const TreeRef* table = S.tref;
    if (table == nullptr) return tree_ref(P, q, k);
    IMOMD_ASSUME_SHARED(table);
    TreeRef t = table[k];
    IMOMD_ASSUME_GLOBAL(t.base);
    IMOMD_ASSUME_GLOBAL(t.cell_head);
    IMOMD_ASSUME_GLOBAL(t.cell_cnt);
    IMOMD_ASSUME_GLOBAL(t.occ);
    IMOMD_ASSUME_GLOBAL(t.rec);
    IMOMD_ASSUME_GLOBAL(t.chunk_next);
    IMOMD_ASSUME_GLOBAL(t.free_stack);
    return t;
}

// The staged K x K matrices live in shared memory; their pointers are re-read from `Shared`, so
// the address space is asserted at every use (see IMOMD_ASSUME_SHARED).
IMOMD_HD inline double* sm_P(const Shared& S) { double* p = S.mP; IMOMD_ASSUME_SHARED(p); return p; }
IMOMD_HD inline double* sm_D(const Shared& S) { double* p = S.mD; IMOMD_ASSUME_SHARED(p); return p; }
IMOMD_HD inline double* sm_MHv(const Shared& S) { double* p = S.mMHv; IMOMD_ASSUME_SHARED(p); return p; }
IMOMD_HD inline uint32_t* sm_MHi(const Shared& S) { uint32_t* p = S.mMHi; IMOMD_ASSUME_SHARED(p); return p; }
IMOMD_HD inline uint32_t* sm_CN(const Shared& S) { uint32_t* p = S.mCN; IMOMD_ASSUME_SHARED(p); return p; }

template <int KW>
struct Seq
{
    const PlannerDev& P;
    QueryDev* Q;
    int q;
    Shared& S;
    uint32_t* arena;

    IMOMD_HD Seq(const PlannerDev& P_, int q_, Shared& S_)
        : P(P_), Q(S_.Q), q(q_), S(S_), arena(S_.arena)
    {
        assume_planner(P);
        IMOMD_ASSUME_SHARED(&S);
        IMOMD_ASSUME_SHARED(Q);      // the launch-staged copy of the hot header
        IMOMD_ASSUME_SHARED(S.mP);
        IMOMD_ASSUME_SHARED(S.mD);
        IMOMD_ASSUME_SHARED(S.mMHv);
        IMOMD_ASSUME_SHARED(S.mMHi);
        IMOMD_ASSUME_SHARED(S.mCN);
        IMOMD_ASSUME_GLOBAL(arena);
    }

    IMOMD_HD uint32_t deg(uint32_t v) const { return P.g.deg[v]; }

    // selectRandomVertex_, src/imomd_rrt_star.cpp:358-400
    IMOMD_HD void sample(int k, const TreeRef& T)
    {
        Q->word_cursor = sample_at(k, T, Q->word_cursor);
    }

    // the same, reading the injected stream from `cursor`; returns the cursor behind the sample
    IMOMD_HD uint64_t sample_at(int k, const TreeRef& T, uint64_t cursor)
    {
        const GraphDev& g = P.g;
        int K = Q->K;
        WordStream ws;
        ws.words = P.words;
        ws.cursor = cursor;
        if (ws_uniform_real(ws, 0, 1) < P.goal_bias)
        {
            double rnd = ws_uniform_real(ws, 0, 1);
            int i = 0;
            while (rnd >= 0)
            {
                if (i >= K)
                {
                    Q->status = 4;   // IMOMD_RUN_CDF_OVERRUN
                    break;
                }
                rnd -= sm_P(S)[k * K + i++];
            }
            uint32_t random_dest = Q->view_root[i - 1];
            double rlat = g.lat[random_dest];
            double rlon = g.lon[random_dest];
            S.sample_is_vertex = 1;
            if (T.parent(random_dest) != IMOMD_NONE)
            {
                S.sample_is_vertex = 0;   // the sample is moved towards the root (:383-390)
                double ratio = ws_uniform_real(ws, 0.0, 0.5);
                uint32_t root = Q->view_root[k];
                rlat *= ratio;
                rlon *= ratio;
                rlat += (1.0 - ratio) * g.lat[root];
                rlon += (1.0 - ratio) * g.lon[root];
            }
            S.rand_id = random_dest;
            S.qlat = rlat;
            S.qlon = rlon;
        }
        else
        {
            uint32_t v = ws_uniform_u32(ws, g.n - 1);
            S.sample_is_vertex = 1;
            S.rand_id = v;
            S.qlat = g.lat[v];
            S.qlon = g.lon[v];
        }
        return ws.cursor;
    }

    // ---- batched mode: what an expansion does beyond its own attach is queued ------------------
    // the rewiring of x_new (relaxation after the step), the connection checks of the attached
    // vertices (after the relaxation, with the final costs), and updateSelectionProbability with
    // the heuristic entry of that moment, in program order.
    IMOMD_HD void queue_rewire(uint32_t x)
    {
        if (S.dirty_n >= S.dirty_cap)
        {
            Q->status = 3;   // IMOMD_RUN_ARENA_FULL
            return;
        }
        S.dirty[S.dirty_n++] = x;
    }
    IMOMD_HD void queue_check(uint32_t x)
    {
        if (S.chk_n >= S.chk_cap)
        {
            Q->status = 3;
            return;
        }
        S.chk[S.chk_n++] = x;
    }
    IMOMD_HD void log_event(int k, uint32_t kind, uint32_t x, uint32_t aux, double val, uint32_t extra)
    {
        uint32_t i = S.ev_n;
        if (i >= P.ev_cap)
        {
            Q->status = 3;   // IMOMD_RUN_ARENA_FULL: event log of the batch exhausted
            return;
        }
        BEvent* e = S.ev + i;
        e->x = x;
        e->aux = aux;
        e->val = val;
        e->extra = extra;
        e->kind = kind;
        S.ev_n = i + 1;
    }

    // which (k, o) min-heuristic holders equal x  (updateExpandables :503-512)
    IMOMD_HD void collect_rescans(int k, uint32_t x)
    {
        int K = Q->K;
        S.n_rescan = 0;
        for (int o = 0; o < K; ++o)
        {
            if (o == k) continue;
            if (sm_MHi(S)[k * K + o] == x) S.rescan_o[S.n_rescan++] = o;
        }
    }

    // updateConnectionTree_, src/imomd_rrt_star.cpp:645-706.  The other trees' parent
    // and cost at x were gathered into S.conn_par / S.conn_cost by the CTA
    // (svc_conn_gather); they cannot change while tree k is being expanded.
    IMOMD_HD void apply_connection(int k, const TreeRef& T, uint32_t x, double cx)
    {
        int K = Q->K;
        S.stat[6] += 1;
        for (int o = 0; o < K; ++o)
        {
            if (o == k) continue;
            if (S.conn_par[o] == IMOMD_NONE) continue;
            int set_idx = (o > k) ? o * (o - 1) / 2 + k : k * (k - 1) / 2 + o;
            if (!Q->contact[set_idx])
            {
                Q->contact[set_idx] = 1;
                if (P.pseudo) log_connection((uint32_t)set_idx, x);
                connect_two_trees(k, o);
            }
            double distance = cx + S.conn_cost[o];
            if (distance < sm_D(S)[k * K + o] - 0.1)
            {
                sm_D(S)[k * K + o] = distance;
                sm_D(S)[o * K + k] = distance;
                sm_CN(S)[k * K + o] = x;
                sm_CN(S)[o * K + k] = x;
                Q->dm_updated = 1;
            }
            if (P.pseudo)
            {
                TreeRef O = tree_ref_of(P, q, o, S);
                uint32_t x_parent = T.parent(x);
                uint32_t opp = O.parent(x_parent);
                if (opp == IMOMD_NONE || (opp != x && S.conn_par[o] != x_parent))
                {
                    log_connection((uint32_t)set_idx, x);
                }
            }
        }
    }

    // pseudo mode: the order in which a tree visits its vertices is the insertion order of
    // the reference's `tree.parent` map, which mergeTwoTree_ iterates (:860)
    IMOMD_HD void log_visit(int k, uint32_t x)
    {
        if (!P.pseudo) return;
        P.visit_log[((size_t)q * P.K + k) * P.g.n + (Q->tree_size[k] - 1)] = x;
    }

    IMOMD_HD void log_connection(uint32_t set_idx, uint32_t x)
    {
        uint32_t* lg = P.log + (size_t)q * 2 * P.log_entries;
        uint32_t i = Q->log_count;
        if (i < P.log_entries)
        {
            lg[2 * i] = set_idx;
            lg[2 * i + 1] = x;
        }
        Q->log_count = i + 1;   // an overflow is visible to the host as count > capacity
    }

    // connectTwoTree_, src/imomd_rrt_star.cpp:708-742 (flat relabel: same labels)
    IMOMD_HD void connect_two_trees(int a, int b)
    {
        if (Q->connected) return;
        int K = Q->K;
        int lo = Q->ds_parent[a], hi = Q->ds_parent[b];
        if (lo == hi) return;
        if (hi < lo)
        {
            int t = lo;
            lo = hi;
            hi = t;
        }
        for (int i = 0; i < K; ++i)
        {
            if (Q->ds_parent[i] == hi) Q->ds_parent[i] = lo;
        }
        int all = 1;
        for (int i = 1; i < K; ++i)
        {
            if (Q->ds_parent[i] != Q->ds_parent[0]) all = 0;
        }
        Q->connected = all;
    }

    // One hop of the jump-point search of steerNewNode_ (src/imomd_rrt_star.cpp:427-459).
    // Returns true when a chain vertex was attached (:443-445 / :453-455): its
    // connection check (:446 / :456) is then pending and S.jps_next holds the vertex the
    // search continues from.  Returns false when the search is over at S.x_new.
    IMOMD_HD bool jps_hop(int k, const TreeRef& T)
    {
        const GraphDev& g = P.g;
        uint32_t x_new = S.x_new;
        if (deg(x_new) != 2) return false;
        uint32_t e = row_begin(g, x_new);
        uint32_t x_1 = g.col[e];
        uint32_t x_2 = g.col[e + 1];
        uint32_t via, next;
        if (T.parent(x_1) != IMOMD_NONE)
        {
            if (T.parent(x_2) != IMOMD_NONE) return false;   // both sides visited: done
            via = x_1;
            next = x_2;
        }
        else
        {
            via = x_2;
            next = x_1;
        }
        if (T.parent(x_new) == IMOMD_NONE)
        {
            Q->tree_size[k] += 1;
            log_visit(k, x_new);
            dirty_mark(P, T, x_new);
        }
        T.set_parent(x_new, via);
        if (!S.batch) chl_assign_single_w<KW>(T.chl(via), x_new);
        double cnew = T.cost(via) + edge_weight(g, via, x_new);
        T.set_cost(x_new, cnew);
        S.conn_x = x_new;
        S.conn_cx = cnew;
        S.jps_next = next;
        return true;
    }

    // connectNewNode_, src/imomd_rrt_star.cpp:465-496.  All neighbour records are
    // requested before the first one is looked at (one round trip instead of one per
    // neighbour); the decision loop then runs in adjacency order with strict `<`.
    template <int W>
    IMOMD_HD void connect_new_node_w(int k, const TreeRef& T, uint32_t x_new)
    {
        const GraphDev& g = P.g;
        uint32_t e0 = row_begin(g, x_new), e1 = row_end(g, x_new);
        uint32_t ys[W];
        double ws[W];
        THead hs[W];
#pragma unroll
        for (int i = 0; i < W; ++i)
        {
            if (e0 + i < e1)
            {
                ys[i] = g.col[e0 + i];
                ws[i] = g.w[e0 + i];
            }
        }
#pragma unroll
        for (int i = 0; i < W; ++i)
        {
            if (e0 + i < e1) hs[i] = T.head(ys[i]);
        }
        uint32_t x_parent = IMOMD_NONE;
        double min_cost = INFINITY;
#pragma unroll
        for (int i = 0; i < W; ++i)
        {
            if (e0 + i < e1 && hs[i].parent != IMOMD_NONE)
            {
                double c = hs[i].cost + ws[i];
                if (c < min_cost)
                {
                    x_parent = ys[i];
                    min_cost = c;
                }
            }
        }
        if (x_parent != IMOMD_NONE)
        {
            if (T.parent(x_new) == IMOMD_NONE)
            {
                Q->tree_size[k] += 1;
                log_visit(k, x_new);
                dirty_mark(P, T, x_new);
            }
            T.set_parent(x_new, x_parent);
            if (!S.batch) chl_insert_w<KW>(T.chl(x_parent), x_new);
            T.set_cost(x_new, min_cost);
        }
    }

    IMOMD_HD void connect_new_node(int k, const TreeRef& T, uint32_t x_new)
    {
        connect_new_node_w<KW>(k, T, x_new);
    }

    // neighbour part of updateExpandables, src/imomd_rrt_star.cpp:533-558 (the
    // heuristic updates of the listed vertices run as the REQ_MHINS service)
    template <int W>
    IMOMD_HD void insert_neighbours_w(int k, const TreeRef& T, uint32_t x_new)
    {
        const GraphDev& g = P.g;
        S.n_ys = 0;
        uint32_t e0 = row_begin(g, x_new), e1 = row_end(g, x_new);
        uint32_t ys[W];
        uint32_t par[W];
#pragma unroll
        for (int i = 0; i < W; ++i)
        {
            if (e0 + i < e1) ys[i] = g.col[e0 + i];
        }
#pragma unroll
        for (int i = 0; i < W; ++i)
        {
            if (e0 + i < e1) par[i] = T.parent(ys[i]);
        }
#pragma unroll
        for (int i = 0; i < W; ++i)
        {
            if (e0 + i < e1 && par[i] == IMOMD_NONE)
            {
                if (!fr_insert(P, Q, T, k, ys[i])) return;
                S.ys[S.n_ys++] = ys[i];
            }
        }
    }

    IMOMD_HD void insert_neighbours(int k, const TreeRef& T, uint32_t x_new)
    {
        insert_neighbours_w<KW>(k, T, x_new);
    }

    // ---- rewireTree_, src/imomd_rrt_star.cpp:561-600 ------------------------------
    // The recursion is unrolled onto an explicit LIFO in the arena.  A frame is
    // {prev, x, e, end, skip, lo, cur}; above the frame that owns it sits the
    // `updated_nodes` list [lo, cur) of the active re-parenting, consumed from the back
    // (:589-596).  Two steps are handed to the whole CTA: the cost propagation
    // (tree_t::updateCost, REQ_BFS) and the speculative examination of the next calls of
    // the list, almost all of which change nothing (REQ_CHECK).
    static constexpr uint32_t FR = 7;

    IMOMD_HD bool push_frame(uint32_t x, uint32_t e) { return push_frame(x, e, row_end(P.g, x)); }

    IMOMD_HD bool push_frame(uint32_t x, uint32_t e, uint32_t e_end)
    {
        uint32_t top = S.rw_top;
        if (top + FR > S.arena_cap)
        {
            Q->status = 3;   // IMOMD_RUN_ARENA_FULL
            return false;
        }
        S.stat[2] += 1;
        uint32_t* f = arena + top;
        f[0] = S.rw_fp;
        f[1] = x;
        f[2] = e;
        f[3] = e_end;
        f[4] = IMOMD_NONE;
        f[5] = IMOMD_NONE;
        f[6] = 0;
        S.rw_fp = top;
        S.rw_top = top + FR;
        return true;
    }

    template <int W>
    IMOMD_HD void scan_frame(const TreeRef& T, const uint32_t* f, uint32_t& hit_e, uint32_t& hit_y,
                             uint32_t& hit_py, double& hit_nc)
    {
        const GraphDev& g = P.g;
        const uint32_t x = f[1], e0 = f[2], e1 = f[3];
        uint32_t ys[W];
        double ws[W];
        THead hs[W];
#pragma unroll
        for (int i = 0; i < W; ++i)
        {
            if (e0 + i < e1)
            {
                ys[i] = g.col[e0 + i];
                ws[i] = g.w[e0 + i];
            }
        }
        THead hx = T.head(x);
#pragma unroll
        for (int i = 0; i < W; ++i)
        {
            if (e0 + i < e1) hs[i] = T.head(ys[i]);
        }
        S.stat[3] += e1 - e0;
#pragma unroll
        for (int i = 0; i < W; ++i)
        {
            if (hit_e == IMOMD_NONE && e0 + i < e1 && hs[i].parent != IMOMD_NONE && ys[i] != hx.parent)
            {
                double nc = hx.cost + ws[i];
                if (hs[i].cost > nc)
                {
                    hit_e = e0 + i;
                    hit_y = ys[i];
                    hit_py = hs[i].parent;
                    hit_nc = nc;
                }
            }
        }
    }

    // y leaves py's child row and joins x's (:579-581); both rows are fetched together
    template <int W>
    IMOMD_HD void move_child_w(const TreeRef& T, uint32_t y, uint32_t py, uint32_t x)
    {
        if (py == x)
        {
            chl_erase_w<W>(T.chl(py), y);
            chl_insert_w<W>(T.chl(x), y);
            return;
        }
        ChlRow<W> from = chl_load<W>(T.chl(py));
        ChlRow<W> to = chl_load<W>(T.chl(x));
        chl_row_erase<W>(from, y);
        chl_row_insert<W>(to, y);
        chl_store<W>(T.chl(py), from);
        chl_store<W>(T.chl(x), to);
    }

    IMOMD_HD void reparent(const TreeRef& T, uint32_t y, uint32_t py, uint32_t x, double nc)
    {
        S.stat[4] += 1;
        move_child_w<KW>(T, y, py, x);
        T.set_parent(y, x);
        S.bfs_node = y;
        S.bfs_cost = nc;
        S.bfs_lo = S.rw_top;
        S.req = REQ_BFS;
    }

    // returns true when the cascade is finished, false when a service was requested
    IMOMD_HD bool rewire_run(int k, const TreeRef& T)
    {
        const GraphDev& g = P.g;
        while (S.rw_fp != IMOMD_NONE)
        {
            if (Q->status != 0) return true;
            uint32_t* f = arena + S.rw_fp;
            if (f[5] != IMOMD_NONE)
            {
                // walking updated_nodes in reverse
                if (S.chk_valid)
                {
                    S.chk_valid = 0;
                    S.stat[2] += S.chk_noop;
                    f[6] -= S.chk_noop;          // calls that would have changed nothing
                    if (S.chk_slot != IMOMD_NONE)
                    {
                        // the examined call rewireTree_(u) re-parents y at slot chk_e: enter
                        // its frame just behind that slot and do the re-parenting (:574-584)
                        --f[6];
                        uint32_t u = S.chk_u;
                        if (!push_frame(u, S.chk_e + 1, S.chk_e1)) return true;
                        reparent(T, S.chk_y, S.chk_py, u, S.chk_nc);
                        return false;
                    }
                    continue;
                }
                if (f[6] <= f[5])
                {
                    S.rw_top = f[5];             // list exhausted: release it
                    f[5] = IMOMD_NONE;
                    continue;
                }
                S.chk_hi = f[6];
                S.chk_lo = f[5];
                S.chk_skip = f[4];
                S.req = REQ_CHECK;
                return false;
            }
            if (S.bfs_valid)
            {
                // the re-parenting started below has its updated_nodes list now
                S.bfs_valid = 0;
                if (S.bfs_hi == IMOMD_NONE)
                {
                    Q->status = 3;
                    return true;
                }
                apply_connection(k, T, S.bfs_node, S.bfs_cost);      // :586
                f[4] = S.bfs_node;
                f[5] = S.rw_top;
                f[6] = S.bfs_hi;
                S.rw_top = S.bfs_hi;
                continue;
            }
            if (f[2] >= f[3])
            {
                S.rw_top = S.rw_fp;
                S.rw_fp = f[0];
                continue;
            }
            // serial scan of x's adjacency (:565-572): the remaining slots are fetched
            // together, the first one that re-parents is acted on
            uint32_t hit_e = IMOMD_NONE, hit_y = IMOMD_NONE, hit_py = IMOMD_NONE;
            double hit_nc = 0.0;
            scan_frame<KW>(T, f, hit_e, hit_y, hit_py, hit_nc);
            if (hit_e == IMOMD_NONE)
            {
                f[2] = f[3];
                continue;
            }
            f[2] = hit_e + 1;
            reparent(T, hit_y, hit_py, f[1], hit_nc);
            return false;
        }
        return true;
    }

    // updateSelectionProbability, src/imomd_rrt_star.cpp:602-643
    template <class Cta>
    IMOMD_HD void update_selection_probability(Cta& cta, int k)
    {
        int K = Q->K;
        int other = K;
        for (int i = 0; i < K; ++i)
        {
            if ((uint64_t)Q->dest[i] == S.rand_id)
            {
                other = i;
                break;
            }
        }
        if (other == K) return;
        selection_probability_apply(cta, k, other, sm_MHv(S)[k * K + other], sm_MHi(S)[k * K + other]);
    }

    // the test and the pruning of :617-642 for the pair (k, other), given the min-heuristic entry
    template <class Cta>
    IMOMD_HD void selection_probability_apply(Cta& cta, int k, int other, double mh, uint32_t mh_holder)
    {
        int K = Q->K;
        const double dd = sm_D(S)[k * K + other];
        bool prune = mh > dd;
        if (mh_holder != IMOMD_NONE && dd < INFINITY && fabs(mh - dd) <= P.tie_rel * dd)
        {
            // the heuristic value is CUDA libm's: this close to the distance the comparison is
            // redone on glibc's value of the holder
            const GraphDev& g = P.g;
            const uint32_t v = mh_holder, rk = Q->view_root[k], ro = Q->view_root[other];
            double* a = S.tie_in[0];
            double* b = S.tie_in[1];
            a[0] = g.lat[v]; a[1] = g.lon[v]; a[2] = g.lat[rk]; a[3] = g.lon[rk];
            b[0] = g.lat[v]; b[1] = g.lon[v]; b[2] = g.lat[ro]; b[3] = g.lon[ro];
            if (cta.certify(2, S.tie_in, S.tie_out))
            {
                prune = (S.tie_out[0] + S.tie_out[1]) > dd;
                S.certified_ties += 1;
            }
            else
            {
                S.unresolved_ties += 1;
                double* ti = Q->tie_info;
                ti[0] = 4; ti[1] = k; ti[2] = other; ti[3] = mh; ti[4] = dd; ti[5] = v;
                ti[6] = 0; ti[7] = 0;
            }
        }
        if (prune)
        {
            double* Pm = sm_P(S);
            Pm[k * K + other] = 0.0;
            Pm[other * K + k] = 0.0;
            int dk = 1, dother = 1;
            for (int i = 0; i < K; ++i)
            {
                if (!(Pm[k * K + i] == 0)) dk = 0;
            }
            Q->is_done[k] = dk;
            for (int i = 0; i < K; ++i)
            {
                if (!(Pm[other * K + i] == 0)) dother = 0;
            }
            Q->is_done[other] = dother;
            double sum_t = 0.0, sum_o = 0.0;
            for (int i = 0; i < K; ++i) sum_t += Pm[k * K + i];
            for (int i = 0; i < K; ++i) sum_o += Pm[other * K + i];
            for (int i = 0; i < K; ++i)
            {
                Pm[k * K + i] /= sum_t;
                Pm[other * K + i] /= sum_o;
            }
        }
    }

    // Runs the serial part of expandTree_ (:337-356) until it needs the CTA or the
    // launch is over; S.req tells the CTA what to do next.
    template <class Cta>
    IMOMD_HD void step(Cta& cta, int iters)
    {
        int K = Q->K;
        for (;;)
        {
            if (Q->status != 0)
            {
                S.req = REQ_EXIT;
                return;
            }
            int k = S.k;
            TreeRef T = tree_ref_of(P, q, k < K ? k : 0, S);
            switch (S.phase)
            {
                case PH_NEXT_TREE:
                {
                    if (k >= K)
                    {
                        S.it += 1;
                        S.k = 0;
                        Q->iterations += 1;
                        // done_time is evaluated after every pass (:258-262)
                        if (P.budget_ns && cta.now_ns() - S.t_start_ns > P.budget_ns) S.stop = 1;
                        continue;
                    }
                    if (S.it >= iters || S.stop)
                    {
                        S.req = REQ_EXIT;
                        return;
                    }
                    if (k == 0)
                    {
                        int active = 0;
                        for (int j = 0; j < K; ++j)
                        {
                            if (Q->frontier_count[j] != 0 && !Q->is_done[j]) ++active;
                        }
                        if (active == 0)
                        {
                            // nothing can expand any more: the remaining passes are no-ops
                            Q->iterations += iters - S.it;
                            S.it = iters;
                            S.req = REQ_EXIT;
                            return;
                        }
                    }
                    if (Q->frontier_count[k] == 0 || Q->is_done[k])
                    {
                        S.k = k + 1;
                        continue;
                    }
                    if (Q->word_cursor + IMOMD_WORD_MARGIN > P.words_avail)
                    {
                        Q->status = 1;   // IMOMD_RUN_NEED_WORDS (resumable)
                        continue;
                    }
                    sample(k, T);
                    if (Q->status != 0) continue;
                    S.req = REQ_NN;
                    S.phase = PH_AFTER_NN;
                    return;
                }
                case PH_AFTER_NN:
                {
                    uint32_t x = S.result.id1;
                    S.x_new = x;
                    S.conn_pending = 0;
                    S.phase = PH_JPS;
                    if (deg(x) == 2)
                    {
                        // updateExpandables(tree, x_new, false), :422-425
                        fr_erase(P, Q, T, k, x);
                        collect_rescans(k, x);
                        if (S.n_rescan > 0)
                        {
                            S.req = REQ_RESCAN;
                            return;
                        }
                    }
                    continue;
                }
                case PH_JPS:
                {
                    if (S.conn_pending)
                    {
                        // the hop's updateConnectionTree_ (:446 / :456), then move on
                        apply_connection(k, T, S.conn_x, S.conn_cx);
                        S.conn_pending = 0;
                        S.x_new = S.jps_next;
                    }
                    if (S.batch)
                    {
                        // hops need no service here: their connection checks are only logged
                        while (Q->status == 0 && jps_hop(k, T))
                        {
                            queue_check(S.conn_x);
                            S.x_new = S.jps_next;
                        }
                    }
                    else if (!S.attach_mode && jps_hop(k, T))
                    {
                        S.conn_pending = 1;
                        S.req = REQ_CONN;
                        return;
                    }
                    uint32_t x_new = S.x_new;
                    connect_new_node(k, T, x_new);
                    // updateExpandables(tree, x_new, true), first half
                    fr_erase(P, Q, T, k, x_new);
                    collect_rescans(k, x_new);
                    S.phase = PH_INSERT;
                    if (S.n_rescan > 0)
                    {
                        S.req = REQ_RESCAN;
                        return;
                    }
                    continue;
                }
                case PH_INSERT:
                {
                    insert_neighbours(k, T, S.x_new);
                    S.phase = PH_FINISH;
                    if (Q->status == 0 && S.n_ys > 0)
                    {
                        S.req = REQ_MHINS;
                        return;
                    }
                    continue;
                }
                case PH_FINISH:
                {
                    if (S.batch)
                    {
                        // rewireTree_(x_new) is the relaxation after the step; :354-355 deferred: the
                        // selection-probability test with the heuristic entry as it is NOW, the
                        // connection check of x_new with the costs the relaxation leaves
                        int other = Q->K;
                        for (int i = 0; i < Q->K; ++i)
                        {
                            if ((uint64_t)Q->dest[i] == S.rand_id)
                            {
                                other = i;
                                break;
                            }
                        }
                        if (other < Q->K)
                            log_event(k, 1u, IMOMD_NONE, (uint32_t)other, sm_MHv(S)[k * Q->K + other],
                                      sm_MHi(S)[k * Q->K + other]);
                        if (T.parent(S.x_new) != IMOMD_NONE)
                        {
                            queue_rewire(S.x_new);
                            queue_check(S.x_new);
                        }
                        S.exp_n += 1;
                        S.req = REQ_EXIT;
                        S.phase = PH_NEXT_TREE;
                        return;
                    }
                    S.rw_fp = IMOMD_NONE;
                    S.rw_top = 0;
                    S.bfs_valid = 0;
                    S.chk_valid = 0;
                    push_frame(S.x_new, row_begin(P.g, S.x_new));
                    S.phase = PH_REWIRE;
                    continue;
                }
                case PH_REWIRE:
                {
                    if (!rewire_run(k, T)) return;   // REQ_BFS or REQ_CHECK posted
                    if (Q->status != 0) continue;
                    if (!S.attach_mode) update_selection_probability(cta, k);
                    S.conn_x = S.x_new;
                    S.req = REQ_CONN;
                    S.phase = PH_LAST_CONN;
                    return;
                }
                case PH_LAST_CONN:
                {
                    apply_connection(k, T, S.x_new, T.cost(S.x_new));         // :355
                    if (S.attach_mode)
                    {
                        // report the parent this vertex received, then take the next one
                        P.attach_list[S.attach_i] = T.parent(S.x_new);
                        S.attach_i += 1;
                        if (S.attach_i < S.attach_n && Q->status == 0)
                        {
                            S.x_new = P.attach_list[S.attach_i];
                            S.conn_pending = 0;
                            S.phase = PH_JPS;
                            continue;
                        }
                        S.req = REQ_EXIT;
                        return;
                    }
                    Q->expansions += 1;
                    S.k = k + 1;
                    S.phase = PH_NEXT_TREE;
                    continue;
                }
            }
        }
    }
};

// ====================================================== CTA-wide services ====

// All threads: evaluate the records of the chunk work list with `eval(rec)`.
template <class Cta, class Eval>
IMOMD_HD inline Best scan_list(Cta& cta, const TreeRef& T, const uint32_t* clist, uint32_t first,
                               uint32_t n_list, Eval eval)
{
    Best b = best_empty();
    uint32_t total = (n_list - first) * IMOMD_CHUNK;
    for (uint32_t i = cta.tid(); i < total; i += cta.nt())
    {
        uint32_t li = first + i / IMOMD_CHUNK;
        uint32_t lane = i % IMOMD_CHUNK;
        uint32_t chunk = clist[2 * li];
        uint32_t cnt = clist[2 * li + 1];
        if (lane < cnt)
        {
            const VRec& r = T.rec[(size_t)chunk * IMOMD_CHUNK + lane];
            best_add(b, eval(r), r.id);
        }
    }
    return b;
}

// one thread: append the chunk chain of cell c to the work list
template <class Cta>
IMOMD_HD inline void list_cell(Cta& cta, const TreeRef& T, uint32_t* clist, uint32_t* n_list,
                               uint32_t c, uint32_t cnt)
{
    uint32_t chunk = T.cell_head[c];
    uint32_t in_head = (cnt - 1) % IMOMD_CHUNK + 1;
    uint32_t n = in_head;
    while (chunk != IMOMD_NONE && cnt > 0)
    {
        uint32_t at = cta.counter_add(n_list, 1u);
        clist[2 * at] = chunk;
        clist[2 * at + 1] = n;
        cta.counter_add(n_list + 2, n);   // Shared::n_recs sits two words after n_list
        cnt -= n;
        n = IMOMD_CHUNK;
        chunk = T.chunk_next[chunk];
    }
}

// Pruned arg-min over the frontier hash of one tree.
//   bounds(c, lo, hi)    conservative bounds of the objective over cell c
//   eval(r)              objective of a frontier record
//   bound(m)             largest lower bound a cell may have and still matter when the
//                        minimum is known to be at most m
// The smallest upper bound over the occupied cells caps the minimum, every occupied cell
// whose lower bound does not exceed that cap is listed, its 32-record chunks are ranked one
// record per thread.  Returns the same Best on every thread.  `lb` holds G*G doubles.
template <class Cta, class Bounds, class Eval, class Bound>
IMOMD_HD inline Best pruned_argmin(Cta& cta, const PlannerDev& P, const TreeRef& T, Shared& S,
                                   uint32_t n_occ, double* lb, uint32_t* clist, Bounds bounds,
                                   Eval eval, Bound bound)
{
    if (n_occ == 0) return best_empty();
    double mub = INFINITY;
    for (uint32_t i = cta.tid(); i < n_occ; i += cta.nt())
    {
        uint32_t c = T.occ[i];
        double l, u;
        bounds(c, l, u);
        if (lb) lb[i] = l;      // (no table in the batched mode: a warp recomputes the bound below)
        if (u < mub) mub = u;
    }
    if (cta.tid() == 0)
    {
        S.n_list = 0;
        S.n_cells += n_occ;
    }
    mub = cta.reduce_min_f64(mub, S.red);   // (barriers inside order the writes above)
    const double limit = bound(mub);
    for (uint32_t i = cta.tid(); i < n_occ; i += cta.nt())
    {
        uint32_t c = T.occ[i];
        double l;
        if (lb) l = lb[i];
        else
        {
            double u;
            bounds(c, l, u);
        }
        if (l <= limit) list_cell(cta, T, clist, &S.n_list, c, T.cell_cnt[c]);
    }
    cta.sync();
    Best b = scan_list(cta, T, clist, 0u, S.n_list, eval);
    return cta.reduce(b, S.red);
}

// Helper-thread prefetch, run by the second warp WHILE the serial thread processes the result
// of a nearest search (the other warps would only wait at the barrier).  The serial bookkeeping
// that follows a nearest search touches
// lines nobody has looked at yet (records of unvisited neighbours, their hash cells, the
// chunk that gives up its last record): one dependent DRAM round trip each when thread 0
// meets them alone.  Here the idle threads of the CTA fetch those lines in parallel --
// one thread per neighbour -- so the serial code finds them in L1/L2.  Purely a latency
// optimisation: nothing is written, results are discarded.
template <class Cta>
IMOMD_HD inline void warm_expansion(Cta& cta, const PlannerDev& P, const TreeRef& T, int q, int k,
                                    uint32_t x, Shared& S)
{
    const GraphDev& g = P.g;
    if (x == IMOMD_NONE || cta.nt() < 64 || cta.tid() < 32) return;   // helpers: the second warp on
    const uint32_t e0 = row_begin(g, x), e1 = row_end(g, x);
    uint32_t acc = 0;
    const uint32_t t = (uint32_t)cta.tid() - 32u;
    if (t < e1 - e0)
    {
        uint32_t y = g.col[e0 + t];
        double w = g.w[e0 + t];
        THead h = T.head(y);
        acc += T.chl(y)[0] + h.parent + (w > 0.0 ? 1u : 0u);
        VRec v = g.vrec[y];
        acc += T.cell_cnt[v.cell] + T.cell_head[v.cell];
        acc += g.deg[y] + g.col[row_begin(g, y)];
    }
    else if (t == 32)
    {
        THead h = T.head(x);
        uint32_t c = g.vrec[x].cell;
        uint32_t cnt = T.cell_cnt[c], hd = T.cell_head[c];
        if (cnt != 0 && hd != IMOMD_NONE) acc += T.rec[(size_t)hd * IMOMD_CHUNK + (cnt - 1) % IMOMD_CHUNK].id;
        acc += h.fslot + T.chl(x)[0];
    }
    else if (t == 33)
    {
        uint32_t ft = S.Q->free_top[k];
        if (ft >= 2) acc += T.free_stack[ft - 1] + T.free_stack[ft - 2];
    }
    if (acc == 0x9E3779B9u) S.scan[39] = acc;   // keeps the loads alive, practically never taken
}

// The same idea for imomd_attach_list (pseudo-tree merge): the vertices to attach are known in
// advance, so while thread 0 works on entry i the second warp of the CTA (the other lanes of the
// warp when the kernel is launched with 32 threads) fetches what entries i+1 .. i+8 will touch: the vertex's adjacency row, the tree records and child rows of its
// neighbours, their hash cells.  Nothing is written; stale reads are harmless.
template <class Cta>
IMOMD_HD inline void warm_attach(Cta& cta, const PlannerDev& P, const TreeRef& T, uint32_t at, uint32_t n, Shared& S)
{
    const GraphDev& g = P.g;
    // helpers: the second warp when there is one (no divergence with the serial lane), else the
    // other lanes of the only warp
    uint32_t t = (uint32_t)cta.tid();
    if (cta.nt() >= 64)
    {
        if (t < 32 || t >= 64) return;
        t -= 31;
    }
    if (t == 0 || t > 32) return;
    const uint32_t ahead = 1 + (t - 1) / 4, slot = (t - 1) % 4;
    if (at + ahead >= n) return;
    const uint32_t x = P.attach_list[at + ahead];
    if (x >= g.n) return;
    uint32_t acc = 0;
    const uint32_t e0 = row_begin(g, x), d = g.deg[x];
    for (uint32_t j = slot; j < d; j += 4)
    {
        const uint32_t y = g.col[e0 + j];
        const double w = g.w[e0 + j];
        THead h = T.head(y);
        acc += T.chl(y)[0] + h.parent + (w > 0.0 ? 1u : 0u);
        VRec v = g.vrec[y];
        acc += T.cell_cnt[v.cell] + T.cell_head[v.cell] + g.deg[y];
    }
    if (slot == 0)
    {
        THead h = T.head(x);
        const uint32_t c = g.vrec[x].cell;
        acc += h.fslot + T.chl(x)[0] + T.cell_cnt[c] + T.cell_head[c];
        if (h.fslot != IMOMD_NONE) acc += T.rec[h.fslot].id;
    }
    if (acc == 0x9E3779B9u) S.scan[39] = acc;   // keeps the loads alive
}

// steerNewNode_ arg-min (src/imomd_rrt_star.cpp:406-418) for the sample in S.
template <class Cta>
IMOMD_HD IMOMD_NOINLINE void svc_nearest(Cta& cta, const PlannerDev& P, int q, int k, Shared& S,
                                 double* lb, uint32_t* clist, int64_t* near_ties,
                                 int64_t* unresolved_ties, bool warm)
{
    assume_planner(P);
    IMOMD_ASSUME_SHARED(&S);
    if (lb) IMOMD_ASSUME_SHARED(lb);
    const GraphDev& g = P.g;
    TreeRef T = tree_ref_of(P, q, k, S);
    double qlat = S.qlat, qlon = S.qlon;
    double cl;
    double qx, qy, qz;
    if (warm && S.sample_is_vertex && S.rand_id < (uint64_t)g.n)
    {
        // the planner's samples are graph vertices (selectRandomVertex_): their unit vector was
        // computed at upload with the very expressions below (make_vrec), so one 32-byte load
        // replaces three double-precision sin / cos evaluations in every thread
        const VRec me = g.vrec[(uint32_t)S.rand_id];
        qx = me.x;
        qy = me.y;
        qz = me.z;
        cl = g.coslat[(uint32_t)S.rand_id];
    }
    else
    {
        cl = cos(qlat * kDeg2Rad);
        double sl = sin(qlat * kDeg2Rad);
        double co = cos(qlon * kDeg2Rad), so = sin(qlon * kDeg2Rad);
        qx = cl * co;
        qy = cl * so;
        qz = sl;
    }
    auto bounds = [&](uint32_t c, double& lo, double& hi) {
        cell_bounds_a(g, c, qlat, qlon, cl, &lo, &hi);
        lo *= 4.0;
        hi *= 4.0;
    };
    auto eval = [&](const VRec& r) {
        double dx = r.x - qx, dy = r.y - qy, dz = r.z - qz;
        return dx * dx + dy * dy + dz * dz;
    };
    // rounding band of a squared chord m: unit-vector components carry ~1e-16 each
    // (an upper bound of 2e-12 m + 4e-15 sqrt(m): single-precision root, rounded up)
    // (a chord^2 differs relatively twice as much as the distance: a widened P.tie_rel -- tests --
    // widens the band with it)
    const double band_rel = 2e-12 + (P.tie_rel > kTieRel ? 4.0 * P.tie_rel : 0.0);
    auto band = [band_rel](double m) { return band_rel * m + 4.00001e-15 * (double)sqrtf((float)m) + 1e-33; };
    auto bound = [&](double m) { return m * (1.0 + kPruneSlack) + 4.0 * band(m); };
    if (cta.tid() == 0)
    {
        S.n_cells = 0;
        S.n_recs = 0;
    }
    cta.sync();
    Best b = pruned_argmin(cta, P, T, S, S.Q->n_occ[k], lb, clist, bounds, eval, bound);
    int ambiguous = (b.id1 != IMOMD_NONE) && (b.v2 - b.v1 <= band(b.v1));
    if (ambiguous)
    {
        // slow exact pass: full haversine for everything inside the band
        double cut = b.v1 + 2.0 * band(b.v1);
        auto eval2 = [&](const VRec& r) {
            double dx = r.x - qx, dy = r.y - qy, dz = r.z - qz;
            double c2 = dx * dx + dy * dy + dz * dz;
            if (c2 > cut) return (double)INFINITY;
            return haversine(g.lat[r.id], g.lon[r.id], qlat, qlon);
        };
        Best e = scan_list(cta, T, clist, 0u, S.n_list, eval2);
        e = cta.reduce(e, S.red);
        b = e;
        if (cta.tid() == 0) *near_ties += 1;
        if (e.v2 - e.v1 <= P.tie_rel * e.v1)
        {
            // CUDA's libm and glibc may order these differently: collect every vertex inside the
            // band and let the host decide with the reference's own arithmetic
            if (cta.tid() == 0) S.tie_n = 0;
            cta.sync();
            const double lim = e.v1 + 2.0 * P.tie_rel * e.v1;
            auto collect = [&](const VRec& r) {
                double v = eval2(r);
                if (v <= lim)
                {
                    uint32_t at = cta.counter_add(&S.tie_n, 1u);
                    if (at < IMOMD_TIE_IDS) S.tie_id[at] = r.id;
                }
                return v;
            };
            (void)scan_list(cta, T, clist, 0u, S.n_list, collect);
            cta.sync();
            if (cta.tid() == 0)
            {
                double vals[IMOMD_TIE_IDS];
                bool exact_tie = false;
                int w = certify_argmin(cta, g, S, S.tie_n, false, qlat, qlon, 0.0, 0.0, vals, &exact_tie);
                if (w >= 0)
                {
                    b.id1 = S.tie_id[w];
                    b.v1 = vals[w];
                    S.certified_ties += 1;
                }
                if (w < 0 || exact_tie)
                {
                    *unresolved_ties += 1;
                    double* ti = S.Q->tie_info;
                    ti[0] = 1; ti[1] = k; ti[2] = -1; ti[3] = e.v1; ti[4] = e.v2; ti[5] = b.id1;
                    ti[6] = qlat; ti[7] = qlon;
                }
            }
        }
    }
    if (cta.tid() == 0)
    {
        S.result = b;
        S.ambiguous = ambiguous;
        QueryDev* Qs = S.Q;
        S.stat[0] += S.n_cells;
        S.stat[1] += S.n_recs;
    }
}

// frontier rescans of updateExpandables (src/imomd_rrt_star.cpp:510-524) for the
// pairs (k, o) listed in S.rescan_o
template <class Cta>
IMOMD_HD IMOMD_NOINLINE void svc_rescan(Cta& cta, const PlannerDev& P, int q, int k, Shared& S, double* lb)
{
    assume_planner(P);
    IMOMD_ASSUME_SHARED(&S);
    if (lb) IMOMD_ASSUME_SHARED(lb);
    const GraphDev& g = P.g;
    QueryDev* Q = S.Q;
    TreeRef T = tree_ref_of(P, q, k, S);
    int K = Q->K;
    uint32_t cells = (uint32_t)g.G * (uint32_t)g.G;
    uint32_t* clist = S.clist;
    IMOMD_ASSUME_GLOBAL(clist);
    const double* lbd = P.lbd + (size_t)q * K * cells;
    const double* ubd = P.ubd + (size_t)q * K * cells;
    uint32_t rk = Q->view_root[k];
    double klat = g.lat[rk], klon = g.lon[rk];
    int n = S.n_rescan;
    for (int i = 0; i < n; ++i)
    {
        int o = S.rescan_o[i];
        uint32_t ro = Q->view_root[o];
        double olat = g.lat[ro], olon = g.lon[ro];
        const double* lk = lbd + (size_t)k * cells;
        const double* lo = lbd + (size_t)o * cells;
        const double* uk = ubd + (size_t)k * cells;
        const double* uo = ubd + (size_t)o * cells;
        auto bounds = [&](uint32_t c, double& lower, double& upper) {
            lower = lk[c] + lo[c];
            upper = uk[c] + uo[c];
        };
        const double kcos = g.coslat[rk], ocos = g.coslat[ro];
        auto eval = [&](const VRec& r) {
            double vlat = g.lat[r.id], vlon = g.lon[r.id], vcos = g.coslat[r.id];
            return haversine_pc(vlat, vlon, vcos, klat, klon, kcos) + haversine_pc(vlat, vlon, vcos, olat, olon, ocos);
        };
        auto bound = [&](double m) { return m * (1.0 + kPruneSlack); };
        if (cta.tid() == 0)
        {
            S.n_cells = 0;
            S.n_recs = 0;
        }
        cta.sync();
        Best b = pruned_argmin(cta, P, T, S, Q->n_occ[k], lb, clist, bounds, eval, bound);
        const bool undecided = b.id1 != IMOMD_NONE && b.v2 - b.v1 <= P.tie_rel * b.v1;
        if (undecided)
        {
            if (cta.tid() == 0) S.tie_n = 0;
            cta.sync();
            const double lim = b.v1 + 2.0 * P.tie_rel * b.v1;
            auto collect = [&](const VRec& r) {
                double v = eval(r);
                if (v <= lim)
                {
                    uint32_t at = cta.counter_add(&S.tie_n, 1u);
                    if (at < IMOMD_TIE_IDS) S.tie_id[at] = r.id;
                }
                return v;
            };
            (void)scan_list(cta, T, clist, 0u, S.n_list, collect);
            cta.sync();
        }
        if (cta.tid() == 0)
        {
            S.stat[0] += S.n_cells;
            S.stat[7] += S.n_recs;
            if (undecided)
            {
                double vals[IMOMD_TIE_IDS];
                bool exact_tie = false;
                int w = certify_argmin(cta, g, S, S.tie_n, true, klat, klon, olat, olon, vals, &exact_tie);
                if (w >= 0)
                {
                    // the holder is glibc's; its value stays the device's (1e-12 relative)
                    if (S.tie_id[w] != b.id1)
                    {
                        const uint32_t v = S.tie_id[w];
                        b.id1 = v;
                        b.v1 = haversine_pc(g.lat[v], g.lon[v], g.coslat[v], klat, klon, kcos) +
                               haversine_pc(g.lat[v], g.lon[v], g.coslat[v], olat, olon, ocos);
                    }
                    S.certified_ties += 1;
                }
                if (w < 0 || exact_tie)
                {
                    S.unresolved_ties += 1;
                    double* ti = Q->tie_info;
                    ti[0] = 2; ti[1] = k; ti[2] = o; ti[3] = b.v1; ti[4] = b.v2; ti[5] = b.id1;
                    ti[6] = 0; ti[7] = 0;
                }
            }
            sm_MHi(S)[k * K + o] = b.id1;
            sm_MHv(S)[k * K + o] = b.v1;
        }
        cta.sync();
    }
}

// heuristic updates for the vertices just inserted into the frontier
// (src/imomd_rrt_star.cpp:541-556): h(y, o) = hav(y, root_k) + hav(y, root_o)
template <class Cta>
IMOMD_HD IMOMD_NOINLINE void svc_mhins(Cta& cta, const PlannerDev& P, int q, int k, Shared& S)
{
    assume_planner(P);
    IMOMD_ASSUME_SHARED(&S);
    const GraphDev& g = P.g;
    QueryDev* Q = S.Q;
    int K = Q->K;
    int m = S.n_ys;
    double* hv = S.h;       // [m][K], m <= the graph's adjacency width
    IMOMD_ASSUME_SHARED(hv);
    for (int i = cta.tid(); i < m * K; i += cta.nt())
    {
        int yi = i / K, o = i % K;
        uint32_t y = S.ys[yi];
        uint32_t r = Q->view_root[o];
        hv[i] = haversine_pc(g.lat[y], g.lon[y], g.coslat[y], g.lat[r], g.lon[r], g.coslat[r]);
    }
    if (cta.tid() == 0) S.mh_tie_mask = 0u;
    cta.sync();
    for (int o = cta.tid(); o < K; o += cta.nt())
    {
        if (o == k) continue;
        double best = sm_MHv(S)[k * K + o];
        uint32_t arg = sm_MHi(S)[k * K + o];
        bool close = false;
        for (int yi = 0; yi < m; ++yi)
        {
            double h = hv[yi * K + k] + hv[yi * K + o];
            // (an inserted vertex that already IS the holder compares equal in any arithmetic)
            if (arg != IMOMD_NONE && S.ys[yi] != arg && fabs(h - best) <= P.tie_rel * best) close = true;
            if (h < best)
            {
                best = h;
                arg = S.ys[yi];
            }
        }
        if (close)
        {
            cta.counter_or(&S.mh_tie_mask, 1u << o);   // redone below with the exact values
            continue;
        }
        sm_MHv(S)[k * K + o] = best;
        sm_MHi(S)[k * K + o] = arg;
    }
    cta.sync();
    if (S.mh_tie_mask == 0u) return;
    if (cta.tid() == 0)
    {
        // a candidate within 1e-12 of the holder (or of another candidate): the sequence of strict
        // `<` comparisons of :541-556 is replayed on glibc's values (holder first, then the
        // inserted vertices in adjacency order)
        const uint32_t rk = Q->view_root[k];
        for (int o = 0; o < K; ++o)
        {
            if (!((S.mh_tie_mask >> o) & 1u)) continue;
            const uint32_t ro = Q->view_root[o];
            const uint32_t holder = sm_MHi(S)[k * K + o];
            uint32_t ids[IMOMD_MAXDEG + 1];
            double dev[IMOMD_MAXDEG + 1];
            int n = 0;
            if (holder != IMOMD_NONE)
            {
                ids[n] = holder;
                dev[n] = sm_MHv(S)[k * K + o];
                ++n;
            }
            for (int yi = 0; yi < m; ++yi)
            {
                ids[n] = S.ys[yi];
                dev[n] = hv[yi * K + k] + hv[yi * K + o];
                ++n;
            }
            // (rounds of IMOMD_TIE_PAIRS / 2 candidates per mailbox request)
            double exact[IMOMD_MAXDEG + 1];
            bool ok = true;
            for (int i0 = 0; ok && i0 < n; i0 += IMOMD_TIE_PAIRS / 2)
            {
                const int mm = n - i0 < IMOMD_TIE_PAIRS / 2 ? n - i0 : IMOMD_TIE_PAIRS / 2;
                for (int i = 0; i < mm; ++i)
                {
                    const uint32_t v = ids[i0 + i];
                    double* a = S.tie_in[2 * i];
                    double* b = S.tie_in[2 * i + 1];
                    a[0] = g.lat[v]; a[1] = g.lon[v]; a[2] = g.lat[rk]; a[3] = g.lon[rk];
                    b[0] = g.lat[v]; b[1] = g.lon[v]; b[2] = g.lat[ro]; b[3] = g.lon[ro];
                }
                ok = cta.certify(2 * mm, S.tie_in, S.tie_out);
                for (int i = 0; ok && i < mm; ++i) exact[i0 + i] = S.tie_out[2 * i] + S.tie_out[2 * i + 1];
            }
            int w = 0;
            if (ok)
            {
                double best = exact[0];
                for (int i = 1; i < n; ++i)
                {
                    double h = exact[i];
                    if (h < best)
                    {
                        best = h;
                        w = i;
                    }
                }
                S.certified_ties += 1;
            }
            else
            {
                for (int i = 1; i < n; ++i)
                    if (dev[i] < dev[w]) w = i;
                S.unresolved_ties += 1;
                double* ti = Q->tie_info;
                ti[0] = 3; ti[1] = k; ti[2] = o; ti[3] = dev[0]; ti[4] = dev[w]; ti[5] = ids[w];
                ti[6] = 0; ti[7] = 0;
            }
            sm_MHv(S)[k * K + o] = dev[w];
            sm_MHi(S)[k * K + o] = ids[w];
        }
    }
}

IMOMD_HD inline void conn_gather_one(const PlannerDev& P, int q, uint32_t x, int o, Shared& S)
{
    const size_t qt = (size_t)q * P.K + o;
    const U4 t = *reinterpret_cast<const U4*>(P.trec + (qt * (size_t)P.g.n + x) * (size_t)P.trec_stride);
    unsigned long long bits = ((unsigned long long)t.w << 32) | (unsigned long long)t.z;
    double c;
    memcpy(&c, &bits, 8);
    S.conn_par[o] = t.x;
    S.conn_cost[o] = c;
}

// the other trees' parent / cost at vertex x, one tree per thread
template <class Cta>
IMOMD_HD IMOMD_NOINLINE void svc_conn_gather(Cta& cta, const PlannerDev& P, int q, uint32_t x, Shared& S)
{
    assume_planner(P);
    IMOMD_ASSUME_SHARED(&S);
    int K = P.K;
    for (int o = cta.tid(); o < K; o += cta.nt()) conn_gather_one(P, q, x, o, S);
}

// tree_t::updateCost (include/imomd_rrt_star/imomd_rrt_star.h:102-123) as a level-wise
// parallel BFS.  The FIFO queue of the sequential loop IS the output list
// (`updated_nodes`): the vertices of one level are expanded one per thread, an exclusive
// scan of the child counts places their children behind everything queued before, so the
// list comes out in exactly the sequential order.  Every vertex sits in at most one child
// list, so the cost updates never collide.
//
// Latency is what matters here (deep, narrow subtrees): the current level lives in shared
// memory, a child row is one 128-bit load, the read-modify-write of the children's costs
// is deferred by one round so that only the child-row load is on the critical path, and
// levels of at most 32 vertices are walked by warp 0 alone with shuffles (no block barrier).
constexpr uint32_t kLevelCap = 256;

template <int W>
struct BfsPending
{
    int n;   // (kept for the interface; cost updates ride on the record load now)
};

template <int W>
IMOMD_HD inline void bfs_flush(const TreeRef&, BfsPending<W>&, double, double)
{
}

// one vertex of the current level: a single record load returns its cost (updated in
// place, include/imomd_rrt_star/imomd_rrt_star.h:116) and its child row
template <int W>
IMOMD_HD inline int bfs_expand(const TreeRef& T, uint32_t node, bool active, BfsPending<W>&,
                               double new_cost, double old, ChlRow<W>& row, uint32_t rewired)
{
    int cnt = 0;
    if (active)
    {
        THead h = T.head(node);
        row = chl_load<W>(T.chl(node));
        if (node != rewired) T.set_cost(node, new_cost + h.cost - old);
#pragma unroll
        for (int i = 0; i < W; ++i)
        {
            if (row.v[i] != IMOMD_NONE) cnt = i + 1;
        }
    }
    return cnt;
}

template <int W, class Cta>
IMOMD_HD IMOMD_NOINLINE void svc_bfs_w(Cta& cta, const PlannerDev& P, int q, int k, Shared& S,
                               uint32_t* lvl)
{
    assume_planner(P);
    IMOMD_ASSUME_SHARED(&S);
    IMOMD_ASSUME_SHARED(lvl);
    TreeRef T = tree_ref_of(P, q, k, S);
    uint32_t* arena = S.arena;
    IMOMD_ASSUME_GLOBAL(arena);
    const uint32_t cap = S.arena_cap;
    const uint32_t rewired = S.bfs_node;
    const double new_cost = S.bfs_cost;
    const uint32_t lo = S.bfs_lo;
    const double old = T.cost(rewired);
    BfsPending<W> pd;
    pd.n = 0;
    // The connection check that follows the propagation needs the OTHER trees' records of
    // `rewired` (untouched by this walk): the second warp fetches them while the first one walks.
    const bool early_gather = cta.nt() >= 64 && !S.batch;
    if (early_gather)
    {
        const int o = cta.tid() - 32;
        if (o >= 0)
        {
            for (int oo = o; oo < P.K; oo += cta.nt() - 32)
            {
                if (oo != k) conn_gather_one(P, q, rewired, oo, S);
            }
        }
    }
    if (cta.tid() == 0)
    {
        S.lv_overflow = lo >= cap ? 1 : 0;
        if (!S.lv_overflow)
        {
            arena[lo] = rewired;
            lvl[0] = rewired;
        }
        S.lv_cur = 0;
        S.lv_start = lo;
        S.lv_n = 1;
        S.lv_tail = lo + 1;
    }
    cta.sync();
    for (;;)
    {
        uint32_t n = S.lv_n;
        if (n == 0 || S.lv_overflow) break;
        if (n <= (uint32_t)cta.warp_width())
        {
            // ---- narrow levels: warp 0 only, shuffles, no block barrier ----
            if (cta.warp() == 0)
            {
                uint32_t cur = S.lv_cur, start = S.lv_start, tail = S.lv_tail;
                uint32_t ovf = 0;
                uint32_t levels = 0;
                while (n != 0 && n <= (uint32_t)cta.warp_width() && !ovf)
                {
                    uint32_t* cl = lvl + cur * kLevelCap;
                    uint32_t* nl = lvl + (cur ^ 1u) * kLevelCap;
                    const bool active = (uint32_t)cta.lane() < n;
                    uint32_t node = active ? cl[cta.lane()] : 0u;
                    // Lockstep run: a rewired subtree is mostly a handful of parallel chains
                    // (1.4 vertices per level on a grid).  While every vertex of the level has
                    // exactly one child the next level has the same shape -- lane i's vertex is
                    // replaced by its child -- so the walk stays in registers: one record load,
                    // one vote, one store per level, no scan and no shared-memory round trip.
                    for (;;)
                    {
                        ChlRow<W> row;
                        int cnt = bfs_expand<W>(T, node, active, pd, new_cost, old, row, rewired);
                        if (cta.warp_all(!active || cnt == 1))
                        {
                            if (tail + n > cap)
                            {
                                ovf = 1;
                                break;
                            }
                            if (active) arena[tail + (uint32_t)cta.lane()] = row.v[0];
                            start = tail;
                            tail += n;
                            ++levels;
                            node = row.v[0];
                            continue;
                        }
                        // mixed level: place the children behind everything queued before
                        uint32_t total = 0;
                        uint32_t off = cta.warp_scan_small((uint32_t)cnt, W, &total);
                        if (tail + total > cap)
                        {
                            ovf = 1;
                            break;
                        }
#pragma unroll
                        for (int i = 0; i < W; ++i)
                        {
                            if (i < cnt)
                            {
                                uint32_t child = row.v[i];
                                arena[tail + off + i] = child;
                                if (off + i < kLevelCap) nl[off + i] = child;
                            }
                        }
                        pd.n = cnt;
                        start = tail;
                        tail += total;
                        n = total;
                        cur ^= 1u;
                        ++levels;
                        cta.warp_sync();
                        break;
                    }
                }
                if (cta.lane() == 0)
                {
                    S.Q->prof_count[0] += levels;
                    S.lv_cur = cur;
                    S.lv_start = start;
                    S.lv_tail = tail;
                    S.lv_n = n;
                    if (ovf) S.lv_overflow = 1;
                }
            }
            cta.sync();
            continue;
        }
        // ---- wide level: the whole CTA, chunks of nt vertices ----
        {
            uint32_t cur = S.lv_cur, start = S.lv_start, tail = S.lv_tail;
            uint32_t* cl = lvl + cur * kLevelCap;
            uint32_t* nl = lvl + (cur ^ 1u) * kLevelCap;
            uint32_t produced = 0;
            bool ovf = false;
            for (uint32_t c0 = 0; c0 < n; c0 += (uint32_t)cta.nt())
            {
                uint32_t idx = c0 + (uint32_t)cta.tid();
                bool active = idx < n;
                uint32_t node = 0u;
                if (active) node = idx < kLevelCap ? cl[idx] : arena[start + idx];
                ChlRow<W> row;
                int cnt = bfs_expand<W>(T, node, active, pd, new_cost, old, row, rewired);
                uint32_t total = 0;
                uint32_t off = cta.scan_exclusive((uint32_t)cnt, &total, S.scan);
                if (tail + produced + total > cap)
                {
                    ovf = true;
                    break;
                }
#pragma unroll
                for (int i = 0; i < W; ++i)
                {
                    if (i < cnt)
                    {
                        uint32_t child = row.v[i];
                        uint32_t at = produced + off + i;
                        arena[tail + at] = child;
                        if (at < kLevelCap) nl[at] = child;
                    }
                }
                pd.n = cnt;
                produced += total;
            }
            cta.sync();
            if (cta.tid() == 0)
            {
                S.Q->prof_count[0] += 1;
                S.lv_cur = cur ^ 1u;
                S.lv_start = tail;
                S.lv_tail = tail + produced;
                S.lv_n = produced;
                if (ovf) S.lv_overflow = 1;
            }
            cta.sync();
        }
    }
    bfs_flush<W>(T, pd, new_cost, old);
    if (!early_gather && !S.batch) svc_conn_gather(cta, P, q, rewired, S);
    if (cta.tid() == 0)
    {
        // statistics: [0] = subtree vertices (cycles slot) and levels... see imomd_get_profile
        S.Q->prof_cycles[0] += (uint64_t)(S.lv_tail - lo);
        S.stat[5] += (uint64_t)(S.lv_tail - lo - 1);
        bool overflow = S.lv_overflow != 0;
        if (!overflow) T.set_cost(rewired, new_cost);
        S.bfs_hi = overflow ? IMOMD_NONE : S.lv_tail;
        S.bfs_valid = 1;
        S.chk_hi = S.lv_tail;     // the examination request that follows (run_query)
        S.chk_lo = lo;
        S.chk_skip = rewired;
    }
}

// Speculative examination of the next calls `rewireTree_(u)` of an updated_nodes list
// (src/imomd_rrt_star.cpp:589-596): a call changes nothing unless some visited neighbour
// y != parent[u] has cost[y] > cost[u] + w (:569-572).  All (call, adjacency slot) pairs of
// a batch are tested at once against the current tree; the first pair that would re-parent
// (in call order, then slot order) is reported, everything before it is skipped exactly.
template <int KW, class Cta>
IMOMD_HD IMOMD_NOINLINE void svc_check(Cta& cta, const PlannerDev& P, int q, int k, Shared& S)
{
    assume_planner(P);
    IMOMD_ASSUME_SHARED(&S);
    const GraphDev& g = P.g;
    TreeRef T = tree_ref_of(P, q, k, S);
    const uint32_t* arena = S.arena;
    IMOMD_ASSUME_GLOBAL(arena);
    const uint32_t W = (uint32_t)KW;
    uint32_t n = S.chk_hi - S.chk_lo;
    if (n > kCheckBatch) n = kCheckBatch;
    const uint32_t hi = S.chk_hi, skip = S.chk_skip;
    uint32_t key = IMOMD_NONE;
    uint32_t my_u = 0, my_e = 0, my_e1 = 0, my_y = 0, my_py = 0;
    double my_nc = 0.0;
    for (uint32_t idx = cta.tid(); idx < n * W; idx += cta.nt())
    {
        uint32_t i = idx / W, j = idx % W;
        uint32_t u = arena[hi - 1 - i];
        if (u == skip) continue;
        // padded rows: the slot's address follows from u alone, empty slots hold NONE
        uint32_t e = row_begin(g, u) + j;
        uint32_t y = g.col[e];
        double w = g.w[e];
        if (y == IMOMD_NONE) continue;
        THead hu = T.head(u);
        THead hy = T.head(y);
        if (hy.parent == IMOMD_NONE || y == hu.parent) continue;
        double nc = hu.cost + w;
        if (hy.cost > nc && idx < key)
        {
            key = idx;
            my_u = u;
            my_e = e;
            my_e1 = row_end(g, u);
            my_y = y;
            my_py = hy.parent;
            my_nc = nc;
        }
    }
    const uint32_t mine = key;
    key = cta.reduce_min_u32(key, S.scan);
    if (key != IMOMD_NONE && mine == key)
    {
        S.chk_u = my_u;
        S.chk_e = my_e;
        S.chk_e1 = my_e1;
        S.chk_y = my_y;
        S.chk_py = my_py;
        S.chk_nc = my_nc;
    }
    if (cta.tid() == 0)
    {
        S.chk_noop = key == IMOMD_NONE ? n : key / W;
        S.chk_slot = key == IMOMD_NONE ? IMOMD_NONE : key % W;
        S.chk_valid = 1;
    }
}

// Tree construction of the planner constructor (src/imomd_rrt_star.cpp:73-78):
// tree i gets its root and updateExpandables(root, true) while trees j > i still
// read as id 0 / root 0.  Serial; runs once per query.
IMOMD_HD inline void construct_trees(const PlannerDev& P, int q, int hq)
{
    // q: the slot that holds the trees; hq: the query (header) they belong to
    const GraphDev& g = P.g;
    QueryDev* Q = P.query + hq;
    int K = Q->K;
    for (int i = 0; i < K; ++i)
    {
        Q->view_id[i] = 0;
        Q->view_root[i] = 0;
    }
    for (int i = 0; i < K; ++i)
    {
        TreeRef T = tree_ref(P, q, i);
        uint32_t root = Q->dest[i];
        Q->view_id[i] = i;
        Q->view_root[i] = root;
        T.set_parent(root, root);
        T.set_cost(root, 0.0);
        dirty_mark(P, T, root);
        Q->tree_size[i] = 1;
        if (P.pseudo) P.visit_log[((size_t)q * K + i) * g.n] = root;
        Q->ds_parent[i] = i;
        // erase(root) is a no-op and no holder equals the root yet; neighbours:
        uint32_t e0 = row_begin(g, root), e1 = row_end(g, root);
        for (uint32_t e = e0; e < e1; ++e)
        {
            uint32_t y = g.col[e];
            if (T.parent(y) != IMOMD_NONE) continue;
            if (!fr_insert(P, Q, T, i, y)) return;
            for (int o = 0; o < K; ++o)
            {
                if (Q->view_id[o] == i) continue;
                int oid = Q->view_id[o];
                uint32_t oroot = Q->view_root[o];
                double h = haversine(g.lat[y], g.lon[y], g.lat[root], g.lon[root]) +
                           haversine(g.lat[y], g.lon[y], g.lat[oroot], g.lon[oroot]);
                if (h < Q->MHv[i * K + oid])
                {
                    Q->MHv[i * K + oid] = h;
                    Q->MHi[i * K + oid] = y;
                }
            }
        }
    }
}

// ------------------------------------------------ slot recycling (streaming planners) ---
// A streaming planner has fewer sets of per-query device state ("slots") than queries: a CTA
// takes the next query from a work queue, rebuilds its slot for it, runs the query to completion
// and keeps only the header (K x K matrices, counters, tree hash).

IMOMD_HD inline int lowest_bit(uint32_t w)
{
#ifdef __CUDA_ARCH__
    return __ffs((int)w) - 1;
#else
    return __builtin_ctz(w);
#endif
}
IMOMD_HD inline uint32_t count_bits(uint32_t w)
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__popc(w);
#else
    return (uint32_t)__builtin_popcount(w);
#endif
}

// All threads: the slot back to "nothing visited, empty frontier hashes, all chunks free".
// The records start out all-ones (planner creation) and every record a query writes is marked in
// its tree's dirty bitmap (dirty_mark), so only the marked records are rewritten: a query touches
// a few percent of the K n records of a slot.
template <class Cta>
IMOMD_HD inline void reset_slot(Cta& cta, const PlannerDev& P, int slot)
{
    const size_t K = (size_t)P.K;
    const size_t cells = (size_t)P.g.G * P.g.G;
    // all-ones = not visited / not in the frontier / no children
    U4 ones;
    ones.x = ones.y = ones.z = ones.w = IMOMD_NONE;
    const uint32_t vec_per_rec = P.trec_stride / 16;
    if (P.dirty_words != 0)
    {
        const size_t dw = P.dirty_words;
        for (size_t i = (size_t)cta.tid(); i < K * dw; i += (size_t)cta.nt())
        {
            const size_t qt = (size_t)slot * K + i / dw;
            uint32_t* word = P.occ + qt * (2 * cells + dw) + 2 * cells + i % dw;
            uint32_t bits = *word;
            if (bits == 0) continue;
            *word = 0u;
            unsigned char* base = P.trec + qt * (size_t)P.g.n * P.trec_stride;
            const size_t v0 = (i % dw) * 32;
            while (bits != 0)
            {
                const int b = lowest_bit(bits);
                bits &= bits - 1;
                U4* r = reinterpret_cast<U4*>(base + (v0 + (size_t)b) * P.trec_stride);
                for (uint32_t j = 0; j < vec_per_rec; ++j) r[j] = ones;
            }
        }
    }
    else
    {
        // no bitmap: all K n records, with streaming 128-bit stores (they must not evict the
        // resident queries' working sets from L2)
        unsigned char* base = P.trec + (size_t)slot * K * (size_t)P.g.n * P.trec_stride;
        const size_t vecs = K * (size_t)P.g.n * P.trec_stride / 16;
        U4* dst = reinterpret_cast<U4*>(base);
        for (size_t i = (size_t)cta.tid(); i < vecs; i += (size_t)cta.nt()) cta.store_streaming(dst + i, ones);
    }
    uint32_t* ch = P.cell_head + (size_t)slot * K * cells;
    uint32_t* cc = P.cell_cnt + (size_t)slot * K * cells;
    for (size_t i = (size_t)cta.tid(); i < K * cells; i += (size_t)cta.nt())
    {
        ch[i] = IMOMD_NONE;
        cc[i] = 0u;
    }
    const size_t chunks = (size_t)P.chunks;
    uint32_t* fs = P.free_stack + (size_t)slot * K * chunks;
    for (size_t i = (size_t)cta.tid(); i < K * chunks; i += (size_t)cta.nt())
        fs[i] = (uint32_t)(chunks - 1 - i % chunks);   // pop order 0, 1, 2, ...
}

// All threads: root -> cell distance bounds of the slot for the roots of query hq
template <class Cta>
IMOMD_HD inline void fill_bounds(Cta& cta, const PlannerDev& P, int slot, int hq)
{
    const size_t cells = (size_t)P.g.G * P.g.G;
    const size_t total = (size_t)P.K * cells;
    double* lbd = P.lbd + (size_t)slot * total;
    double* ubd = P.ubd + (size_t)slot * total;
    const QueryDev* Q = P.query + hq;
    for (size_t i = (size_t)cta.tid(); i < total; i += (size_t)cta.nt())
    {
        const uint32_t root = Q->dest[i / cells];
        const uint32_t c = (uint32_t)(i % cells);
        lbd[i] = lbd_of(P.g, root, c);
        ubd[i] = ubd_of(P.g, root, c);
    }
}

IMOMD_HD inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

// one (tree, vertex, parent, cost bits) entry of the order-free tree checksum
IMOMD_HD inline uint64_t tree_sum_term(int k, uint32_t v, uint32_t parent, uint64_t cost_bits)
{
    return mix64(mix64(((uint64_t)(uint32_t)k << 32 | v) + 0x9E3779B97F4A7C15ull) ^
                 mix64(((uint64_t)parent << 1) + 1u) ^ mix64(cost_bits + 0xD6E8FEB86659FD93ull));
}

// TREEHASH of SURVEY.md Appendix A: FNV-1a-style fold over the trees in id order, each tree's
// visited vertices in ascending id, of the three 64-bit values (vertex, parent, cost bits).
// All threads compact a tile's visited records, in vertex order, into `buf` (u32 quadruples:
// vertex, parent, cost bits lo / hi); thread 0 folds them.  Also returns the order-free checksum.
// `buf` holds `cap` u32 (the slot's rewire arena).  Results are valid on thread 0.
constexpr uint32_t kHashPerThread = 8;
template <class Cta>
IMOMD_HD inline void tree_hash(Cta& cta, const PlannerDev& P, int slot, Shared& S, uint32_t* buf, uint32_t cap,
                               uint64_t* hash_out, uint64_t* sum_out)
{
    const uint32_t n = P.g.n;
    uint64_t h = 1469598103934665603ull;
    uint64_t sum = 0;
    uint32_t filled = 0;          // entries waiting in buf (uniform across the CTA)
    // thread 0 folds the waiting entries (the FNV chain is sequential)
    auto fold = [&]() {
        cta.sync();
        if (cta.tid() == 0)
        {
            for (uint32_t i = 0; i < filled; ++i)
            {
                const uint32_t* e = buf + 4 * (size_t)i;
                h = (h ^ (uint64_t)e[0]) * 1099511628211ull;
                h = (h ^ (uint64_t)e[1]) * 1099511628211ull;
                h = (h ^ (((uint64_t)e[3] << 32) | e[2])) * 1099511628211ull;
            }
        }
        filled = 0;
        cta.sync();
    };
    auto put = [&](int k, uint32_t at, uint32_t v, const U4& hd) {
        uint32_t* e = buf + 4 * (size_t)at;
        e[0] = v;
        e[1] = hd.x;
        e[2] = hd.z;
        e[3] = hd.w;
        sum += tree_sum_term(k, v, hd.x, ((uint64_t)hd.w << 32) | hd.z);
    };
    for (int k = 0; k < P.K; ++k)
    {
        const size_t qt = (size_t)slot * P.K + k;
        const unsigned char* base = P.trec + qt * (size_t)n * P.trec_stride;
        if (P.dirty_words != 0)
        {
            // streaming planner: only the records marked in the tree's dirty bitmap can be
            // visited; one bitmap word (32 vertices) per thread and tile, in vertex order
            const uint32_t dw = P.dirty_words;
            const uint32_t* bm = P.occ + qt * (2 * (size_t)P.g.G * P.g.G + dw) + 2 * (size_t)P.g.G * P.g.G;
            const uint32_t tile = (uint32_t)cta.nt();
            for (uint32_t w0 = 0; w0 < dw; w0 += tile)
            {
                const uint32_t wi = w0 + (uint32_t)cta.tid();
                uint32_t bits = wi < dw ? bm[wi] : 0u;
                uint32_t vis = 0;
                while (bits != 0)
                {
                    const int b = lowest_bit(bits);
                    bits &= bits - 1;
                    const uint32_t v = wi * 32u + (uint32_t)b;
                    if (*reinterpret_cast<const uint32_t*>(base + (size_t)v * P.trec_stride) != IMOMD_NONE) vis |= 1u << b;
                }
                uint32_t total = 0;
                uint32_t off = cta.scan_exclusive(count_bits(vis), &total, S.scan);
                while (vis != 0)
                {
                    const int b = lowest_bit(vis);
                    vis &= vis - 1;
                    const uint32_t v = wi * 32u + (uint32_t)b;
                    put(k, filled + off, v, *reinterpret_cast<const U4*>(base + (size_t)v * P.trec_stride));
                    ++off;
                }
                filled += total;
                // fold at the end of every tree and whenever the next tile might not fit
                if (w0 + tile >= dw || 4 * ((size_t)filled + (size_t)tile * 32) > cap) fold();
            }
            continue;
        }
        const uint32_t tile = (uint32_t)cta.nt() * kHashPerThread;
        for (uint32_t v0 = 0; v0 < n; v0 += tile)
        {
            U4 hd[kHashPerThread];
            uint32_t cnt = 0, total = 0;
            const uint32_t first = v0 + (uint32_t)cta.tid() * kHashPerThread;
#pragma unroll
            for (uint32_t j = 0; j < kHashPerThread; ++j)
            {
                hd[j].x = IMOMD_NONE;
                if (first + j < n)
                    hd[j] = cta.load_streaming(reinterpret_cast<const U4*>(base + (size_t)(first + j) * P.trec_stride));
                if (hd[j].x != IMOMD_NONE) ++cnt;
            }
            uint32_t off = cta.scan_exclusive(cnt, &total, S.scan);
#pragma unroll
            for (uint32_t j = 0; j < kHashPerThread; ++j)
            {
                if (hd[j].x != IMOMD_NONE)
                {
                    put(k, filled + off, first + j, hd[j]);
                    ++off;
                }
            }
            filled += total;
            // fold at the end of every tree and whenever the next tile might not fit
            if (v0 + tile >= n || 4 * (size_t)(filled + tile) > cap) fold();
        }
    }
    sum = cta.reduce_sum_u64(sum, S.red);
    *hash_out = h;
    *sum_out = sum;
}

// The per-query driver: the loop every CTA runs (device) / the host simulation runs.
// q: the slot holding the query's trees and scratch; hq: the query (header / report index).
template <int KW, class Cta>
IMOMD_HD inline void run_query(Cta& cta, const PlannerDev& P, int q, int hq, int iters, Shared& S,
                               double* lb, uint32_t* lvl, double* mat, unsigned char* hot,
                               int attach_tree = -1, uint32_t attach_count = 0)
{
    assume_planner(P);
    IMOMD_ASSUME_SHARED(&S);
    IMOMD_ASSUME_SHARED(lb);
    IMOMD_ASSUME_SHARED(lvl);
    IMOMD_ASSUME_SHARED(mat);
    IMOMD_ASSUME_SHARED(hot);
    QueryDev* Qg = P.query + hq;
    QueryDev* Q = reinterpret_cast<QueryDev*>(hot);
    const int KK = P.K * P.K;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(Qg);
        uint32_t* dst = reinterpret_cast<uint32_t*>(hot);
        for (uint32_t i = cta.tid(); i < IMOMD_QUERY_HOT_BYTES / 4; i += cta.nt()) dst[i] = src[i];
    }
    TreeRef* tref = reinterpret_cast<TreeRef*>(hot + ((IMOMD_QUERY_HOT_BYTES + 15) / 16) * 16);
    for (int k = cta.tid(); k < P.K; k += cta.nt()) tref[k] = tree_ref(P, q, k);
    if (cta.tid() == 0)
    {
        S.Q = Q;
        S.arena = P.arena + (size_t)q * P.arena_entries;
        S.arena_cap = P.arena_entries;
        S.clist = P.clist + (size_t)q * 2 * P.chunks;
        S.batch = 0;
        for (int i = 0; i < 8; ++i) S.stat[i] = 0;
        S.near_ties = S.unresolved_ties = S.certified_ties = 0;
        S.tref = tref;
        S.h = reinterpret_cast<double*>(tref + P.K);
        S.mP = mat;
        S.mD = mat + KK;
        S.mMHv = mat + 2 * KK;
        S.mMHi = reinterpret_cast<uint32_t*>(mat + 3 * KK);
        S.mCN = S.mMHi + KK;
    }
    cta.sync();
    for (int i = cta.tid(); i < KK; i += cta.nt())
    {
        sm_P(S)[i] = Qg->P[i];
        sm_D(S)[i] = Qg->D[i];
        sm_MHv(S)[i] = Qg->MHv[i];
        sm_MHi(S)[i] = Qg->MHi[i];
        sm_CN(S)[i] = Qg->CN[i];
    }
    if (cta.tid() == 0)
    {
        S.phase = PH_NEXT_TREE;
        S.it = 0;
        S.stop = 0;
        S.t_start_ns = P.budget_ns ? cta.now_ns() : 0;
        S.k = Q->resume_k;         // resume point inside an interrupted pass
        S.attach_mode = 0;
        if (attach_tree >= 0)
        {
            // mergeTwoTree_ (:848-851): connectNewNode_, updateExpandables, rewireTree_,
            // updateConnectionTree_ of the listed vertices, in order, in one given tree
            S.attach_mode = 1;
            S.k = attach_tree;
            S.attach_i = 0;
            S.attach_n = attach_count;
            S.x_new = P.attach_list[0];
            S.conn_pending = 0;
            S.phase = PH_JPS;
        }
        S.req = REQ_EXIT;
        if (Q->status == 1) Q->status = 0;   // NEED_WORDS is resumable
        Q->last_iterations = Q->iterations;
        Q->last_expansions = Q->expansions;
    }
    cta.sync();
    const long long busy0 = cta.clock();
    uint32_t attach_seen = IMOMD_NONE;
    int warm_k = -1;              // tree whose nearest search has just finished (helpers prefetch)
    uint32_t warm_x = IMOMD_NONE;
    const bool prof = P.profile != 0;   // per-request cycle counters (imomd_set_profiling): development aid
    for (;;)
    {
        long long t0 = prof ? cta.clock() : 0;
        if (warm_k >= 0) warm_x = S.result.id1;
        // attach mode: warm the upcoming list entries once per entry (when the position moved)
        uint32_t attach_at = IMOMD_NONE;
        if (attach_tree >= 0 && S.attach_i != attach_seen)
        {
            attach_at = S.attach_i;
            attach_seen = attach_at;
        }
        if (cta.tid() == 0)
        {
            int ph0 = S.phase & 7;
            Seq<KW> seq(P, q, S);
            seq.step(cta, iters);
            if (prof)
            {
                Q->phase_cycles[ph0] += (uint64_t)(cta.clock() - t0);
                Q->phase_count[ph0] += 1;
            }
        }
        else if (warm_k >= 0)
        {
            // a nearest search has just finished: fetch what the serial thread is about to touch
            warm_expansion(cta, P, tree_ref_of(P, q, warm_k, S), q, warm_k, warm_x, S);
        }
        else if (attach_tree >= 0 && attach_at != IMOMD_NONE)
        {
            warm_attach(cta, P, tree_ref_of(P, q, attach_tree, S), attach_at, attach_count, S);
        }
        cta.sync();
        int req = S.req;
        int k = S.k;
        warm_k = -1;
        long long t1 = prof ? cta.clock() : 0;
        if (prof && cta.tid() == 0)
        {
            Q->prof_cycles[7] += (uint64_t)(t1 - t0);
            Q->prof_count[7] += 1;
        }
        if (req == REQ_EXIT) break;
        if (req == REQ_NN)
        {
            svc_nearest(cta, P, q, k, S, Q->n_occ[k] <= lb_entries(P) ? lb : nullptr, S.clist, &S.near_ties,
                        &S.unresolved_ties, true);
            warm_k = k;
        }
        else if (req == REQ_RESCAN) svc_rescan(cta, P, q, k, S, Q->n_occ[k] <= lb_entries(P) ? lb : nullptr);
        else if (req == REQ_MHINS) svc_mhins(cta, P, q, k, S);
        else if (req == REQ_CONN) svc_conn_gather(cta, P, q, S.conn_x, S);
        else if (req == REQ_BFS)
        {
            svc_bfs_w<KW>(cta, P, q, k, S, lvl);
            cta.sync();
            // the serial thread would ask for this next: examine the fresh list's calls (the
            // request was filled in by svc_bfs_w's last step)
            if (S.bfs_hi != IMOMD_NONE) svc_check<KW>(cta, P, q, k, S);
        }
        else if (req == REQ_CHECK) svc_check<KW>(cta, P, q, k, S);
        cta.sync();
        if (prof && cta.tid() == 0)
        {
            Q->prof_cycles[req & 7] += (uint64_t)(cta.clock() - t1);
            Q->prof_count[req & 7] += 1;
        }
    }
    for (int i = cta.tid(); i < KK; i += cta.nt())
    {
        Qg->P[i] = sm_P(S)[i];
        Qg->D[i] = sm_D(S)[i];
        Qg->MHv[i] = sm_MHv(S)[i];
        Qg->MHi[i] = sm_MHi(S)[i];
        Qg->CN[i] = sm_CN(S)[i];
    }
    if (cta.tid() == 0)
    {
        if (attach_tree < 0) Q->resume_k = S.k < Q->K ? S.k : 0;
        for (int i = 0; i < 8; ++i) Q->stat[i] += S.stat[i];
        Q->near_ties += S.near_ties;
        Q->unresolved_ties += S.unresolved_ties;
        Q->certified_ties += S.certified_ties;
        Q->busy_cycles = (uint64_t)(cta.clock() - busy0);
        Q->last_iterations = Q->iterations - Q->last_iterations;
        Q->last_expansions = Q->expansions - Q->last_expansions;
        int active = 0;
        int64_t tv = 0;
        for (int j = 0; j < Q->K; ++j)
        {
            if (Q->frontier_count[j] != 0 && !Q->is_done[j]) ++active;
            tv += (int64_t)Q->tree_size[j];
        }
        Q->active_trees = active;
        ReportDev* R = P.reports + hq;
        R->status = Q->status;
        R->active_trees = active;
        R->iterations = Q->last_iterations;
        R->expansions = Q->last_expansions;
        R->tree_vertices = tv;
        R->words_consumed = (int64_t)Q->word_cursor;
        R->connected = Q->connected;
        R->dm_updated = Q->dm_updated;
        R->near_ties = Q->near_ties;
        R->unresolved_ties = Q->unresolved_ties;
        R->log_entries = (int64_t)Q->log_count;
        R->certified_ties = Q->certified_ties;
        R->tree_hash = Q->tree_hash;
        R->busy_cycles = Q->busy_cycles;
    }
    cta.sync();
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(hot);
        uint32_t* dst = reinterpret_cast<uint32_t*>(Qg);
        for (uint32_t i = cta.tid(); i < IMOMD_QUERY_HOT_BYTES / 4; i += cta.nt()) dst[i] = src[i];
    }
    cta.sync();
}

// One query of a streaming planner, start to finish, in slot `slot`: header from the freshly
// constructed copy, slot rebuilt, trees constructed (ctor :73-78), `iters` passes, tree hash.
template <int KW, class Cta>
IMOMD_HD inline void run_streaming_query(Cta& cta, const PlannerDev& P, int slot, int hq, int iters, Shared& S,
                                         double* lb, uint32_t* lvl, double* mat, unsigned char* hot)
{
    const long long t0 = cta.clock();
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(P.query_init + hq);
        uint32_t* dst = reinterpret_cast<uint32_t*>(P.query + hq);
        for (uint32_t i = cta.tid(); i < sizeof(QueryDev) / 4; i += cta.nt()) dst[i] = src[i];
    }
    reset_slot(cta, P, slot);
    fill_bounds(cta, P, slot, hq);
    cta.sync();
    if (cta.tid() == 0)
    {
        QueryDev* Q = P.query + hq;
        for (int k = 0; k < P.K; ++k) Q->free_top[k] = (uint32_t)P.chunks;
        construct_trees(P, slot, hq);
    }
    cta.sync();
    run_query<KW>(cta, P, slot, hq, iters, S, lb, lvl, mat, hot);
    cta.sync();
    uint64_t h = 0, sum = 0;
    tree_hash(cta, P, slot, S, P.arena + (size_t)slot * P.arena_entries, P.arena_entries, &h, &sum);
    if (cta.tid() == 0)
    {
        QueryDev* Q = P.query + hq;
        Q->tree_hash = h;
        Q->tree_sum = sum;
        Q->busy_cycles = (uint64_t)(cta.clock() - t0);
        P.reports[hq].tree_hash = h;
        P.reports[hq].busy_cycles = Q->busy_cycles;
    }
    cta.sync();
}

// ====================================================== batched expansion ====
// imomd_expand_batch: B expansions per tree per STEP (north_star (2), (4)).  A step has four parts:
//
//  1. ATTACH, the K trees of a query CONCURRENTLY, one cooperating group (a warp on the device)
//     per tree taking trees from a counter: sample, nearest frontier vertex, jump-point hops,
//     choose-parent, frontier / min-heuristic maintenance -- the sequential code of the exact
//     mode on the tree's own records and its own row of the heuristic matrix.  The new vertex is
//     only QUEUED for rewiring.  Samples come from fixed slots of kBatchWords words per (pass,
//     tree) of the injected stream, so that trees draw independently.
//  2. REWIRE, the whole CTA over all trees at once: a label-correcting relaxation from the queued
//     vertices.  A vertex whose cost went down offers cost + w to its visited neighbours; a
//     neighbour takes an offer when its label (cost, parent id) gets lexicographically smaller --
//     one 128-bit compare-and-swap on the {parent, frontier slot, cost} head of its record.  The
//     fixed point (every visited vertex at its cheapest cost through visited vertices, lowest
//     parent id among equal offers) does not depend on the order in which offers arrive: the
//     deterministic resolution of concurrent cost / parent updates.  This replaces the recursive
//     cascade of rewireTree_ (:561-600), whose order is inherently sequential.
//  3. CONNECT: the connection check (updateConnectionTree_ :645-706) of every vertex attached or
//     changed in the step against the other trees, with the final costs of the step: per pair of
//     trees the smallest cost sum (lowest vertex id on ties) replaces the distance when it beats
//     it by more than 0.1 (:680).
//  4. The selection-probability tests (:602-643) in tree order and program order, with the
//     heuristic entry logged at the time of the expansion.
//
// Deterministic: the result depends on neither the number of groups nor their timing.  Child rows
// are not maintained (the relaxation walks the road graph, not the tree), so a planner that has
// run batched steps cannot continue in the exact mode before a reset.
constexpr uint32_t kBatchWords = 8;
// (the exact mode's CTA keeps up to kCtaLbCap of them: lb_entries)
constexpr uint32_t kGroupLbCap = 256;   // cell lower bounds a group keeps in shared memory between the two
                                        // passes of a pruned search (frontiers with more occupied cells
                                        // recompute them)

struct BatchShared
{
    uint32_t ev_count[IMOMD_KMAX];
    uint32_t exp_count[IMOMD_KMAX];
    int32_t active[IMOMD_KMAX];
    int32_t n_active;
    int32_t stop;
    uint32_t next_tree;                  // trees are handed to the groups from this counter
    uint32_t relax_items, relax_changes, relax_levels;   // work counters of the step
    uint32_t dirty_n[2][IMOMD_KMAX];     // relaxation queues (ping-pong) per tree
    uint32_t chk_n[IMOMD_KMAX];          // vertices whose connection check is due
    uint64_t tree_cycles[IMOMD_KMAX];    // cycles the tree's group spent on its batch in this step
    uint64_t relax_cycles[IMOMD_KMAX];   // ... and on the tree's relaxation
    // connection candidates of the step per unordered pair (a > b): index a(a-1)/2 + b
    // (K (K - 1) / 2 entries each; they live in the groups' cell-bound caches, which are idle
    // while the block runs the connection checks)
    unsigned long long* cand_d;
    uint32_t* cand_x;
};

// the scratch of tree k inside the slot's arena: [rewire queue A | rewire queue B | check list].
// (Measured: queue entries that carry the cost they were queued with save the dependent load of the
// vertex's own record but make offers from costs that are already out of date -- 60 % more offer
// rounds, the relaxation 10 % slower.  An offer round reads the vertex's CURRENT cost.)
IMOMD_HD inline uint32_t batch_share(const PlannerDev& P) { return P.arena_entries / (uint32_t)P.K; }
IMOMD_HD inline uint32_t batch_queue_cap(const PlannerDev& P) { return batch_share(P) / 4; }
IMOMD_HD inline uint32_t* batch_queue(const PlannerDev& P, int slot, int k, int which)
{
    const uint32_t share = batch_share(P);
    return P.arena + (size_t)slot * P.arena_entries + (size_t)k * share + (size_t)which * (share / 4);
}
IMOMD_HD inline uint32_t* batch_checks(const PlannerDev& P, int slot, int k)
{
    const uint32_t share = batch_share(P);
    return P.arena + (size_t)slot * P.arena_entries + (size_t)k * share + (size_t)(share / 2);
}

// One group: up to B attaches of tree k (fewer when its frontier runs empty).
template <int KW, class Grp>
IMOMD_HD inline void batch_expand_tree(Grp& grp, const PlannerDev& P, int slot, int k, Shared& S, double* lbg, int B,
                                       uint64_t pass0)
{
    const int K = P.K;
    if (grp.tid() == 0)
    {
        const uint32_t share = batch_share(P);
        S.arena = P.arena + (size_t)slot * P.arena_entries;   // (no rewire frames in this mode)
        S.arena_cap = 0;
        S.clist = P.clist_tree + ((size_t)slot * K + k) * 2 * (size_t)P.chunks;
        S.ev = P.events + ((size_t)slot * K + k) * (size_t)P.ev_cap;
        S.ev_n = 0;
        S.exp_n = 0;
        S.dirty = batch_queue(P, slot, k, 0);
        S.dirty_n = 0;
        S.dirty_cap = batch_queue_cap(P);
        S.chk = batch_checks(P, slot, k);
        S.chk_n = 0;
        S.chk_cap = share / 2;
        S.batch = 1;
        S.attach_mode = 0;
        S.k = k;
    }
    grp.sync();
    for (int b = 0; b < B; ++b)
    {
        if (grp.tid() == 0)
        {
            QueryDev* Q = S.Q;
            if (Q->frontier_count[k] == 0 || Q->status != 0) S.req = REQ_EXIT;
            else
            {
                Seq<KW> seq(P, slot, S);
                TreeRef T = tree_ref_of(P, slot, k, S);
                seq.sample_at(k, T, ((pass0 + (uint64_t)b) * (uint64_t)K + (uint64_t)k) * kBatchWords);
                S.req = Q->status == 0 ? REQ_NN : REQ_EXIT;
                S.phase = PH_AFTER_NN;
                S.conn_pending = 0;
            }
        }
        grp.sync();
        if (S.req == REQ_EXIT) break;
        for (;;)
        {
            const int req = S.req;
            if (req == REQ_EXIT) break;
            const bool prof = P.profile != 0;
            const long long t0 = prof ? grp.clock() : 0;
            double* lb = (lbg != nullptr && S.Q->n_occ[k] <= kGroupLbCap) ? lbg : nullptr;
            if (req == REQ_NN) svc_nearest(grp, P, slot, k, S, lb, S.clist, &S.near_ties, &S.unresolved_ties, true);
            else if (req == REQ_RESCAN) svc_rescan(grp, P, slot, k, S, lb);
            else if (req == REQ_MHINS) svc_mhins(grp, P, slot, k, S);
            grp.sync();
            if (grp.tid() == 0)
            {
                const long long t1 = prof ? grp.clock() : 0;
                if (prof) S.gprof[req & 7] += (uint64_t)(t1 - t0);
                Seq<KW> seq(P, slot, S);
                seq.step(grp, 1);
                if (prof) S.gprof[7] += (uint64_t)(grp.clock() - t1);
            }
            grp.sync();
        }
    }
    grp.sync();
}

// One offer round of the relaxation for vertex u of tree T: u's cost + w to every visited
// neighbour, accepted when (cost, parent) gets lexicographically smaller.  A neighbour whose COST
// went down joins the next queue; every neighbour that changed joins the check list.
// Measured alternatives (bench workload, B = 16, relaxation 187-194 ms per query as it stands: a
// level costs about 5.9 k cycles + 1.2 k per vertex, 188 levels per step): all first swaps of a
// vertex in flight together + rotating queues with one barrier per level -- no change (194 ms);
// queue entries that carry their cost -- offers from costs already out of date, 60 % more rounds,
// 210 ms; the thread following the chain itself instead of queueing (depth-first order) -- 75 %
// more accepted offers, 231 ms.  Breadth-first levels are the cheap order for this relaxation.
template <int W, class Blk>
IMOMD_HD inline void relax_vertex(Blk& blk, const PlannerDev& P, const TreeRef& T, uint32_t u,
                                  uint32_t* next_q, uint32_t* next_n, uint32_t q_cap, uint32_t* chk, uint32_t* chk_n,
                                  uint32_t chk_cap, QueryDev* Q, uint32_t* changes)
{
    const GraphDev& g = P.g;
    const uint32_t e0 = row_begin(g, u), d = g.deg[u];
    // (heads are read past the L1: other threads change them with compare-and-swap at the L2)
    const double cu = head_from(blk.load_cg(reinterpret_cast<const U4*>(T.hp(u)))).cost;
    uint32_t ys[W];
    double ws[W];
    U4 hs[W];
#pragma unroll
    for (int i = 0; i < W; ++i)
    {
        if ((uint32_t)i < d)
        {
            ys[i] = g.col[e0 + i];
            ws[i] = g.w[e0 + i];
        }
    }
#pragma unroll
    for (int i = 0; i < W; ++i)
    {
        if ((uint32_t)i < d) hs[i] = blk.load_cg(reinterpret_cast<const U4*>(T.hp(ys[i])));
    }
#pragma unroll
    for (int i = 0; i < W; ++i)
    {
        if ((uint32_t)i >= d) continue;
        const uint32_t y = ys[i];
        const double p = cu + ws[i];
        U4 cur = hs[i];
        for (;;)
        {
            if (cur.x == IMOMD_NONE) break;                  // not visited
            unsigned long long bits = ((unsigned long long)cur.w << 32) | (unsigned long long)cur.z;
            double cy;
            memcpy(&cy, &bits, 8);
            const bool lower = p < cy;
            if (!(lower || (p == cy && u < cur.x))) break;
            unsigned long long pb;
            memcpy(&pb, &p, 8);
            U4 want;
            want.x = u;
            want.y = cur.y;
            want.z = (uint32_t)pb;
            want.w = (uint32_t)(pb >> 32);
            const U4 seen = blk.cas_head(reinterpret_cast<U4*>(T.hp(y)), cur, want);
            if (seen.x == cur.x && seen.y == cur.y && seen.z == cur.z && seen.w == cur.w)
            {
                blk.counter_add(changes, 1u);
                uint32_t at = blk.counter_add(chk_n, 1u);
                if (at < chk_cap) chk[at] = y;
                else Q->status = 3;
                if (lower)
                {
                    // y's own offer round comes a level later: its adjacency row on the way to the L2
                    blk.prefetch_l2(g.col + row_begin(g, y));
                    blk.prefetch_l2(g.w + row_begin(g, y));
                    at = blk.counter_add(next_n, 1u);
                    if (at < q_cap) next_q[at] = y;
                    else Q->status = 3;
                }
                break;
            }
            cur = seen;
        }
    }
}

// One group: the rewiring relaxation of tree k to its fixed point, level by level, right after the
// tree's attaches.  A tree's records are touched by nobody else during the step, so the trees relax
// independently of each other: no block barrier per level, and a deep chain in one tree does not
// hold the other trees' levels back (the fixed point does not depend on the order of the offers).
// Returns the cycles spent.
template <int KW, class Grp>
IMOMD_HD inline void batch_relax_tree(Grp& grp, const PlannerDev& P, int slot, int k, Shared& S, BatchShared& BS)
{
    QueryDev* Q = S.Q;
    const uint32_t q_cap = batch_queue_cap(P), chk_cap = batch_share(P) / 2;
    const TreeRef T = tree_ref_of(P, slot, k, S);
    uint32_t* chk = batch_checks(P, slot, k);
    uint32_t items = 0, levels = 0;
    int cur = 0;
    for (;;)
    {
        uint32_t n = BS.dirty_n[cur][k];   // (uniform across the group: the level being read is complete)
        n = n < q_cap ? n : q_cap;
        if (n == 0) break;
        const uint32_t* queue = batch_queue(P, slot, k, cur);
        for (uint32_t i = (uint32_t)grp.tid(); i < n; i += (uint32_t)grp.nt())
            relax_vertex<KW>(grp, P, T, queue[i], batch_queue(P, slot, k, cur ^ 1), &BS.dirty_n[cur ^ 1][k], q_cap, chk,
                             &BS.chk_n[k], chk_cap, Q, &BS.relax_changes);
        grp.sync();
        if (grp.tid() == 0) BS.dirty_n[cur][k] = 0;
        items += n;
        levels += 1;
        cur ^= 1;
        grp.sync();
    }
    if (grp.tid() == 0 && levels != 0)
    {
        grp.counter_add(&BS.relax_items, items);
        grp.counter_add(&BS.relax_levels, levels);
    }
}

IMOMD_HD inline int pair_index(int a, int b) { return a > b ? a * (a - 1) / 2 + b : b * (b - 1) / 2 + a; }

// The whole CTA: connection checks of the step's attached / changed vertices with the final costs.
template <int KW, class Blk>
IMOMD_HD inline void batch_connect(Blk& blk, const PlannerDev& P, int slot, Shared& S0, BatchShared& BS)
{
    QueryDev* Q = S0.Q;
    const int K = P.K;
    const int pairs = K * (K - 1) / 2;
    const uint32_t chk_cap = batch_share(P) / 2;
    for (int i = blk.tid(); i < pairs; i += blk.nt())
    {
        BS.cand_d[i] = ~0ull;
        BS.cand_x[i] = IMOMD_NONE;
    }
    blk.sync();
    uint32_t total = 0;
    for (int k = 0; k < K; ++k) total += BS.chk_n[k] < chk_cap ? BS.chk_n[k] : chk_cap;
    for (int pass = 0; pass < 2; ++pass)
    {
        // pass 0: the smallest sum per pair; pass 1: the lowest vertex that attains it
        for (uint32_t i = (uint32_t)blk.tid(); i < total; i += (uint32_t)blk.nt())
        {
            int k = 0;
            uint32_t j = i;
            for (;;)
            {
                uint32_t n = BS.chk_n[k] < chk_cap ? BS.chk_n[k] : chk_cap;
                if (j < n) break;
                j -= n;
                ++k;
            }
            const uint32_t x = batch_checks(P, slot, k)[j];
            const double ck = head_from(blk.load_cg(reinterpret_cast<const U4*>(tree_ref_of(P, slot, k, S0).hp(x)))).cost;
            for (int o = 0; o < K; ++o)
            {
                if (o == k) continue;
                const THead ho = head_from(blk.load_cg(reinterpret_cast<const U4*>(tree_ref_of(P, slot, o, S0).hp(x))));
                if (ho.parent == IMOMD_NONE) continue;
                const double dsum = ck + ho.cost;
                unsigned long long bits;
                memcpy(&bits, &dsum, 8);
                const int pi = pair_index(k, o);
                if (pass == 0) blk.atomic_min_u64(&BS.cand_d[pi], bits);
                else if (bits == BS.cand_d[pi]) blk.atomic_min_u32(&BS.cand_x[pi], x);
            }
        }
        blk.sync();
    }
    if (blk.tid() == 0)
    {
        S0.stat[6] += total;
        Seq<KW> seq(P, slot, S0);
        for (int a = 0; a < K; ++a)
        {
            for (int b = 0; b < a; ++b)
            {
                const int pi = a * (a - 1) / 2 + b;
                if (BS.cand_d[pi] == ~0ull) continue;
                double best;
                memcpy(&best, &BS.cand_d[pi], 8);
                if (!Q->contact[pi])
                {
                    Q->contact[pi] = 1;
                    seq.connect_two_trees(b, a);
                }
                if (best < sm_D(S0)[a * K + b] - 0.1)
                {
                    sm_D(S0)[a * K + b] = best;
                    sm_D(S0)[b * K + a] = best;
                    sm_CN(S0)[a * K + b] = BS.cand_x[pi];
                    sm_CN(S0)[b * K + a] = BS.cand_x[pi];
                    Q->dm_updated = 1;
                }
            }
        }
    }
    blk.sync();
}

// One thread: the logged selection-probability tests of the step, trees in id order, each tree's
// tests in the order its expansions produced them.
template <int KW, class Blk>
IMOMD_HD inline void batch_apply_tests(Blk& blk, const PlannerDev& P, int slot, Shared& S0, BatchShared& BS)
{
    QueryDev* Q = S0.Q;
    const int K = P.K;
    if (blk.tid() == 0)
    {
        Seq<KW> seq(P, slot, S0);
        for (int k = 0; k < K; ++k)
        {
            const uint32_t n = BS.ev_count[k];
            const BEvent* ev = P.events + ((size_t)slot * K + k) * (size_t)P.ev_cap;
            for (uint32_t e = 0; e < n; ++e)
            {
                const BEvent E = ev[e];
                seq.selection_probability_apply(blk, k, (int)E.aux, E.val, E.extra);
            }
            Q->expansions += (int64_t)BS.exp_count[k];
        }
    }
    blk.sync();
}

// `steps` steps of B expansions per tree for query hq in slot `slot`.  SG: one Shared per group.
template <int KW, class Blk, class Grp>
IMOMD_HD inline void run_query_batched(Blk& blk, Grp& grp, const PlannerDev& P, int slot, int hq, int steps, int B,
                                       Shared* SG, BatchShared& BS, double* lb_all, double* mat, unsigned char* hot)
{
    assume_planner(P);
    QueryDev* Qg = P.query + hq;
    QueryDev* Q = reinterpret_cast<QueryDev*>(hot);
    const int K = P.K, KK = K * K;
    const int G = grp.n_groups();
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(Qg);
        uint32_t* dst = reinterpret_cast<uint32_t*>(hot);
        for (uint32_t i = blk.tid(); i < IMOMD_QUERY_HOT_BYTES / 4; i += blk.nt()) dst[i] = src[i];
    }
    TreeRef* tref = reinterpret_cast<TreeRef*>(hot + ((IMOMD_QUERY_HOT_BYTES + 15) / 16) * 16);
    for (int k = blk.tid(); k < K; k += blk.nt()) tref[k] = tree_ref(P, slot, k);
    if (blk.tid() == 0)
    {
        // (kGroupLbCap doubles per group: 2 KB at least, K (K - 1) / 2 <= 496 pairs of 12 bytes need 6 KB
        // of the G >= 4 groups' 8 KB; the host simulation sizes its one buffer accordingly)
        BS.cand_d = reinterpret_cast<unsigned long long*>(lb_all);
        BS.cand_x = reinterpret_cast<uint32_t*>(BS.cand_d + K * (K - 1) / 2);
    }
    for (int g = blk.tid(); g < G; g += blk.nt())
    {
        Shared& S = SG[g];
        S.Q = Q;
        S.tref = tref;
        S.h = reinterpret_cast<double*>(tref + K) + (size_t)g * P.g.width * K;
        S.mP = mat;
        S.mD = mat + KK;
        S.mMHv = mat + 2 * KK;
        S.mMHi = reinterpret_cast<uint32_t*>(mat + 3 * KK);
        S.mCN = S.mMHi + KK;
        S.batch = 1;
        S.attach_mode = 0;
        S.stop = 0;
        S.it = 0;
        for (int i = 0; i < 8; ++i) S.stat[i] = 0;
        for (int i = 0; i < 8; ++i) S.gprof[i] = 0;
        S.near_ties = S.unresolved_ties = S.certified_ties = 0;
        S.arena = P.arena + (size_t)slot * P.arena_entries;
        S.arena_cap = 0;
        S.clist = nullptr;
        S.ev = nullptr;
        S.ev_n = S.exp_n = 0;
        S.dirty = S.chk = nullptr;
        S.dirty_n = S.dirty_cap = S.chk_n = S.chk_cap = 0;
    }
    blk.sync();
    Shared& S0 = SG[0];
    for (int i = blk.tid(); i < KK; i += blk.nt())
    {
        sm_P(S0)[i] = Qg->P[i];
        sm_D(S0)[i] = Qg->D[i];
        sm_MHv(S0)[i] = Qg->MHv[i];
        sm_MHi(S0)[i] = Qg->MHi[i];
        sm_CN(S0)[i] = Qg->CN[i];
    }
    if (blk.tid() == 0)
    {
        if (Q->status == 1) Q->status = 0;
        Q->last_iterations = Q->iterations;
        Q->last_expansions = Q->expansions;
    }
    blk.sync();
    const long long busy0 = blk.clock();
    const uint64_t t_start = P.budget_ns ? blk.now_ns() : 0;
    for (int step = 0; step < steps; ++step)
    {
        if (blk.tid() == 0)
        {
            int active = 0;
            for (int k = 0; k < K; ++k)
            {
                BS.active[k] = (Q->frontier_count[k] != 0 && !Q->is_done[k]) ? 1 : 0;
                BS.ev_count[k] = 0;
                BS.exp_count[k] = 0;
                BS.dirty_n[0][k] = BS.dirty_n[1][k] = 0;
                BS.chk_n[k] = 0;
                active += BS.active[k];
            }
            BS.n_active = active;
            BS.next_tree = 0;
            BS.relax_items = BS.relax_changes = BS.relax_levels = 0;
            BS.stop = 0;
            if (Q->status != 0) BS.stop = 1;
            else if (active == 0)
            {
                Q->iterations += (int64_t)(steps - step) * B;   // nothing can expand any more
                BS.stop = 1;
            }
            else if (((uint64_t)Q->iterations + (uint64_t)B) * (uint64_t)K * kBatchWords + IMOMD_WORD_MARGIN > P.words_avail)
            {
                Q->status = 1;   // IMOMD_RUN_NEED_WORDS
                BS.stop = 1;
            }
            else if (P.budget_ns && step > 0 && blk.now_ns() - t_start > P.budget_ns) BS.stop = 1;
        }
        blk.sync();
        if (BS.stop) break;
        const uint64_t pass0 = (uint64_t)Q->iterations;
        blk.sync();
        const long long ts0 = blk.clock();
        // 1. attach: the groups take trees from the counter
        Shared& S = SG[grp.group()];
        for (;;)
        {
            uint32_t k = 0;
            if (grp.tid() == 0) k = blk.counter_add(&BS.next_tree, 1u);
            k = grp.bcast(k);
            if (k >= (uint32_t)K) break;
            if (!BS.active[k]) continue;
            const long long tk0 = grp.clock();
            batch_expand_tree<KW>(grp, P, slot, (int)k, S, lb_all ? lb_all + (size_t)grp.group() * kGroupLbCap : nullptr, B,
                                  pass0);
            const long long tk1 = grp.clock();
            if (grp.tid() == 0)
            {
                BS.ev_count[k] = S.ev_n;
                BS.exp_count[k] = S.exp_n;
                BS.dirty_n[0][k] = S.dirty_n;
                BS.chk_n[k] = S.chk_n;
                BS.tree_cycles[k] = (uint64_t)(tk1 - tk0);
            }
            grp.sync();
            // 2. rewire: the tree's relaxation, by the same group
            batch_relax_tree<KW>(grp, P, slot, (int)k, S, BS);
            if (grp.tid() == 0) BS.relax_cycles[k] = (uint64_t)(grp.clock() - tk1);
        }
        blk.sync();
        blk.fence();     // heads changed by compare-and-swap at the L2; the next reads go through the L1
        blk.sync();
        const long long ts2 = blk.clock();
        // 3. connect, 4. selection-probability tests
        batch_connect<KW>(blk, P, slot, S0, BS);
        batch_apply_tests<KW>(blk, P, slot, S0, BS);
        if (blk.tid() == 0)
        {
            Q->iterations += B;
            Q->word_cursor = (uint64_t)Q->iterations * (uint64_t)K * kBatchWords;
            S0.stat[2] += BS.relax_items;
            S0.stat[3] += (uint64_t)BS.relax_items * (uint64_t)P.g.width;
            S0.stat[4] += BS.relax_changes;
            // where a step goes (imomd_get_profile): 1 = the trees' attach batches summed (group
            // cycles), 2 = the slowest tree's attach + relaxation, 3 = attach + relaxation as the
            // block saw them, 4 = the trees' relaxations summed (group cycles),
            // 5 = connection checks + selection-probability tests
            uint64_t sum = 0, mx = 0, rsum = 0;
            for (int k = 0; k < K; ++k)
            {
                if (!BS.active[k]) continue;
                sum += BS.tree_cycles[k];
                rsum += BS.relax_cycles[k];
                if (BS.tree_cycles[k] + BS.relax_cycles[k] > mx) mx = BS.tree_cycles[k] + BS.relax_cycles[k];
                Q->prof_count[1] += 1;
            }
            Q->prof_cycles[1] += sum;
            Q->prof_cycles[2] += mx;
            Q->prof_count[2] += 1;
            Q->prof_cycles[3] += (uint64_t)(ts2 - ts0);
            Q->prof_count[3] += BS.relax_levels;
            Q->prof_cycles[4] += rsum;
            Q->prof_count[4] += BS.relax_items;
            Q->prof_cycles[5] += (uint64_t)(blk.clock() - ts2);
            Q->prof_count[5] += 1;
        }
        blk.sync();
    }
    for (int i = blk.tid(); i < KK; i += blk.nt())
    {
        Qg->P[i] = sm_P(S0)[i];
        Qg->D[i] = sm_D(S0)[i];
        Qg->MHv[i] = sm_MHv(S0)[i];
        Qg->MHi[i] = sm_MHi(S0)[i];
        Qg->CN[i] = sm_CN(S0)[i];
    }
    if (blk.tid() == 0)
    {
        for (int g = 0; g < G; ++g)
        {
            for (int i = 0; i < 8; ++i) Q->stat[i] += SG[g].stat[i];
            Q->near_ties += SG[g].near_ties;
            Q->unresolved_ties += SG[g].unresolved_ties;
            Q->certified_ties += SG[g].certified_ties;
            for (int i = 0; i < 8; ++i) Q->phase_cycles[i] += SG[g].gprof[i];
        }
        Q->resume_k = 0;
        Q->busy_cycles = (uint64_t)(blk.clock() - busy0);
        Q->last_iterations = Q->iterations - Q->last_iterations;
        Q->last_expansions = Q->expansions - Q->last_expansions;
        int active = 0;
        int64_t tv = 0;
        for (int j = 0; j < K; ++j)
        {
            if (Q->frontier_count[j] != 0 && !Q->is_done[j]) ++active;
            tv += (int64_t)Q->tree_size[j];
        }
        Q->active_trees = active;
        ReportDev* R = P.reports + hq;
        R->status = Q->status;
        R->active_trees = active;
        R->iterations = Q->last_iterations;
        R->expansions = Q->last_expansions;
        R->tree_vertices = tv;
        R->words_consumed = (int64_t)Q->word_cursor;
        R->connected = Q->connected;
        R->dm_updated = Q->dm_updated;
        R->near_ties = Q->near_ties;
        R->unresolved_ties = Q->unresolved_ties;
        R->log_entries = (int64_t)Q->log_count;
        R->certified_ties = Q->certified_ties;
        R->tree_hash = Q->tree_hash;
        R->busy_cycles = Q->busy_cycles;
    }
    blk.sync();
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(hot);
        uint32_t* dst = reinterpret_cast<uint32_t*>(Qg);
        for (uint32_t i = blk.tid(); i < IMOMD_QUERY_HOT_BYTES / 4; i += blk.nt()) dst[i] = src[i];
    }
    blk.sync();
}

// a streaming planner's query in the batched mode (cf. run_streaming_query)
template <int KW, class Blk, class Grp>
IMOMD_HD inline void run_streaming_query_batched(Blk& blk, Grp& grp, const PlannerDev& P, int slot, int hq, int steps,
                                                 int B, Shared* SG, BatchShared& BS, double* lb_all, double* mat,
                                                 unsigned char* hot)
{
    const long long t0 = blk.clock();
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(P.query_init + hq);
        uint32_t* dst = reinterpret_cast<uint32_t*>(P.query + hq);
        for (uint32_t i = blk.tid(); i < sizeof(QueryDev) / 4; i += blk.nt()) dst[i] = src[i];
    }
    reset_slot(blk, P, slot);
    fill_bounds(blk, P, slot, hq);
    blk.sync();
    if (blk.tid() == 0)
    {
        QueryDev* Q = P.query + hq;
        for (int k = 0; k < P.K; ++k) Q->free_top[k] = (uint32_t)P.chunks;
        construct_trees(P, slot, hq);
    }
    blk.sync();
    run_query_batched<KW>(blk, grp, P, slot, hq, steps, B, SG, BS, lb_all, mat, hot);
    blk.sync();
    uint64_t h = 0, sum = 0;
    tree_hash(blk, P, slot, SG[0], P.arena + (size_t)slot * P.arena_entries, P.arena_entries, &h, &sum);
    if (blk.tid() == 0)
    {
        QueryDev* Q = P.query + hq;
        Q->tree_hash = h;
        Q->tree_sum = sum;
        Q->busy_cycles = (uint64_t)(blk.clock() - t0);
        P.reports[hq].tree_hash = h;
        P.reports[hq].busy_cycles = Q->busy_cycles;
    }
    blk.sync();
}

}  // namespace imomd

#endif  // IMOMD_CORE_CUH
