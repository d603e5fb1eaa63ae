// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles the device logic of imomd-rrtstar_b200/csrc/imomd_core.cuh for the host
// with a one-thread `Cta` policy, so that the state machine, the frontier hash, the
// pruned searches and the rewire stack can be checked against the oracle on the CPU
// box (no GPU there).  It is NOT a fallback: the shipped library
// (libimomd_b200.so) contains no host execution path and fails without a device.
// Only tests/ builds and loads this file.

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../imomd-rrtstar_b200/csrc/imomd_core.cuh"
#include "../../imomd-rrtstar_b200/csrc/imomd_setup.h"

using namespace imomd;

namespace {

struct CtaHost
{
    int tid() const { return 0; }
    int nt() const { return 1; }
    void sync() const {}
    long long clock() const { return 0; }
    int warp() const { return 0; }
    int lane() const { return 0; }
    int warp_width() const { return 1; }
    void warp_sync() const {}
    uint32_t warp_bcast(uint32_t v) const { return v; }
    bool warp_all(bool p) const { return p; }
    uint32_t warp_scan_small(uint32_t v, int, uint32_t* total) const
    {
        *total = v;
        return 0;
    }
    uint32_t warp_scan_exclusive(uint32_t v, uint32_t* total) const
    {
        *total = v;
        return 0;
    }
    uint32_t counter_add(uint32_t* p, uint32_t v) const
    {
        uint32_t old = *p;
        *p += v;
        return old;
    }
    Best reduce(Best b, Best*) const { return b; }
    double reduce_min_f64(double v, Best*) const { return v; }
    uint32_t scan_exclusive(uint32_t v, uint32_t* total, uint32_t*) const
    {
        *total = v;
        return 0;
    }
    uint32_t reduce_min_u32(uint32_t v, uint32_t*) const { return v; }
    void counter_or(uint32_t* p, uint32_t v) const { *p |= v; }
    int group() const { return 0; }       // batched mode: one group walks the trees one by one
    uint32_t bcast(uint32_t v) const { return v; }
    U4 cas_head(U4* addr, const U4& expect, const U4& want) const
    {
        U4 seen = *addr;
        if (seen.x == expect.x && seen.y == expect.y && seen.z == expect.z && seen.w == expect.w) *addr = want;
        return seen;
    }
    U4 load_cg(const U4* p) const { return *p; }
    void atomic_min_u64(unsigned long long* p, unsigned long long v) const { if (v < *p) *p = v; }
    void atomic_min_u32(uint32_t* p, uint32_t v) const { if (v < *p) *p = v; }
    void fence() const {}
    void prefetch_l2(const void*) const {}
    int n_groups() const { return 1; }
    mutable uint64_t fake_ns = 0;
    uint64_t now_ns() const { return fake_ns += 1000; }   // every reading is 1 us later
    void store_streaming(U4* p, const U4& v) const { *p = v; }
    U4 load_streaming(const U4* p) const { return *p; }
    uint64_t reduce_sum_u64(uint64_t v, Best*) const { return v; }
    // the exact-value service: here the caller already IS the host (glibc)
    mutable uint64_t certified = 0;
    bool enabled = true;
    bool certify(int n, double (*in)[4], double* out) const
    {
        if (!enabled) return false;
        for (int i = 0; i < n; ++i) out[i] = haversine(in[i][0], in[i][1], in[i][2], in[i][3]);
        ++certified;
        return true;
    }
};

struct HsPlanner
{
    PlannerDev d;
    std::vector<uint32_t> row_ptr, col, pcol;
    std::vector<double> pw, coslat;
    std::vector<uint8_t> pdeg;
    std::vector<double> w, lat, lon, cosrow;
    std::vector<VRec> vrec;
    std::vector<QueryDev> query;
    std::vector<unsigned char> trec;
    std::vector<uint32_t> cell_head, cell_cnt, chunk_next, free_stack, arena,
        clist, log, words;
    std::vector<double> lbd, ubd;
    std::vector<uint32_t> occ, visit_log, attach_list;
    std::vector<VRec> rec;
    std::vector<ReportDev> reports;
    std::vector<double> lb;
    std::vector<uint32_t> lvl;
    std::vector<double> mat;
    std::vector<unsigned char> hot;
    std::vector<QueryDev> query_init;
    std::vector<BEvent> events;
    std::vector<uint32_t> clist_tree;
    BatchShared BS;
    Shared S;
    CtaHost cta;
};

}  // namespace

extern "C" {

void* hs_create(uint32_t n, const uint32_t* row_ptr, const uint32_t* col, const double* w,
                const double* lat, const double* lon, int Q, int K, const uint32_t* dest,
                const double* prob, double goal_bias, int pseudo, int grid_dim, int slots, double tie_band,
                char* err, int errlen)
{
    uint32_t md = 0;
    std::string verr;
    if (!validate_graph(n, row_ptr, col, &md, &verr))
    {
        std::strncpy(err, verr.c_str(), errlen - 1);
        return nullptr;
    }
    HsPlanner* p = new HsPlanner;
    uint32_t arcs = row_ptr[n];
    p->row_ptr.assign(row_ptr, row_ptr + n + 1);
    p->col.assign(col, col + arcs);
    p->w.assign(w, w + arcs);
    p->lat.assign(lat, lat + n);
    p->lon.assign(lon, lon + n);
    PlannerDev& d = p->d;
    std::memset(&d, 0, sizeof(d));
    d.g.n = n;
    d.g.arcs = arcs;
    setup_grid(d.g, n, lat, lon, grid_dim, md, &p->cosrow);
    p->vrec.resize(n);
    for (uint32_t v = 0; v < n; ++v)
        p->vrec[v] = make_vrec(v, lat[v], lon[v], d.g.G, d.g.lat0, d.g.lon0, d.g.hlat, d.g.hlon);
    p->coslat.resize(n);
    for (uint32_t v = 0; v < n; ++v) p->coslat[v] = std::cos(lat[v] * kDeg2Rad);
    d.g.coslat = p->coslat.data();
    build_padded_rows(n, row_ptr, col, w, d.g.width, &p->pcol, &p->pw, &p->pdeg);
    d.g.deg = p->pdeg.data();
    d.g.col = p->pcol.data();
    d.g.w = p->pw.data();
    d.g.lat = p->lat.data();
    d.g.lon = p->lon.data();
    d.g.vrec = p->vrec.data();
    d.g.cosrow = p->cosrow.data();
    d.Q = Q;
    d.K = K;
    d.goal_bias = goal_bias;
    d.tie_rel = tie_band > 0.0 ? tie_band : kTieRel;
    d.pseudo = pseudo;
    default_sizes(d, 0, 0, 0);
    if (slots <= 0 || slots > Q) slots = Q;
    d.slots = slots;
    d.streaming = slots < Q ? 1 : 0;
    // heavy state is per SLOT (resident planner: slot == query)
    size_t qk = (size_t)slots * K, cells = (size_t)d.g.G * d.g.G, chunks = d.chunks;
    p->query.resize(Q);
    for (int q = 0; q < Q; ++q)
        fill_initial_header(p->query[q], K, dest + (size_t)q * K, prob + (size_t)q * K * K);
    d.trec_stride = 16u + 4u * (uint32_t)d.g.width;
    p->trec.assign(qk * (size_t)n * d.trec_stride + 16, 0xFF);
    p->cell_head.assign(qk * cells, IMOMD_NONE);
    p->cell_cnt.assign(qk * cells, 0);
    p->rec.resize(qk * chunks * IMOMD_CHUNK);
    p->chunk_next.assign(qk * chunks, IMOMD_NONE);
    p->free_stack.resize(qk * chunks);
    p->lbd.resize(qk * cells);
    p->ubd.resize(qk * cells);
    d.dirty_words = d.streaming ? (uint32_t)((n + 31) / 32) : 0u;
    p->occ.assign(qk * (2 * cells + d.dirty_words), 0);
    p->arena.resize((size_t)slots * d.arena_entries);
    p->clist.resize((size_t)slots * 2 * chunks);
    p->log.resize((size_t)slots * 2 * d.log_entries);
    p->reports.resize(Q);
    p->lb.resize(std::max<size_t>(cells, 1024));   // (also holds the batched mode's connection candidates)
    p->lvl.resize(2 * kLevelCap);
    p->mat.resize((size_t)K * K * 4);
    p->hot.resize(IMOMD_QUERY_HOT_BYTES + 64 + (size_t)K * sizeof(TreeRef) + (size_t)d.g.width * K * sizeof(double));
    d.query = p->query.data();
    d.trec = reinterpret_cast<unsigned char*>(((uintptr_t)p->trec.data() + 15) & ~(uintptr_t)15);
    d.cell_head = p->cell_head.data();
    d.cell_cnt = p->cell_cnt.data();
    d.rec = p->rec.data();
    d.chunk_next = p->chunk_next.data();
    d.free_stack = p->free_stack.data();
    d.lbd = p->lbd.data();
    d.ubd = p->ubd.data();
    d.occ = p->occ.data();
    if (pseudo)
    {
        p->visit_log.assign(qk * (size_t)n, IMOMD_NONE);
        d.visit_log = p->visit_log.data();
        p->attach_list.assign(n, 0);
        d.attach_list = p->attach_list.data();
    }
    d.arena = p->arena.data();
    d.clist = p->clist.data();
    d.log = p->log.data();
    d.reports = p->reports.data();
    p->query_init = p->query;
    d.query_init = p->query_init.data();
    d.ev_cap = 4096;
    p->events.resize(qk * d.ev_cap);
    d.events = p->events.data();
    p->clist_tree.resize(qk * 2 * chunks);
    d.clist_tree = p->clist_tree.data();
    if (d.streaming) return p;     // slots are rebuilt when a query takes them (run_streaming_query)
    for (size_t i = 0; i < qk; ++i)
        for (size_t c = 0; c < chunks; ++c) p->free_stack[i * chunks + c] = (uint32_t)(chunks - 1 - c);
    for (int q = 0; q < Q; ++q)
    {
        for (int k = 0; k < K; ++k) p->query[q].free_top[k] = (uint32_t)chunks;
        construct_trees(d, q, q);
    }
    for (size_t i = 0; i < qk * cells; ++i)
    {
        p->lbd[i] = lbd_entry(d, i);
        p->ubd[i] = ubd_entry(d, i);
    }
    return p;
}

void hs_destroy(void* hp) { delete (HsPlanner*)hp; }

int hs_grid_dim(void* hp) { return ((HsPlanner*)hp)->d.g.G; }

void hs_feed_words(void* hp, const uint32_t* words, size_t n)
{
    HsPlanner* p = (HsPlanner*)hp;
    p->words.insert(p->words.end(), words, words + n);
    p->d.words = p->words.data();
    p->d.words_avail = p->words.size();
}

void hs_expand(void* hp, int iters, ReportDev* reports)
{
    HsPlanner* p = (HsPlanner*)hp;
    CtaHost& cta = p->cta;
    unsigned char* hot = reinterpret_cast<unsigned char*>(((uintptr_t)p->hot.data() + 15) & ~(uintptr_t)15);
    if (p->d.streaming)
    {
        // the device hands queries to whichever slot falls free; queries are independent, so any
        // assignment gives the same per-query results -- here query hq reuses slot hq % slots
        for (int hq = 0; hq < p->d.Q; ++hq)
        {
            const int slot = hq % p->d.slots;
            if (p->d.g.width == 4)
                run_streaming_query<4>(cta, p->d, slot, hq, iters, p->S, p->lb.data(), p->lvl.data(), p->mat.data(), hot);
            else if (p->d.g.width == 8)
                run_streaming_query<8>(cta, p->d, slot, hq, iters, p->S, p->lb.data(), p->lvl.data(), p->mat.data(), hot);
            else
                run_streaming_query<16>(cta, p->d, slot, hq, iters, p->S, p->lb.data(), p->lvl.data(), p->mat.data(), hot);
        }
        if (reports) std::memcpy(reports, p->reports.data(), p->reports.size() * sizeof(ReportDev));
        return;
    }
    for (int q = 0; q < p->d.Q; ++q)
    {
        if (p->d.g.width == 4)
            run_query<4>(cta, p->d, q, q, iters, p->S, p->lb.data(), p->lvl.data(), p->mat.data(), hot);
        else if (p->d.g.width == 8)
            run_query<8>(cta, p->d, q, q, iters, p->S, p->lb.data(), p->lvl.data(), p->mat.data(), hot);
        else
            run_query<16>(cta, p->d, q, q, iters, p->S, p->lb.data(), p->lvl.data(), p->mat.data(), hot);
    }
    if (reports) std::memcpy(reports, p->reports.data(), p->reports.size() * sizeof(ReportDev));
}

// imomd_expand_batch on the host simulation: `steps` steps of B expansions per tree
void hs_expand_batch(void* hp, int steps, int B, ReportDev* reports)
{
    HsPlanner* p = (HsPlanner*)hp;
    CtaHost& cta = p->cta;
    unsigned char* hot = reinterpret_cast<unsigned char*>(((uintptr_t)p->hot.data() + 15) & ~(uintptr_t)15);
    for (int hq = 0; hq < p->d.Q; ++hq)
    {
        const int slot = p->d.streaming ? hq % p->d.slots : hq;
#define HS_BATCH(W)                                                                                              \
    if (p->d.streaming)                                                                                          \
        run_streaming_query_batched<W>(cta, cta, p->d, slot, hq, steps, B, &p->S, p->BS, p->lb.data(), p->mat.data(), hot);    \
    else                                                                                                         \
        run_query_batched<W>(cta, cta, p->d, slot, hq, steps, B, &p->S, p->BS, p->lb.data(), p->mat.data(), hot);
        if (p->d.g.width == 4) { HS_BATCH(4) }
        else if (p->d.g.width == 8) { HS_BATCH(8) }
        else { HS_BATCH(16) }
#undef HS_BATCH
    }
    if (reports) std::memcpy(reports, p->reports.data(), p->reports.size() * sizeof(ReportDev));
}

// imomd_attach_list on the host simulation (pseudo-mode planners): the merge's per-vertex work
void hs_attach_list(void* hp, int q, int k, const uint32_t* vertices, uint32_t count, uint32_t* parents_out)
{
    HsPlanner* p = (HsPlanner*)hp;
    if (!p->d.attach_list || count == 0) return;
    std::memcpy(p->d.attach_list, vertices, (size_t)count * 4);
    CtaHost& cta = p->cta;
    unsigned char* hot = reinterpret_cast<unsigned char*>(((uintptr_t)p->hot.data() + 15) & ~(uintptr_t)15);
    if (p->d.g.width == 4)
        run_query<4>(cta, p->d, q, q, 0, p->S, p->lb.data(), p->lvl.data(), p->mat.data(), hot, k, count);
    else if (p->d.g.width == 8)
        run_query<8>(cta, p->d, q, q, 0, p->S, p->lb.data(), p->lvl.data(), p->mat.data(), hot, k, count);
    else
        run_query<16>(cta, p->d, q, q, 0, p->S, p->lb.data(), p->lvl.data(), p->mat.data(), hot, k, count);
    if (parents_out) std::memcpy(parents_out, p->d.attach_list, (size_t)count * 4);
}

uint32_t hs_visit_order(void* hp, int q, int k, uint32_t* out)
{
    HsPlanner* p = (HsPlanner*)hp;
    if (!p->d.visit_log) return 0;
    uint32_t n = (uint32_t)p->query[q].tree_size[k];
    std::memcpy(out, p->d.visit_log + ((size_t)q * p->d.K + k) * p->d.g.n, (size_t)n * 4);
    return n;
}

void hs_get_tree(void* hp, int q, int k, uint32_t* parent, double* cost)
{
    HsPlanner* p = (HsPlanner*)hp;
    size_t n = p->d.g.n;
    TreeRef T = tree_ref(p->d, q, k);
    for (size_t v = 0; v < n; ++v)
    {
        parent[v] = T.parent((uint32_t)v);
        cost[v] = parent[v] == IMOMD_NONE ? INFINITY : T.cost((uint32_t)v);
    }
}

void hs_get_state(void* hp, int q, double* D, uint32_t* cn, double* prob, uint32_t* mhi,
                  double* mhv, int32_t* is_done, int32_t* ds_parent, int32_t* flags)
{
    HsPlanner* p = (HsPlanner*)hp;
    const QueryDev& h = p->query[q];
    int K = p->d.K;
    for (int i = 0; i < K * K; ++i)
    {
        D[i] = h.D[i];
        cn[i] = h.CN[i];
        prob[i] = h.P[i];
        mhi[i] = h.MHi[i];
        mhv[i] = h.MHv[i];
    }
    for (int k = 0; k < K; ++k)
    {
        is_done[k] = h.is_done[k];
        ds_parent[k] = h.ds_parent[k];
    }
    flags[0] = h.connected;
    flags[1] = h.dm_updated;
}

void hs_tie_info(void* hp, int q, double* out)
{
    HsPlanner* p = (HsPlanner*)hp;
    for (int i = 0; i < 8; ++i) out[i] = p->query[q].tie_info[i];
}

uint32_t hs_get_frontier(void* hp, int q, int k, uint32_t* out)
{
    HsPlanner* p = (HsPlanner*)hp;
    size_t n = p->d.g.n;
    TreeRef T = tree_ref(p->d, q, k);
    uint32_t c = 0;
    for (size_t v = 0; v < n; ++v)
        if (T.fslot((uint32_t)v) != IMOMD_NONE) out[c++] = (uint32_t)v;
    return c;
}

// consistency of the frontier hash: every listed record sits in the right cell, slots
// agree with fslot, counts agree; returns 0 when consistent
int hs_check_hash(void* hp, int q, int k)
{
    HsPlanner* p = (HsPlanner*)hp;
    TreeRef T = tree_ref(p->d, q, k);
    size_t cells = (size_t)p->d.g.G * p->d.g.G;
    uint32_t total = 0;
    for (uint32_t c = 0; c < cells; ++c)
    {
        uint32_t cnt = T.cell_cnt[c];
        if (!cnt) continue;
        uint32_t chunk = T.cell_head[c];
        uint32_t nrec = (cnt - 1) % IMOMD_CHUNK + 1, left = cnt;
        while (left)
        {
            if (chunk == IMOMD_NONE) return 1;
            for (uint32_t i = 0; i < nrec; ++i)
            {
                const VRec& r = T.rec[(size_t)chunk * IMOMD_CHUNK + i];
                if (r.cell != c) return 2;
                if (T.fslot(r.id) != chunk * IMOMD_CHUNK + i) return 3;
                if (T.parent(r.id) != IMOMD_NONE) return 4;
            }
            left -= nrec;
            total += nrec;
            nrec = IMOMD_CHUNK;
            chunk = T.chunk_next[chunk];
        }
    }
    if (total != p->query[q].frontier_count[k]) return 5;
    return 0;
}

void hs_nearest(void* hp, int q, int k, double lat, double lon, uint32_t* idx, double* dist,
                int64_t* stats)
{
    HsPlanner* p = (HsPlanner*)hp;
    CtaHost& cta = p->cta;
    p->S.Q = &p->query[q];
    p->S.tref = nullptr;
    p->S.qlat = lat;
    p->S.qlon = lon;
    svc_nearest(cta, p->d, q, k, p->S, p->lb.data(), p->clist.data() + (size_t)q * 2 * p->d.chunks,
                &stats[0], &stats[1], false);
    uint32_t v = p->S.result.id1;
    *idx = v;
    *dist = v == IMOMD_NONE ? INFINITY : haversine(p->lat[v], p->lon[v], lat, lon);
    stats[2] = p->S.n_list;   // chunks examined by the pruned search
}

// TREEHASH / checksum of a resident query's trees through the device routine
void hs_treehash(void* hp, int q, uint64_t* hash, uint64_t* sum)
{
    HsPlanner* p = (HsPlanner*)hp;
    if (p->d.streaming)
    {
        *hash = p->query[q].tree_hash;
        *sum = p->query[q].tree_sum;
        return;
    }
    tree_hash(p->cta, p->d, q, p->S, p->d.arena + (size_t)q * p->d.arena_entries, p->d.arena_entries, hash, sum);
}

// exact-value service: requests answered so far; enable = 0 detaches it (device decisions kept)
uint64_t hs_certified(void* hp) { return ((HsPlanner*)hp)->cta.certified; }
void hs_set_certify(void* hp, int enable) { ((HsPlanner*)hp)->cta.enabled = enable != 0; }

uint32_t hs_connection_log(void* hp, int q, uint32_t* set_idx, uint32_t* vertex, uint32_t cap)
{
    HsPlanner* p = (HsPlanner*)hp;
    uint32_t c = p->query[q].log_count;
    const uint32_t* lg = p->log.data() + (size_t)q * 2 * p->d.log_entries;
    for (uint32_t i = 0; i < c && i < cap && i < p->d.log_entries; ++i)
    {
        set_idx[i] = lg[2 * i];
        vertex[i] = lg[2 * i + 1];
    }
    return c;
}

}  // extern "C"
