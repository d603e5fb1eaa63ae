// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin C API over the *unmodified* reference implementation.  This file is
// compiled together with the reference's own translation units, taken where
// they lie under /root/reference (never copied into this repository), into
// oracle/_ref/libimomd_ref.so and oracle/_ref/ref_probe by oracle/Makefile.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
// arm may load it.  Nothing under imomd-rrtstar_b200/ links or calls it.
//
// Private members of ImomdRRT / EciGenSolver are reached with the usual
// `#define private public` trick so that the expansion
// loop can be driven and inspected without touching the reference sources.
//
// Reference entry points exercised here (path:line under /root/reference):
//   ImomdRRT::ImomdRRT                 src/imomd_rrt_star.cpp:39-223
//   ImomdRRT::expandTreeLayers_        src/imomd_rrt_star.cpp:329-335
//   ImomdRRT::expandTree_ (step by step for traces)   :337-356
//   ImomdRRT::solveRTSP_               src/imomd_rrt_star.cpp:966-1003
//   EciGenSolver::solveRTSP            src/eci_gen_tsp_solver.cpp:76-101
//   EciGenSolver::calculatePathCost_   src/eci_gen_tsp_solver.cpp:442-450
//   computeHaversineDistance           include/osm_converter/compute_haversine.hpp:41-58
//   OSMParser::parse                   src/osm_parser.cpp:48-76

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <list>
#include <memory>
#include <numeric>
#include <queue>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <unistd.h>

#define private public
#include "imomd_rrt_star/imomd_rrt_star.h"
#undef private
#include "osm_converter/osm_parser.h"
#include "fake_map/fake_map.h"

// The reference library declares this `extern` (include/utils/debugger.h:44);
// a message prints iff level >= DEBUG_LEVEL, so 11 silences everything.
int DEBUG_LEVEL = 11;

namespace {

struct RefGraph
{
    std::shared_ptr<std::vector<location_t>> map;
    std::shared_ptr<std::vector<std::unordered_map<size_t, double>>> conn;
};

struct RefPlanner
{
    RefGraph* g;
    ImomdRRT* rrt;   // never destroyed: ~CSVFile flushes an unopened stream with
                     // exceptions enabled and aborts
    int K;
};

const uint32_t NONE32 = 0xFFFFFFFFu;

}  // namespace

extern "C" {

// ---------------------------------------------------------------- graphs ---

void* ref_graph_from_edges(uint64_t n, const double* lat, const double* lon,
                           uint64_t m, const uint64_t* eu, const uint64_t* ev,
                           const double* ew)
{
    RefGraph* g = new RefGraph;
    g->map = std::make_shared<std::vector<location_t>>(n);
    g->conn = std::make_shared<std::vector<std::unordered_map<size_t, double>>>(n);
    for (uint64_t i = 0; i < n; ++i)
    {
        (*g->map)[i] = location_t(i, lat[i], lon[i]);
    }
    for (uint64_t e = 0; e < m; ++e)
    {
        // same two inserts as OSMParser::createNetwork_ (src/osm_parser.cpp:179-180)
        double w = ew ? ew[e]
                      : computeHaversineDistance((*g->map)[eu[e]], (*g->map)[ev[e]]);
        (*g->conn)[eu[e]].insert({ev[e], w});
        (*g->conn)[ev[e]].insert({eu[e], w});
    }
    return g;
}

// Jittered W x W lattice exactly as in SURVEY.md Appendix A step 2.
void* ref_graph_grid(int W)
{
    RefGraph* g = new RefGraph;
    size_t n = (size_t)W * W;
    g->map = std::make_shared<std::vector<location_t>>(n);
    g->conn = std::make_shared<std::vector<std::unordered_map<size_t, double>>>(n);
    std::mt19937 jr(123);
    std::uniform_real_distribution<> jit(-2e-5, 2e-5);
    for (int r = 0; r < W; ++r)
    {
        for (int c = 0; c < W; ++c)
        {
            size_t i = (size_t)r * W + c;
            // g++ evaluates the constructor arguments of the survey's probe right to
            // left, so the longitude jitter is drawn first.
            double jlon = jit(jr);
            double jlat = jit(jr);
            (*g->map)[i] = location_t(i, 37.0 + r * 1e-4 + jlat, -122.0 + c * 1e-4 + jlon);
        }
    }
    for (int r = 0; r < W; ++r)
    {
        for (int c = 0; c < W; ++c)
        {
            size_t i = (size_t)r * W + c;
            if (c + 1 < W)
            {
                double w = computeHaversineDistance((*g->map)[i], (*g->map)[i + 1]);
                (*g->conn)[i].insert({i + 1, w});
                (*g->conn)[i + 1].insert({i, w});
            }
            if (r + 1 < W)
            {
                double w = computeHaversineDistance((*g->map)[i], (*g->map)[i + W]);
                (*g->conn)[i].insert({i + W, w});
                (*g->conn)[i + W].insert({i, w});
            }
        }
    }
    return g;
}

// OSM file parsed with the `osm_all` key filter (config/osm_way_config.yaml:5-10,
// main.cpp:114-117).
void* ref_graph_from_osm(const char* path)
{
    map_properties_t mp;
    mp.key = {"highway", "footway", "sidewalk", "cycleway"};
    OSMParser parser(path, mp);
    try
    {
        parser.parse();
    }
    catch (const std::exception& ex)
    {
        std::fprintf(stderr, "ref_graph_from_osm: %s\n", ex.what());
        return nullptr;
    }
    RefGraph* g = new RefGraph;
    g->map = parser.getMap();
    g->conn = parser.getConnection();
    return g;
}

void* ref_graph_fake(int type)
{
    FakeMap fm(type);
    RefGraph* g = new RefGraph;
    g->map = fm.getMap();
    g->conn = fm.getGraph();
    return g;
}

uint64_t ref_graph_num_nodes(void* gp) { return ((RefGraph*)gp)->map->size(); }

uint64_t ref_graph_num_arcs(void* gp)
{
    uint64_t s = 0;
    for (const auto& m : *((RefGraph*)gp)->conn) s += m.size();
    return s;
}

// CSR in the containers' *native iteration order* (SURVEY.md section 8a:
// "CSR order = native adjacency iteration order").
void ref_graph_export(void* gp, double* lat, double* lon, uint64_t* ids,
                      uint32_t* row_ptr, uint32_t* col, double* w)
{
    RefGraph* g = (RefGraph*)gp;
    size_t n = g->map->size();
    uint32_t pos = 0;
    for (size_t i = 0; i < n; ++i)
    {
        lat[i] = (*g->map)[i].latitude;
        lon[i] = (*g->map)[i].longitude;
        if (ids) ids[i] = (*g->map)[i].id;
        row_ptr[i] = pos;
        for (const auto& kv : (*g->conn)[i])
        {
            col[pos] = (uint32_t)kv.first;
            w[pos] = kv.second;
            ++pos;
        }
    }
    row_ptr[n] = pos;
}

double ref_haversine(double lat1, double lon1, double lat2, double lon2)
{
    location_t a(lat1, lon1), b(lat2, lon2);
    return computeHaversineDistance(a, b);
}

double ref_bearing(double lat1, double lon1, double lat2, double lon2)
{
    location_t a(lat1, lon1), b(lat2, lon2);
    return computeBearing(a, b);
}

// Destination picking of SURVEY.md Appendix A step 3.
void ref_pick_destinations(void* gp, uint32_t seed, int count, uint64_t* out)
{
    RefGraph* g = (RefGraph*)gp;
    std::mt19937 pick(seed);
    std::uniform_int_distribution<size_t> u(0, g->map->size() - 1);
    int got = 0;
    while (got < count)
    {
        size_t v = u(pick);
        if (!(*g->conn)[v].empty()) out[got++] = v;
    }
}

// Connected-component label per vertex (BFS), returns the number of components.
uint64_t ref_graph_components(void* gp, uint32_t* label)
{
    RefGraph* g = (RefGraph*)gp;
    size_t n = g->map->size();
    std::fill(label, label + n, NONE32);
    uint32_t c = 0;
    std::vector<size_t> q;
    for (size_t s = 0; s < n; ++s)
    {
        if (label[s] != NONE32) continue;
        q.clear();
        q.push_back(s);
        label[s] = c;
        for (size_t h = 0; h < q.size(); ++h)
        {
            for (const auto& kv : (*g->conn)[q[h]])
            {
                if (label[kv.first] == NONE32)
                {
                    label[kv.first] = c;
                    q.push_back(kv.first);
                }
            }
        }
        ++c;
    }
    return c;
}

// --------------------------------------------------------------- planner ---

void* ref_planner_create(void* gp, uint64_t source, uint64_t target,
                         const uint64_t* objectives, int n_obj, double goal_bias,
                         int pseudo_mode, int max_iter, int max_time, int swapping,
                         int genetic, int ga_mutation_iter, int ga_population,
                         int ga_generation)
{
    RefGraph* g = (RefGraph*)gp;
    imomd_setting_t s;
    s.goal_bias = goal_bias;
    s.max_iter = max_iter;
    s.max_time = max_time;
    s.random_seed = false;
    s.pseudo_mode = pseudo_mode != 0;
    s.log_data = 0;
    s.rtsp_setting.swapping = swapping != 0;
    s.rtsp_setting.genetic = genetic != 0;
    s.rtsp_setting.ga_mutation_iter = ga_mutation_iter;
    s.rtsp_setting.ga_population = ga_population;
    s.rtsp_setting.ga_generation = ga_generation;
    s.rtsp_setting.ga_random_seed = false;
    std::vector<size_t> obj(objectives, objectives + n_obj);
    RefPlanner* p = new RefPlanner;
    p->g = g;
    p->rrt = new ImomdRRT(source, target, obj, g->map, g->conn, s);
    p->K = n_obj + 2;
    return p;
}

int ref_planner_K(void* pp) { return ((RefPlanner*)pp)->K; }

// `iters` calls of expandTreeLayers_(); returns the number of expansions
// (non-early-returning expandTree_ calls).
uint64_t ref_planner_iterate(void* pp, int iters)
{
    RefPlanner* p = (RefPlanner*)pp;
    uint64_t expansions = 0;
    for (int it = 0; it < iters; ++it)
    {
        // body of expandTreeLayers_ (:331-334) with the early-return test of
        // expandTree_ (:339) evaluated first so that expansions are counted exactly
        for (tree_t& t : p->rrt->tree_layers_)
        {
            if (!(t.expandables.empty() || t.is_done)) ++expansions;
            p->rrt->expandTree_(t);
        }
    }
    return expansions;
}

struct ref_trace_t
{
    int32_t tree;
    int32_t skipped;
    uint64_t rand_id;
    double rand_lat;
    double rand_lon;
    uint64_t x_nearest;   // arg-min of steerNewNode_ before JPS (recomputed here)
    double nearest_dist;
    uint64_t x_new;       // after JPS
    uint64_t parent_of_new;
    double cost_of_new;
    uint64_t tree_size_after;
    uint64_t frontier_after;
};

// One expandTree_ (src/imomd_rrt_star.cpp:337-356) executed step by step so the
// intermediate values can be recorded.  The sequence of calls is the body of
// the reference function; only the recording is added.
int ref_planner_expand_traced(void* pp, int k, ref_trace_t* tr)
{
    RefPlanner* p = (RefPlanner*)pp;
    ImomdRRT& r = *p->rrt;
    tree_t& tree = r.tree_layers_[k];
    std::memset(tr, 0, sizeof(*tr));
    tr->tree = k;
    if (tree.expandables.empty() || tree.is_done)
    {
        tr->skipped = 1;
        return 0;
    }
    location_t random_point = r.selectRandomVertex_(tree);
    tr->rand_id = random_point.id;
    tr->rand_lat = random_point.latitude;
    tr->rand_lon = random_point.longitude;
    {
        double best = INFINITY;
        size_t arg = (size_t)-1;
        for (size_t v : tree.expandables)
        {
            double d = computeHaversineDistance((*r.map_ptr_)[v], random_point);
            if (d < best)
            {
                best = d;
                arg = v;
            }
        }
        tr->x_nearest = arg;
        tr->nearest_dist = best;
    }
    size_t x_new = r.steerNewNode_(tree, random_point);
    r.connectNewNode_(tree, x_new);
    r.updateExpandables(tree, x_new, true);
    r.rewireTree_(tree, x_new);
    r.updateSelectionProbability(tree, random_point);
    r.updateConnectionTree_(tree, x_new);
    tr->x_new = x_new;
    tr->parent_of_new = tree.parent.count(x_new) ? tree.parent[x_new] : (uint64_t)-1;
    tr->cost_of_new = tree.cost.count(x_new) ? tree.cost[x_new] : NAN;
    tr->tree_size_after = tree.parent.size();
    tr->frontier_after = tree.expandables.size();
    return 1;
}

void ref_planner_get_tree(void* pp, int k, uint32_t* parent, double* cost)
{
    RefPlanner* p = (RefPlanner*)pp;
    size_t n = p->g->map->size();
    const tree_t& t = p->rrt->tree_layers_[k];
    for (size_t i = 0; i < n; ++i)
    {
        parent[i] = NONE32;
        cost[i] = INFINITY;
    }
    for (const auto& kv : t.parent) parent[kv.first] = (uint32_t)kv.second;
    for (const auto& kv : t.cost) cost[kv.first] = kv.second;
}

uint64_t ref_planner_tree_size(void* pp, int k)
{
    return ((RefPlanner*)pp)->rrt->tree_layers_[k].parent.size();
}

// children of `v` in tree k in the container's iteration order; returns count.
int ref_planner_children(void* pp, int k, uint64_t v, uint64_t* out, int cap)
{
    tree_t& t = ((RefPlanner*)pp)->rrt->tree_layers_[k];
    auto it = t.children.find(v);
    if (it == t.children.end()) return -1;
    int c = 0;
    for (size_t ch : it->second)
    {
        if (c < cap) out[c] = ch;
        ++c;
    }
    return c;
}

// FNV-style hash of SURVEY.md Appendix A step 6.
uint64_t ref_planner_treehash(void* pp)
{
    RefPlanner* p = (RefPlanner*)pp;
    uint64_t h = 1469598103934665603ull;
    for (const tree_t& t : p->rrt->tree_layers_)
    {
        std::vector<size_t> keys;
        keys.reserve(t.parent.size());
        for (const auto& kv : t.parent) keys.push_back(kv.first);
        std::sort(keys.begin(), keys.end());
        for (size_t v : keys)
        {
            uint64_t par = t.parent.at(v);
            double c = t.cost.at(v);
            uint64_t bits;
            std::memcpy(&bits, &c, 8);
            h = (h ^ (uint64_t)v) * 1099511628211ull;
            h = (h ^ par) * 1099511628211ull;
            h = (h ^ bits) * 1099511628211ull;
        }
    }
    return h;
}

void ref_planner_get_matrices(void* pp, double* D, uint64_t* conn_node, double* prob,
                              uint64_t* mh_id, double* mh_val)
{
    RefPlanner* p = (RefPlanner*)pp;
    ImomdRRT& r = *p->rrt;
    int K = p->K;
    for (int i = 0; i < K; ++i)
    {
        for (int j = 0; j < K; ++j)
        {
            if (D) D[i * K + j] = r.distance_matrix_[i][j];
            if (conn_node) conn_node[i * K + j] = r.connection_node_matrix_[i][j];
            if (prob) prob[i * K + j] = r.probability_matrix_[i][j];
            if (mh_id) mh_id[i * K + j] = r.expandables_min_heuristic_matrix_[i][j].first;
            if (mh_val) mh_val[i * K + j] = r.expandables_min_heuristic_matrix_[i][j].second;
        }
    }
}

void ref_planner_get_flags(void* pp, int32_t* is_done, int32_t* connected,
                           int32_t* dm_updated, int32_t* ds_parent)
{
    RefPlanner* p = (RefPlanner*)pp;
    for (int k = 0; k < p->K; ++k)
    {
        if (is_done) is_done[k] = p->rrt->tree_layers_[k].is_done ? 1 : 0;
        if (ds_parent) ds_parent[k] = p->rrt->disjoint_set_parent_[k];
    }
    if (connected) *connected = p->rrt->is_connected_graph_ ? 1 : 0;
    if (dm_updated) *dm_updated = p->rrt->is_distance_matrix_updated_ ? 1 : 0;
}

uint64_t ref_planner_frontier(void* pp, int k, uint64_t* out, uint64_t cap)
{
    const tree_t& t = ((RefPlanner*)pp)->rrt->tree_layers_[k];
    uint64_t c = 0;
    for (size_t v : t.expandables)
    {
        if (c < cap) out[c] = v;
        ++c;
    }
    return c;
}

// Pseudo-mode connection-node sets (lower-triangular index i*(i-1)/2+j).
uint64_t ref_planner_connection_set(void* pp, int idx, uint64_t* out, uint64_t cap)
{
    RefPlanner* p = (RefPlanner*)pp;
    uint64_t c = 0;
    for (size_t v : *p->rrt->connection_nodes_set_[idx])
    {
        if (c < cap) out[c] = v;
        ++c;
    }
    return c;
}

// The arg-min loop of steerNewNode_ (src/imomd_rrt_star.cpp:406-418) on the
// tree's real `expandables` container, without the JPS side effects.
// n_ties = number of frontier vertices whose distance equals the minimum.
int ref_planner_nearest(void* pp, int k, double lat, double lon, uint64_t* idx,
                        double* dist, int32_t* n_ties, double* second)
{
    RefPlanner* p = (RefPlanner*)pp;
    const tree_t& t = p->rrt->tree_layers_[k];
    location_t q(lat, lon);
    double best = INFINITY, sec = INFINITY;
    size_t arg = (size_t)-1;
    int ties = 0;
    for (size_t v : t.expandables)
    {
        double d = computeHaversineDistance((*p->rrt->map_ptr_)[v], q);
        if (d < best)
        {
            sec = best;
            best = d;
            arg = v;
            ties = 1;
        }
        else if (d == best)
        {
            ++ties;
        }
        else if (d < sec)
        {
            sec = d;
        }
    }
    *idx = arg;
    *dist = best;
    if (n_ties) *n_ties = ties;
    if (second) *second = sec;
    return t.expandables.empty() ? 0 : 1;
}

// ------------------------------------------------- anytime loop / RTSP -----

struct ref_log_row_t
{
    double time_s;
    double path_cost;
    int64_t tree_size;
    int64_t iteration;
};

// The loop of ImomdRRT::findShortestPath (src/imomd_rrt_star.cpp:225-283) driven
// from here (the original ends in pthread_exit and cannot return).  One row is
// recorded at every moment the reference calls logData_: when solveRTSP_ lowers
// shortest_path_cost_ (:978-984) or, in pseudo mode, on every solveRTSP_ that
// finds the graph connected and the distance matrix updated (:990-1001, the
// cost may then stay the same).  Returns the number of rows.
int ref_planner_run(void* pp, ref_log_row_t* rows, int cap, int64_t* iterations,
                    double* elapsed_s, int64_t* expansions)
{
    RefPlanner* p = (RefPlanner*)pp;
    ImomdRRT& r = *p->rrt;
    int n_rows = 0;
    int64_t n_exp = 0;
    auto record = [&](double before, bool pseudo_log) {
        if ((r.setting_.pseudo_mode ? pseudo_log : r.shortest_path_cost_ < before) && n_rows < cap)
        {
            int64_t ts = 0;
            for (const auto& t : r.tree_layers_) ts += (int64_t)t.parent.size();
            rows[n_rows].time_s = bipedlab::timing::spendElapsedTime(r.start_time_);
            rows[n_rows].path_cost = r.shortest_path_cost_;
            rows[n_rows].tree_size = ts;
            rows[n_rows].iteration = r.iteration_count_;
            ++n_rows;
        }
    };
    bool done_iter = false, done_time = false, done_exp = false;
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());   // logData_ prints unconditionally (:1014)
    while (!done_iter && !done_time && !done_exp)
    {
        for (tree_t& t : r.tree_layers_)   // expandTreeLayers_ (:331-334), counted
        {
            if (!(t.expandables.empty() || t.is_done)) ++n_exp;
            r.expandTree_(t);
        }
        if (r.iteration_count_ % 100 == 0)
        {
            double before = r.shortest_path_cost_;
            bool pseudo_log = r.is_connected_graph_ && r.is_distance_matrix_updated_;
            r.solveRTSP_();
            record(before, pseudo_log);
            done_exp = true;
            for (const tree_t& t : r.tree_layers_)
            {
                if (!t.is_done) done_exp = false;
            }
        }
        if (r.iteration_count_++ > r.setting_.max_iter)
        {
            done_iter = true;
        }
        else if (bipedlab::timing::spendElapsedTime(r.start_time_) > r.setting_.max_time)
        {
            done_time = true;
        }
    }
    for (int c = 0; c < 10; ++c)
    {
        double before = r.shortest_path_cost_;
        bool pseudo_log = r.is_connected_graph_ && r.is_distance_matrix_updated_;
        r.solveRTSP_();
        record(before, pseudo_log);
    }
    std::cout.rdbuf(old);
    if (iterations) *iterations = r.iteration_count_;
    if (elapsed_s) *elapsed_s = bipedlab::timing::spendElapsedTime(r.start_time_);
    if (expansions) *expansions = n_exp;
    return n_rows;
}

// The per-vertex work of mergeTwoTree_ (src/imomd_rrt_star.cpp:848-851 / :883-886) on tree k:
// connectNewNode_, updateExpandables(.., true), rewireTree_, updateConnectionTree_.
void ref_planner_attach(void* pp, int k, uint64_t x)
{
    RefPlanner* p = (RefPlanner*)pp;
    tree_t& t = p->rrt->tree_layers_[k];
    p->rrt->connectNewNode_(t, x);
    p->rrt->updateExpandables(t, x, true);
    p->rrt->rewireTree_(t, x);
    p->rrt->updateConnectionTree_(t, x);
}

// keys of tree k's parent map in iteration order (what mergeTwoTree_ :860 walks)
uint64_t ref_planner_parent_order(void* pp, int k, uint64_t* out, uint64_t cap)
{
    RefPlanner* p = (RefPlanner*)pp;
    uint64_t c = 0;
    for (const auto& kv : p->rrt->tree_layers_[k].parent)
    {
        if (c < cap) out[c] = kv.first;
        ++c;
    }
    return c;
}

// One solveRTSP_ call (src/imomd_rrt_star.cpp:966-1003); returns 1 when the
// path improved.
int ref_planner_solve_rtsp(void* pp)
{
    RefPlanner* p = (RefPlanner*)pp;
    double before = p->rrt->shortest_path_cost_;
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    p->rrt->solveRTSP_();
    std::cout.rdbuf(old);
    return p->rrt->shortest_path_cost_ < before ? 1 : 0;
}

double ref_planner_path_cost(void* pp) { return ((RefPlanner*)pp)->rrt->shortest_path_cost_; }

uint64_t ref_planner_path(void* pp, uint64_t* out, uint64_t cap)
{
    RefPlanner* p = (RefPlanner*)pp;
    uint64_t c = 0;
    for (size_t v : p->rrt->shortest_path_)
    {
        if (c < cap) out[c] = v;
        ++c;
    }
    return c;
}

int ref_planner_sequence(void* pp, int32_t* out, int cap)
{
    RefPlanner* p = (RefPlanner*)pp;
    if (!p->rrt->sequence_of_tree_id_rtsp_) return 0;
    int c = 0;
    for (int v : *p->rrt->sequence_of_tree_id_rtsp_)
    {
        if (c < cap) out[c] = v;
        ++c;
    }
    return c;
}

// ------------------------------------------------------- ECI-Gen solver ----

void* ref_solver_create(int swapping, int genetic, int ga_mutation_iter, int ga_population,
                        int ga_generation)
{
    rtsp_setting_t s;
    s.swapping = swapping != 0;
    s.genetic = genetic != 0;
    s.ga_mutation_iter = ga_mutation_iter;
    s.ga_population = ga_population;
    s.ga_generation = ga_generation;
    s.ga_random_seed = false;
    return new EciGenSolver(s);
}

// EciGenSolver::solveRTSP (src/eci_gen_tsp_solver.cpp:76-101) on a row-major
// K x K matrix.  Returns the sequence length.
int ref_solver_solve(void* sp, int K, const double* D, int source, int target, double* cost,
                     int32_t* seq, int cap)
{
    EciGenSolver* s = (EciGenSolver*)sp;
    auto M = std::make_shared<std::vector<std::vector<double>>>(K, std::vector<double>(K));
    for (int i = 0; i < K; ++i)
    {
        for (int j = 0; j < K; ++j) (*M)[i][j] = D[i * K + j];
    }
    auto res = s->solveRTSP(M, source, target);
    *cost = res.first;
    int c = 0;
    for (int v : *res.second)
    {
        if (c < cap) seq[c] = v;
        ++c;
    }
    return c;
}

// EciGenSolver::calculatePathCost_ (src/eci_gen_tsp_solver.cpp:442-450) for a
// batch of tours stored with a fixed stride.
void ref_path_costs(int K, const double* D, int n_tours, int stride, const int32_t* tours,
                    const int32_t* lens, double* out)
{
    rtsp_setting_t st;
    st.swapping = true;
    st.genetic = false;
    st.ga_mutation_iter = 0;
    st.ga_population = 0;
    st.ga_generation = 0;
    st.ga_random_seed = false;
    EciGenSolver s(st);
    s.distance_matrix_ptr_ =
        std::make_shared<std::vector<std::vector<double>>>(K, std::vector<double>(K));
    for (int i = 0; i < K; ++i)
    {
        for (int j = 0; j < K; ++j) (*s.distance_matrix_ptr_)[i][j] = D[i * K + j];
    }
    for (int t = 0; t < n_tours; ++t)
    {
        std::vector<int> seq(tours + (size_t)t * stride, tours + (size_t)t * stride + lens[t]);
        out[t] = s.calculatePathCost_(seq);
    }
}

// Raw std::mt19937 words and the libstdc++ distributions the reference draws
// from them -- used to pin the device-side re-implementation.
void ref_mt19937_words(uint32_t seed, uint64_t n, uint32_t* out)
{
    std::mt19937 g(seed);
    for (uint64_t i = 0; i < n; ++i) out[i] = (uint32_t)g();
}

void ref_uniform_real(uint32_t seed, double a, double b, uint64_t n, double* out)
{
    std::mt19937 g(seed);
    std::uniform_real_distribution<> d(a, b);
    for (uint64_t i = 0; i < n; ++i) out[i] = d(g);
}

void ref_uniform_size_t(uint32_t seed, uint64_t lo, uint64_t hi, uint64_t n, uint64_t* out)
{
    std::mt19937 g(seed);
    std::uniform_int_distribution<size_t> d(lo, hi);
    for (uint64_t i = 0; i < n; ++i) out[i] = d(g);
}

void ref_uniform_int(uint32_t seed, int lo, int hi, uint64_t n, int32_t* out)
{
    std::mt19937 g(seed);
    std::uniform_int_distribution<> d(lo, hi);
    for (uint64_t i = 0; i < n; ++i) out[i] = d(g);
}

const char* ref_toolchain(void)
{
    static char buf[256];
    std::snprintf(buf, sizeof(buf), "g++ %d.%d.%d, libstdc++ %d, glibc %d.%d", __GNUC__,
                  __GNUC_MINOR__, __GNUC_PATCHLEVEL__, (int)__GLIBCXX__, __GLIBC__,
                  __GLIBC_MINOR__);
    return buf;
}

}  // extern "C"
