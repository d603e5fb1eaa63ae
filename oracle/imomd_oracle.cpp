// oracle/imomd_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's multi-tree expansion hot path and of the
// ECI-Gen relaxed-TSP solver, written against dense arrays and a CSR graph so
// that nothing depends on libstdc++ hash-container iteration order:
//
//   * adjacency order   = the CSR handed in (the reference containers' native
//                         iteration order, exported by the caller);
//   * children order    = the explicit 13-bucket rule of SURVEY.md section 8a-7
//                         (what unordered_set<size_t> does for <= 13 elements);
//   * frontier order    = irrelevant; exact distance ties go to the LOWEST
//                         VERTEX ID and are counted (orc_planner_tie_count);
//   * random numbers    = own mt19937 + the two libstdc++ formulas
//                         (generate_canonical, Lemire multiply-shift).
//
// Every function cites the reference lines it follows (paths under
// /root/reference).  PARITY PIN: tests/test_oracle_vs_ref.py compares this file
// with the unmodified reference (oracle/_ref) on bundled OSM maps, jittered
// grids and both fake maps: parents, costs (bitwise), matrices, flags, RTSP
// tours; tests/golden/ holds the vectors generated from the reference for the
// GPU box.  Arithmetic must match bit for bit, so this file is compiled with
// -ffp-contract=off and no -march (the reference's build.sh:2 has neither).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
// this library.

#include "imomd_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <unordered_set>
#include <utility>
#include <vector>

namespace {

const uint32_t NONE = 0xFFFFFFFFu;

// ------------------------------------------------------------- geometry ----

// include/osm_converter/compute_haversine.hpp:41-58 (EARTH_RADIUS is an int).
double haversine(double lat1, double lon1, double lat2, double lon2)
{
    static constexpr double DEG2RAD = M_PI / 180;
    static const int EARTH_RADIUS = 6371e3;
    double lat1_rad = lat1 * DEG2RAD;
    double lat2_rad = lat2 * DEG2RAD;
    double diff_lat_rad = (lat2 - lat1) * DEG2RAD;
    double diff_long_rad = (lon2 - lon1) * DEG2RAD;
    double a = std::pow(std::sin(diff_lat_rad / 2), 2) +
               std::cos(lat1_rad) * std::cos(lat2_rad) * std::pow(std::sin(diff_long_rad / 2), 2);
    double c = 2 * std::atan2(std::sqrt(a), std::sqrt(1 - a));
    return EARTH_RADIUS * c;
}

// include/osm_converter/compute_bearing.hpp:45-61.
double bearing(double lat1, double lon1, double lat2, double lon2)
{
    static constexpr double DEG2RAD = M_PI / 180;
    double lat1_rad = lat1 * DEG2RAD;
    double lat2_rad = lat2 * DEG2RAD;
    double diff_lon_rad = (lon2 - lon1) * DEG2RAD;
    double y = std::sin(diff_lon_rad) * std::cos(lat2_rad);
    double x = std::cos(lat1_rad) * std::sin(lat2_rad) -
               std::sin(lat1_rad) * std::cos(lat2_rad) * std::cos(diff_lon_rad);
    double b = std::atan2(y, x);
    b += b < 0 ? 2 * M_PI : 0;
    return b;
}

// --------------------------------------------------------------- random ----

// std::mt19937 (32-bit Mersenne twister, the engine the reference seeds with 0 at
// src/imomd_rrt_star.cpp:200 and src/eci_gen_tsp_solver.cpp:51).
struct Mt19937
{
    uint32_t mt[624];
    int idx;
    explicit Mt19937(uint32_t seed = 5489u) { reseed(seed); }
    void reseed(uint32_t seed)
    {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i)
        {
            mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        }
        idx = 624;
    }
    uint32_t next()
    {
        if (idx >= 624)
        {
            for (int i = 0; i < 624; ++i)
            {
                uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
};

// std::generate_canonical<double,53>(mt19937): two words, low word first
// (libstdc++ bits/random.tcc:3349-3381).
double canonical(Mt19937& g)
{
    double sum = 0.0, tmp = 1.0;
    for (int k = 0; k < 2; ++k)
    {
        sum += (double)g.next() * tmp;
        tmp *= 4294967296.0;
    }
    double ret = sum / tmp;
    if (ret >= 1.0) ret = std::nextafter(1.0, 0.0);
    return ret;
}

// std::uniform_real_distribution<double>(a, b)(g)
double uniform_real(Mt19937& g, double a, double b) { return canonical(g) * (b - a) + a; }

// std::uniform_int_distribution<T>(lo, lo+urange)(g) for urange < 2^32 - 1: Lemire's
// multiply-shift with rejection (libstdc++ bits/uniform_int_dist.h:270-283, 319-325).
uint64_t uniform_u(Mt19937& g, uint64_t urange)
{
    if (urange == 0xFFFFFFFFull) return g.next();
    if (urange > 0xFFFFFFFFull)
    {
        std::fprintf(stderr, "oracle: uniform range beyond 32 bits is not restated\n");
        std::abort();
    }
    uint32_t range = (uint32_t)(urange + 1);
    uint64_t product = (uint64_t)g.next() * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range)
    {
        uint32_t threshold = (uint32_t)(0u - range) % range;
        while (low < threshold)
        {
            product = (uint64_t)g.next() * (uint64_t)range;
            low = (uint32_t)product;
        }
    }
    return product >> 32;
}

int uniform_int(Mt19937& g, int lo, int hi)
{
    uint64_t urange = (uint64_t)(int64_t)hi - (uint64_t)(int64_t)lo;
    return (int)(uniform_u(g, urange) + (uint64_t)(int64_t)lo);
}

}  // namespace

// ---------------------------------------------------------------- graph ----

struct orc_graph
{
    uint32_t n;
    std::vector<uint32_t> row_ptr, col;
    std::vector<double> w, lat, lon;
    uint32_t deg(uint32_t v) const { return row_ptr[v + 1] - row_ptr[v]; }
    // (*connection_ptr_)[a][b]
    double weight(uint32_t a, uint32_t b) const
    {
        for (uint32_t e = row_ptr[a]; e < row_ptr[a + 1]; ++e)
        {
            if (col[e] == b) return w[e];
        }
        std::fprintf(stderr, "oracle: asymmetric adjacency %u -> %u\n", a, b);
        std::abort();
    }
};

namespace {

// One tree (include/imomd_rrt_star/imomd_rrt_star.h:66-124) on dense arrays.
struct Tree
{
    int id = 0;
    uint32_t root = 0;
    bool is_done = false;
    bool constructed = false;
    std::vector<uint32_t> parent;      // NONE = not visited
    std::vector<double> cost;          // 0.0 where absent (operator[] default)
    std::vector<uint32_t> chl;         // CSR-aligned child lists in container order
    std::vector<uint8_t> nchl;
    std::vector<uint32_t> fr_items;    // frontier ("expandables")
    std::vector<uint32_t> fr_pos;
    uint64_t size = 0;                 // parent.size()

    bool visited(uint32_t v) const { return parent[v] != NONE; }
    bool in_frontier(uint32_t v) const { return fr_pos[v] != NONE; }
    void fr_insert(uint32_t v)
    {
        if (fr_pos[v] != NONE) return;
        fr_pos[v] = (uint32_t)fr_items.size();
        fr_items.push_back(v);
    }
    void fr_erase(uint32_t v)
    {
        uint32_t p = fr_pos[v];
        if (p == NONE) return;
        uint32_t last = fr_items.back();
        fr_items[p] = last;
        fr_pos[last] = p;
        fr_items.pop_back();
        fr_pos[v] = NONE;
    }
};

struct MinHeur
{
    uint64_t id;
    double val;
};

}  // namespace

struct orc_planner
{
    const orc_graph* g;
    int K;
    double goal_bias;
    bool pseudo;
    std::vector<uint32_t> dest;
    std::vector<Tree> trees;
    std::vector<double> P;          // probability_matrix_
    std::vector<MinHeur> MH;        // expandables_min_heuristic_matrix_
    std::vector<double> D;          // distance_matrix_
    std::vector<uint64_t> CN;       // connection_node_matrix_
    std::vector<uint8_t> contact;   // connection_nodes_set_[idx] non-empty
    std::vector<std::pair<uint32_t, uint64_t>> conn_log;   // pseudo-mode set insertions
    std::vector<int> ds_parent;
    bool connected = false;
    bool dm_updated = false;
    Mt19937 rng{0};
    uint64_t ties = 0;
    uint64_t cdf_overrun = 0;

    // ---- batched mode (restatement of imomd_expand_batch's semantics, DESIGN.md "Batched mode"):
    // a STEP = B expansions of every tree.  (1) Every tree, independently: sample (fixed 8-word
    // slots of the mt19937{0} stream), nearest frontier vertex, jump-point hops, choose-parent,
    // frontier / min-heuristic maintenance -- the new vertex is only queued for rewiring.  (2) Per
    // tree, the rewiring as a label-correcting relaxation from the queued vertices: a vertex whose
    // cost went down offers cost + w to its visited neighbours; a neighbour takes the offer when
    // (cost, parent id) gets lexicographically smaller.  The fixed point does not depend on the
    // order of the offers.  (3) Connection checks of every vertex added or changed in the step
    // against the other trees with the FINAL costs: per pair the smallest sum (lowest vertex on
    // ties) replaces the distance when it beats it by more than 0.1.  (4) The selection-probability
    // tests in tree order and program order.  Children lists are not maintained in this mode.
    struct Event
    {
        int kind;          // 0 connection check, 1 selection-probability test
        uint32_t x;
        int other;
        double val;        // cost of x / min-heuristic value
    };
    bool defer = false;
    bool no_children = false;           // batched mode: child rows are neither read nor written
    std::vector<Event> events;          // of the tree being expanded
    std::vector<uint32_t> word_cache;   // outputs 0, 1, 2, ... of mt19937{0}
    Mt19937 word_gen{0};
    uint64_t batch_passes = 0;          // passes performed in batched mode so far

    uint32_t word(uint64_t i)
    {
        while (word_cache.size() <= i) word_cache.push_back(word_gen.next());
        return word_cache[(size_t)i];
    }

    // ---- children container order: SURVEY.md 8a-7 (unordered_set<size_t>, <= 13) ----
    void chl_insert(Tree& t, uint32_t p, uint32_t x)
    {
        if (no_children) return;
        uint32_t base = g->row_ptr[p];
        uint32_t n = t.nchl[p];
        uint32_t* row = &t.chl[base];
        for (uint32_t i = 0; i < n; ++i)
        {
            if (row[i] == x) return;
        }
        if (n >= g->deg(p) || n >= 13)
        {
            std::fprintf(stderr, "oracle: child list overflow at %u\n", p);
            std::abort();
        }
        uint32_t at = 0;
        for (uint32_t i = 0; i < n; ++i)
        {
            if (row[i] % 13 == x % 13)
            {
                at = i;
                break;
            }
        }
        for (uint32_t i = n; i > at; --i) row[i] = row[i - 1];
        row[at] = x;
        t.nchl[p] = (uint8_t)(n + 1);
    }
    void chl_erase(Tree& t, uint32_t p, uint32_t x)
    {
        if (no_children) return;
        uint32_t base = g->row_ptr[p];
        uint32_t n = t.nchl[p];
        uint32_t* row = &t.chl[base];
        for (uint32_t i = 0; i < n; ++i)
        {
            if (row[i] == x)
            {
                for (uint32_t j = i; j + 1 < n; ++j) row[j] = row[j + 1];
                t.nchl[p] = (uint8_t)(n - 1);
                return;
            }
        }
    }
    void chl_assign_single(Tree& t, uint32_t p, uint32_t x)
    {
        if (no_children) return;
        t.chl[g->row_ptr[p]] = x;
        t.nchl[p] = 1;
    }

    double hav_vv(uint32_t a, uint32_t b) const
    {
        return haversine(g->lat[a], g->lon[a], g->lat[b], g->lon[b]);
    }

    // tree_t::updateCost (include/imomd_rrt_star/imomd_rrt_star.h:102-123)
    void update_cost(Tree& t, uint32_t rewired, double new_cost, std::vector<uint32_t>& updated)
    {
        updated.clear();
        size_t head = 0;
        updated.push_back(rewired);      // the vector doubles as the FIFO queue
        while (head < updated.size())
        {
            uint32_t par = updated[head++];
            uint32_t base = g->row_ptr[par];
            for (uint32_t i = 0; i < t.nchl[par]; ++i)
            {
                uint32_t child = t.chl[base + i];
                t.cost[child] = new_cost + t.cost[child] - t.cost[rewired];
                updated.push_back(child);
            }
        }
        t.cost[rewired] = new_cost;
    }

    // src/imomd_rrt_star.cpp:358-400
    void select_random_vertex(Tree& t, uint64_t& rid, double& rlat, double& rlon)
    {
        if (uniform_real(rng, 0, 1) < goal_bias)
        {
            double rand = uniform_real(rng, 0, 1);
            int i = 0;
            while (rand >= 0)
            {
                if (i >= K)
                {
                    ++cdf_overrun;   // the reference reads past the row here (UB)
                    break;
                }
                rand -= P[(size_t)t.id * K + i++];
            }
            uint32_t random_dest = trees[i - 1].root;
            rid = random_dest;
            rlat = g->lat[random_dest];
            rlon = g->lon[random_dest];
            if (t.visited(random_dest))
            {
                double ratio = uniform_real(rng, 0.0, 0.5);
                rlat *= ratio;
                rlon *= ratio;
                rlat += (1.0 - ratio) * g->lat[t.root];
                rlon += (1.0 - ratio) * g->lon[t.root];
            }
        }
        else
        {
            uint32_t v = (uint32_t)uniform_u(rng, (uint64_t)g->n - 1);
            rid = v;
            rlat = g->lat[v];
            rlon = g->lon[v];
        }
    }

    // arg-min of src/imomd_rrt_star.cpp:406-418; ties -> lowest id (counted)
    bool nearest(const Tree& t, double qlat, double qlon, uint32_t& arg, double& best,
                 int* n_ties, double* second)
    {
        best = INFINITY;
        arg = NONE;
        double sec = INFINITY;
        int tc = 0;
        for (uint32_t v : t.fr_items)
        {
            double d = haversine(g->lat[v], g->lon[v], qlat, qlon);
            if (d < best)
            {
                sec = best;
                best = d;
                arg = v;
                tc = 1;
            }
            else if (d == best)
            {
                ++tc;
                if (v < arg) arg = v;
            }
            else if (d < sec)
            {
                sec = d;
            }
        }
        if (n_ties) *n_ties = tc;
        if (second) *second = sec;
        return !t.fr_items.empty();
    }

    // src/imomd_rrt_star.cpp:402-463 (x_nearest reported for the traces)
    uint32_t steer(Tree& t, double qlat, double qlon, uint32_t* x_nearest, double* nd)
    {
        uint32_t x_new;
        double best;
        int tc = 0;
        nearest(t, qlat, qlon, x_new, best, &tc, nullptr);
        if (tc > 1) ties += (uint64_t)(tc - 1);
        if (x_nearest) *x_nearest = x_new;
        if (nd) *nd = best;

        if (g->deg(x_new) == 2) update_expandables(t, x_new, false);

        bool jps_done = false;
        while (g->deg(x_new) == 2 && !jps_done)
        {
            uint32_t e = g->row_ptr[x_new];
            uint32_t x_1 = g->col[e];
            uint32_t x_2 = g->col[e + 1];
            if (t.visited(x_1))
            {
                if (t.visited(x_2))
                {
                    jps_done = true;
                }
                else
                {
                    attach_chain(t, x_new, x_1);
                    x_new = x_2;
                }
            }
            else
            {
                attach_chain(t, x_new, x_2);
                x_new = x_1;
            }
        }
        return x_new;
    }

    // the four statements of src/imomd_rrt_star.cpp:443-446 / :453-456
    void attach_chain(Tree& t, uint32_t x_new, uint32_t via)
    {
        if (!t.visited(x_new)) ++t.size;
        t.parent[x_new] = via;
        chl_assign_single(t, via, x_new);
        t.cost[x_new] = t.cost[via] + g->weight(via, x_new);
        update_connection(t, x_new);
    }

    // src/imomd_rrt_star.cpp:465-496
    void connect_new_node(Tree& t, uint32_t x_new)
    {
        uint32_t x_parent = NONE;
        double min_cost = INFINITY;
        bool connectable = false;
        for (uint32_t e = g->row_ptr[x_new]; e < g->row_ptr[x_new + 1]; ++e)
        {
            uint32_t y = g->col[e];
            if (t.visited(y))
            {
                double c = t.cost[y] + g->w[e];
                if (c < min_cost)
                {
                    x_parent = y;
                    min_cost = c;
                    connectable = true;
                }
            }
        }
        if (connectable)
        {
            if (!t.visited(x_new)) ++t.size;
            t.parent[x_new] = x_parent;
            chl_insert(t, x_parent, x_new);
            t.cost[x_new] = min_cost;
        }
    }

    // src/imomd_rrt_star.cpp:498-559.  `other.id` / `other.root` are read from the
    // tree objects as they are, so during construction (:73-78) trees that do not
    // exist yet still have id 0 and root 0.
    void update_expandables(Tree& t, uint32_t x_new, bool search_neighbor)
    {
        t.fr_erase(x_new);
        for (int o = 0; o < K; ++o)
        {
            const Tree& other = trees[o];
            if (other.id == t.id) continue;
            MinHeur& mh = MH[(size_t)t.id * K + other.id];
            if ((uint64_t)x_new == mh.id)
            {
                mh.id = (uint64_t)-1;
                mh.val = INFINITY;
                for (uint32_t v : t.fr_items)
                {
                    double h = hav_vv(v, t.root) + hav_vv(v, other.root);
                    if (h < mh.val)
                    {
                        mh.id = v;
                        mh.val = h;
                    }
                    else if (h == mh.val && mh.id != (uint64_t)-1)
                    {
                        ++ties;
                        if (v < mh.id) mh.id = v;
                    }
                }
            }
        }
        if (!search_neighbor) return;
        for (uint32_t e = g->row_ptr[x_new]; e < g->row_ptr[x_new + 1]; ++e)
        {
            uint32_t y = g->col[e];
            if (!t.visited(y))
            {
                t.fr_insert(y);
                for (int o = 0; o < K; ++o)
                {
                    const Tree& other = trees[o];
                    if (other.id == t.id) continue;
                    double h = hav_vv(y, t.root) + hav_vv(y, other.root);
                    MinHeur& mh = MH[(size_t)t.id * K + other.id];
                    if (h < mh.val)
                    {
                        mh.id = y;
                        mh.val = h;
                    }
                }
            }
        }
    }

    // src/imomd_rrt_star.cpp:561-600, recursion unrolled onto an explicit stack
    void rewire(Tree& t, uint32_t x_start)
    {
        struct Frame
        {
            uint32_t x;
            uint32_t e;           // next adjacency slot to examine
            uint32_t skip;        // x_near of the active rewire
            int64_t cur;          // cursor into `updated` (reverse order), -1 = none
            std::vector<uint32_t> updated;
        };
        std::vector<Frame> st;
        st.push_back(Frame{x_start, g->row_ptr[x_start], NONE, -1, {}});
        while (!st.empty())
        {
            Frame& f = st.back();
            if (f.cur >= 0)
            {
                uint32_t u = f.updated[(size_t)f.cur];
                --f.cur;
                if (u == f.skip) continue;
                st.push_back(Frame{u, g->row_ptr[u], NONE, -1, {}});
                continue;
            }
            if (f.e >= g->row_ptr[f.x + 1])
            {
                st.pop_back();
                continue;
            }
            uint32_t e = f.e++;
            uint32_t y = g->col[e];
            if (t.visited(y) && y != t.parent[f.x])
            {
                double nc = t.cost[f.x] + g->w[e];
                if (t.cost[y] > nc)
                {
                    chl_erase(t, t.parent[y], y);
                    t.parent[y] = f.x;
                    chl_insert(t, f.x, y);
                    update_cost(t, y, nc, f.updated);
                    update_connection(t, y);
                    f.skip = y;
                    f.cur = (int64_t)f.updated.size() - 1;
                }
            }
        }
    }

    // src/imomd_rrt_star.cpp:602-643
    void update_selection_probability(Tree& t, uint64_t rand_id)
    {
        int other_id = K;
        for (int i = 0; i < K; ++i)
        {
            if ((uint64_t)dest[i] == rand_id)
            {
                other_id = i;
                break;
            }
        }
        if (other_id == K) return;
        if (defer)
        {
            events.push_back(Event{1, NONE, other_id, MH[(size_t)t.id * K + other_id].val});
            return;
        }
        apply_selection_probability(t, other_id, MH[(size_t)t.id * K + other_id].val);
    }

    void apply_selection_probability(Tree& t, int other_id, double mh_val)
    {
        if (mh_val > D[(size_t)t.id * K + other_id])
        {
            Tree& other = trees[other_id];
            P[(size_t)t.id * K + other.id] = 0.0;
            P[(size_t)other.id * K + t.id] = 0.0;
            auto all_zero = [&](int r) {
                for (int i = 0; i < K; ++i)
                {
                    if (!(P[(size_t)r * K + i] == 0)) return false;
                }
                return true;
            };
            t.is_done = all_zero(t.id);
            other.is_done = all_zero(other.id);
            double sum_t = 0.0, sum_o = 0.0;
            for (int i = 0; i < K; ++i) sum_t += P[(size_t)t.id * K + i];
            for (int i = 0; i < K; ++i) sum_o += P[(size_t)other.id * K + i];
            for (int i = 0; i < K; ++i)
            {
                P[(size_t)t.id * K + i] /= sum_t;
                P[(size_t)other.id * K + i] /= sum_o;
            }
        }
    }

    // src/imomd_rrt_star.cpp:645-706
    void update_connection(Tree& t, uint32_t x_new)
    {
        if (defer)
        {
            events.push_back(Event{0, x_new, 0, t.cost[x_new]});
            return;
        }
        apply_connection(t, x_new, t.cost[x_new]);
    }

    // the body of updateConnectionTree_ with the tree's cost at x_new given
    void apply_connection(Tree& t, uint32_t x_new, double cost_here)
    {
        for (int o = 0; o < K; ++o)
        {
            Tree& other = trees[o];
            if (other.id == t.id) continue;
            if (!other.visited(x_new)) continue;
            int set_idx = (other.id > t.id) ? other.id * (other.id - 1) / 2 + t.id
                                            : t.id * (t.id - 1) / 2 + other.id;
            if (!contact[set_idx])
            {
                contact[set_idx] = 1;
                if (pseudo) conn_log.push_back({(uint32_t)set_idx, x_new});
                connect_two_trees(t.id, other.id);
            }
            double distance = cost_here + other.cost[x_new];
            if (distance < D[(size_t)t.id * K + other.id] - 0.1)
            {
                D[(size_t)t.id * K + other.id] = distance;
                D[(size_t)other.id * K + t.id] = distance;
                CN[(size_t)t.id * K + other.id] = x_new;
                CN[(size_t)other.id * K + t.id] = x_new;
                dm_updated = true;
            }
            if (pseudo)
            {
                uint32_t x_parent = t.parent[x_new];
                if (!other.visited(x_parent) ||
                    (other.parent[x_parent] != x_new && other.parent[x_new] != x_parent))
                {
                    conn_log.push_back({(uint32_t)set_idx, x_new});
                }
            }
        }
    }

    // src/imomd_rrt_star.cpp:708-742 (flat relabel, same resulting labels)
    void connect_two_trees(int a, int b)
    {
        if (connected) return;
        int lo = ds_parent[a], hi = ds_parent[b];
        if (lo == hi) return;
        if (hi < lo) std::swap(lo, hi);
        for (int i = 0; i < K; ++i)
        {
            if (ds_parent[i] == hi) ds_parent[i] = lo;
        }
        connected = true;
        for (int i = 1; i < K; ++i)
        {
            if (ds_parent[i] != ds_parent[0]) connected = false;
        }
    }

    // selectRandomVertex_ reading the stream from a fixed position (batched mode)
    struct SlotWords
    {
        orc_planner* p;
        uint64_t cursor;
        uint32_t next() { return p->word(cursor++); }
    };
    static double canonical_w(SlotWords& w)
    {
        double sum = 0.0, tmp = 1.0;
        for (int k = 0; k < 2; ++k)
        {
            sum += (double)w.next() * tmp;
            tmp *= 4294967296.0;
        }
        double ret = sum / tmp;
        if (ret >= 1.0) ret = std::nextafter(1.0, 0.0);
        return ret;
    }
    static uint32_t uniform_u32_w(SlotWords& w, uint32_t urange)
    {
        if (urange == 0xFFFFFFFFu) return w.next();
        uint32_t range = urange + 1u;
        uint64_t product = (uint64_t)w.next() * (uint64_t)range;
        uint32_t low = (uint32_t)product;
        if (low < range)
        {
            uint32_t threshold = (0u - range) % range;
            while (low < threshold)
            {
                product = (uint64_t)w.next() * (uint64_t)range;
                low = (uint32_t)product;
            }
        }
        return (uint32_t)(product >> 32);
    }
    void select_random_vertex_at(Tree& t, uint64_t cursor, uint64_t& rid, double& rlat, double& rlon)
    {
        SlotWords w{this, cursor};
        if (canonical_w(w) * (1.0 - 0.0) + 0.0 < goal_bias)
        {
            double rand = canonical_w(w) * (1.0 - 0.0) + 0.0;
            int i = 0;
            while (rand >= 0)
            {
                if (i >= K)
                {
                    ++cdf_overrun;
                    break;
                }
                rand -= P[(size_t)t.id * K + i++];
            }
            uint32_t random_dest = trees[i - 1].root;
            rid = random_dest;
            rlat = g->lat[random_dest];
            rlon = g->lon[random_dest];
            if (t.visited(random_dest))
            {
                double ratio = canonical_w(w) * (0.5 - 0.0) + 0.0;
                rlat *= ratio;
                rlon *= ratio;
                rlat += (1.0 - ratio) * g->lat[t.root];
                rlon += (1.0 - ratio) * g->lon[t.root];
            }
        }
        else
        {
            uint32_t v = uniform_u32_w(w, g->n - 1);
            rid = v;
            rlat = g->lat[v];
            rlon = g->lon[v];
        }
    }

    // rewiring of the batched mode: label-correcting relaxation from `work` (appended to as costs
    // go down); every vertex whose (cost, parent) changes is reported in `changed`
    void relax(Tree& t, std::vector<uint32_t>& work, std::vector<uint32_t>& changed)
    {
        size_t head = 0;
        while (head < work.size())
        {
            uint32_t u = work[head++];
            double cu = t.cost[u];
            for (uint32_t e = g->row_ptr[u]; e < g->row_ptr[u + 1]; ++e)
            {
                uint32_t y = g->col[e];
                if (!t.visited(y)) continue;
                double p = cu + g->w[e];
                if (p < t.cost[y] || (p == t.cost[y] && u < t.parent[y]))
                {
                    bool lower = p < t.cost[y];
                    t.cost[y] = p;
                    t.parent[y] = u;
                    changed.push_back(y);
                    if (lower) work.push_back(y);
                }
            }
        }
    }

    // one step of the batched mode: B expansions per tree (see the member comment above)
    uint64_t batched_step(int B)
    {
        uint64_t expansions = 0;
        std::vector<std::vector<Event>> plog(K);
        std::vector<std::vector<uint32_t>> check(K), dirty(K);
        std::vector<char> active(K);
        int n_active = 0;
        for (int k = 0; k < K; ++k)
        {
            active[k] = !trees[k].fr_items.empty() && !trees[k].is_done;
            n_active += active[k];
        }
        if (n_active == 0)
        {
            batch_passes += (uint64_t)B;
            return 0;
        }
        defer = true;
        no_children = true;
        for (int k = 0; k < K; ++k)
        {
            if (!active[k]) continue;
            Tree& t = trees[k];
            events.clear();
            for (int b = 0; b < B; ++b)
            {
                if (t.fr_items.empty()) break;
                uint64_t rid;
                double rlat, rlon;
                select_random_vertex_at(t, ((batch_passes + (uint64_t)b) * (uint64_t)K + (uint64_t)k) * 8u, rid, rlat, rlon);
                uint32_t x_new = steer(t, rlat, rlon, nullptr, nullptr);
                connect_new_node(t, x_new);
                update_expandables(t, x_new, true);
                update_selection_probability(t, rid);
                if (t.visited(x_new))
                {
                    dirty[k].push_back(x_new);
                    update_connection(t, x_new);
                }
                ++expansions;
            }
            for (const Event& e : events)
            {
                if (e.kind == 0) check[k].push_back(e.x);
                else plog[k].push_back(e);
            }
        }
        defer = false;
        for (int k = 0; k < K; ++k) relax(trees[k], dirty[k], check[k]);
        // connection checks with the final costs, per unordered pair of trees
        for (int a = 0; a < K; ++a)
        {
            for (int b = 0; b < a; ++b)
            {
                double best = INFINITY;
                uint32_t arg = NONE;
                bool touch = false;
                for (int side = 0; side < 2; ++side)
                {
                    const Tree& t = trees[side ? b : a];
                    const Tree& o = trees[side ? a : b];
                    for (uint32_t x : check[side ? b : a])
                    {
                        if (!o.visited(x)) continue;
                        touch = true;
                        double d = t.cost[x] + o.cost[x];
                        if (d < best || (d == best && x < arg))
                        {
                            best = d;
                            arg = x;
                        }
                    }
                }
                if (!touch) continue;
                int set_idx = a * (a - 1) / 2 + b;
                if (!contact[set_idx])
                {
                    contact[set_idx] = 1;
                    connect_two_trees(b, a);
                }
                if (best < D[(size_t)a * K + b] - 0.1)
                {
                    D[(size_t)a * K + b] = best;
                    D[(size_t)b * K + a] = best;
                    CN[(size_t)a * K + b] = arg;
                    CN[(size_t)b * K + a] = arg;
                    dm_updated = true;
                }
            }
        }
        for (int k = 0; k < K; ++k)
        {
            for (const Event& e : plog[k]) apply_selection_probability(trees[k], e.other, e.val);
        }
        batch_passes += (uint64_t)B;
        return expansions;
    }

    // src/imomd_rrt_star.cpp:337-356
    bool expand_tree(Tree& t, orc_trace* tr)
    {
        if (t.fr_items.empty() || t.is_done) return false;
        uint64_t rid;
        double rlat, rlon;
        select_random_vertex(t, rid, rlat, rlon);
        uint32_t xn;
        double nd;
        uint32_t x_new = steer(t, rlat, rlon, &xn, &nd);
        connect_new_node(t, x_new);
        update_expandables(t, x_new, true);
        rewire(t, x_new);
        update_selection_probability(t, rid);
        update_connection(t, x_new);
        if (tr)
        {
            tr->rand_id = rid;
            tr->rand_lat = rlat;
            tr->rand_lon = rlon;
            tr->x_nearest = xn;
            tr->nearest_dist = nd;
            tr->x_new = x_new;
            tr->parent_of_new = t.visited(x_new) ? t.parent[x_new] : (uint64_t)-1;
            tr->cost_of_new = t.visited(x_new) ? t.cost[x_new] : NAN;
            tr->tree_size_after = t.size;
            tr->frontier_after = t.fr_items.size();
        }
        return true;
    }
};

extern "C" {

double orc_haversine(double lat1, double lon1, double lat2, double lon2)
{
    return haversine(lat1, lon1, lat2, lon2);
}

double orc_bearing(double lat1, double lon1, double lat2, double lon2)
{
    return bearing(lat1, lon1, lat2, lon2);
}

void orc_mt19937_words(uint32_t seed, uint64_t n, uint32_t* out)
{
    Mt19937 g(seed);
    for (uint64_t i = 0; i < n; ++i) out[i] = g.next();
}

void orc_uniform_real(uint32_t seed, double a, double b, uint64_t n, double* out)
{
    Mt19937 g(seed);
    for (uint64_t i = 0; i < n; ++i) out[i] = uniform_real(g, a, b);
}

void orc_uniform_size_t(uint32_t seed, uint64_t lo, uint64_t hi, uint64_t n, uint64_t* out)
{
    Mt19937 g(seed);
    for (uint64_t i = 0; i < n; ++i) out[i] = uniform_u(g, hi - lo) + lo;
}

void orc_uniform_int(uint32_t seed, int lo, int hi, uint64_t n, int32_t* out)
{
    Mt19937 g(seed);
    for (uint64_t i = 0; i < n; ++i) out[i] = uniform_int(g, lo, hi);
}

orc_graph* orc_graph_create(uint32_t n, const uint32_t* row_ptr, const uint32_t* col,
                            const double* w, const double* lat, const double* lon)
{
    orc_graph* g = new orc_graph;
    g->n = n;
    g->row_ptr.assign(row_ptr, row_ptr + n + 1);
    uint32_t m = row_ptr[n];
    g->col.assign(col, col + m);
    g->w.assign(w, w + m);
    g->lat.assign(lat, lat + n);
    g->lon.assign(lon, lon + n);
    return g;
}

void orc_graph_destroy(orc_graph* g) { delete g; }

// Order in which libstdc++ iterates an unordered_map<size_t,double> with at most
// 13 keys: a new key goes to the head of its bucket's run (bucket = key % 13) or,
// when that bucket is empty, to the head of the whole list; duplicates are ignored
// (`insert` keeps the first weight).  Same rule as the children sets.
void orc_native_csr_from_edges(uint32_t n, uint64_t m, const uint64_t* eu, const uint64_t* ev,
                               const double* ew, uint32_t* row_ptr, uint32_t* col, double* w)
{
    std::vector<std::vector<std::pair<uint32_t, double>>> rows(n);
    auto ins = [&](uint32_t a, uint32_t b, double wt) {
        auto& r = rows[a];
        for (auto& kv : r)
        {
            if (kv.first == b) return;
        }
        if (r.size() >= 13)
        {
            std::fprintf(stderr, "oracle: degree above 13 is not restated (vertex %u)\n", a);
            std::abort();
        }
        size_t at = 0;
        for (size_t i = 0; i < r.size(); ++i)
        {
            if (r[i].first % 13 == b % 13)
            {
                at = i;
                break;
            }
        }
        r.insert(r.begin() + at, {b, wt});
    };
    for (uint64_t e = 0; e < m; ++e)
    {
        ins((uint32_t)eu[e], (uint32_t)ev[e], ew[e]);
        ins((uint32_t)ev[e], (uint32_t)eu[e], ew[e]);
    }
    uint32_t pos = 0;
    for (uint32_t v = 0; v < n; ++v)
    {
        row_ptr[v] = pos;
        for (auto& kv : rows[v])
        {
            col[pos] = kv.first;
            w[pos] = kv.second;
            ++pos;
        }
    }
    row_ptr[n] = pos;
}

// ImomdRRT::ImomdRRT, src/imomd_rrt_star.cpp:39-223
orc_planner* orc_planner_create(orc_graph* g, int K, const uint64_t* dest, double goal_bias,
                                int pseudo_mode)
{
    orc_planner* p = new orc_planner;
    p->g = g;
    p->K = K;
    p->goal_bias = goal_bias;
    p->pseudo = pseudo_mode != 0;
    p->dest.resize(K);
    for (int i = 0; i < K; ++i) p->dest[i] = (uint32_t)dest[i];
    p->trees.resize(K);   // value-initialised: id 0, root 0 (:58)
    p->ds_parent.assign(K, 0);
    p->P.assign((size_t)K * K, 0.0);
    p->MH.assign((size_t)K * K, MinHeur{(uint64_t)-1, INFINITY});
    p->D.assign((size_t)K * K, 0.0);
    p->CN.assign((size_t)K * K, (uint64_t)-1);
    p->contact.assign((size_t)K * (K - 1) / 2 + 1, 0);
    std::vector<std::vector<std::pair<double, int>>> bearing_matrix(K);
    std::vector<double> inv_hav((size_t)K * K, 0.0);
    std::vector<double> sum_hav(K, 0.0);
    uint32_t n = g->n;
    uint32_t m = g->row_ptr[n];
    for (int i = 0; i < K; ++i)
    {
        Tree& t = p->trees[i];
        uint32_t root = p->dest[i];
        t.id = i;
        t.root = root;
        t.is_done = false;
        t.constructed = true;
        t.parent.assign(n, NONE);
        t.cost.assign(n, 0.0);
        t.chl.assign(m, NONE);
        t.nchl.assign(n, 0);
        t.fr_pos.assign(n, NONE);
        t.parent[root] = root;
        t.cost[root] = 0;
        t.size = 1;
        p->update_expandables(t, root, true);
        p->ds_parent[i] = i;
        for (int j = i + 1; j < K; ++j)
        {
            uint32_t a = p->dest[i], b = p->dest[j];
            bearing_matrix[i].push_back({bearing(g->lat[a], g->lon[a], g->lat[b], g->lon[b]), j});
            bearing_matrix[j].push_back({bearing(g->lat[b], g->lon[b], g->lat[a], g->lon[a]), i});
            inv_hav[(size_t)i * K + j] = 1.0 / p->hav_vv(root, b);
            inv_hav[(size_t)j * K + i] = inv_hav[(size_t)i * K + j];
            sum_hav[i] += inv_hav[(size_t)i * K + j];
            sum_hav[j] += inv_hav[(size_t)j * K + i];
            p->D[(size_t)i * K + j] = INFINITY;
            p->D[(size_t)j * K + i] = INFINITY;
        }
    }
    // bearing term, only with more than three destinations (:124-153)
    for (int i = 0; i < K && K > 3; ++i)
    {
        std::sort(bearing_matrix[i].begin(), bearing_matrix[i].end());
        double sum_bearing = 0.0;
        for (int j = 0; j < K - 1; ++j)
        {
            double curr = bearing_matrix[i][j].first;
            double prev = bearing_matrix[i][(j - 2 + K) % (K - 1)].first;
            double next = bearing_matrix[i][(j + 1) % (K - 1)].first;
            double dprev = curr - prev;
            dprev += dprev < 0 ? 2 * M_PI : 0;
            double dnext = next - curr;
            dnext += dnext < 0 ? 2 * M_PI : 0;
            int idx = bearing_matrix[i][j].second;
            p->P[(size_t)i * K + idx] = dprev + dnext;
            sum_bearing += p->P[(size_t)i * K + idx];
        }
        for (int j = 0; j < K; ++j) p->P[(size_t)i * K + j] /= sum_bearing;
    }
    // inverse-haversine term (:157-163)
    for (int i = 0; i < K; ++i)
    {
        for (int j = 0; j < K; ++j) p->P[(size_t)i * K + j] += inv_hav[(size_t)i * K + j] / sum_hav[i];
    }
    // symmetrise and normalise (:169-185)
    std::vector<double> sum_p(K, 0.0);
    for (int i = 0; i < K; ++i)
    {
        for (int j = i + 1; j < K; ++j)
        {
            p->P[(size_t)i * K + j] += p->P[(size_t)j * K + i];
            p->P[(size_t)j * K + i] = p->P[(size_t)i * K + j];
            sum_p[i] += p->P[(size_t)i * K + j];
            sum_p[j] += p->P[(size_t)j * K + i];
        }
        for (int j = 0; j < K; ++j) p->P[(size_t)i * K + j] /= sum_p[i];
    }
    p->rng.reseed(0);
    return p;
}

void orc_planner_destroy(orc_planner* p) { delete p; }

uint64_t orc_planner_iterate(orc_planner* p, int iters)
{
    uint64_t expansions = 0;
    for (int it = 0; it < iters; ++it)
    {
        for (int k = 0; k < p->K; ++k)
        {
            if (p->expand_tree(p->trees[k], nullptr)) ++expansions;
        }
    }
    return expansions;
}

// `steps` steps of the batched mode (B expansions per tree per step); returns the expansions
uint64_t orc_planner_iterate_batched(orc_planner* p, int steps, int B)
{
    uint64_t expansions = 0;
    for (int s = 0; s < steps; ++s) expansions += p->batched_step(B);
    return expansions;
}

int orc_planner_expand_traced(orc_planner* p, int k, orc_trace* tr)
{
    std::memset(tr, 0, sizeof(*tr));
    tr->tree = k;
    if (!p->expand_tree(p->trees[k], tr))
    {
        tr->skipped = 1;
        return 0;
    }
    return 1;
}

void orc_planner_get_tree(orc_planner* p, int k, uint32_t* parent, double* cost)
{
    const Tree& t = p->trees[k];
    for (uint32_t v = 0; v < p->g->n; ++v)
    {
        parent[v] = t.parent[v];
        cost[v] = t.visited(v) ? t.cost[v] : INFINITY;
    }
}

uint64_t orc_planner_tree_size(orc_planner* p, int k) { return p->trees[k].size; }

int orc_planner_children(orc_planner* p, int k, uint64_t v, uint64_t* out, int cap)
{
    const Tree& t = p->trees[k];
    int n = t.nchl[v];
    for (int i = 0; i < n && i < cap; ++i) out[i] = t.chl[p->g->row_ptr[v] + i];
    return n;
}

uint64_t orc_planner_treehash(orc_planner* p)
{
    uint64_t h = 1469598103934665603ull;
    for (const Tree& t : p->trees)
    {
        for (uint32_t v = 0; v < p->g->n; ++v)
        {
            if (!t.visited(v)) continue;
            uint64_t bits;
            std::memcpy(&bits, &t.cost[v], 8);
            h = (h ^ (uint64_t)v) * 1099511628211ull;
            h = (h ^ (uint64_t)t.parent[v]) * 1099511628211ull;
            h = (h ^ bits) * 1099511628211ull;
        }
    }
    return h;
}

void orc_planner_get_matrices(orc_planner* p, double* D, uint64_t* conn_node, double* prob,
                              uint64_t* mh_id, double* mh_val)
{
    size_t kk = (size_t)p->K * p->K;
    for (size_t i = 0; i < kk; ++i)
    {
        if (D) D[i] = p->D[i];
        if (conn_node) conn_node[i] = p->CN[i];
        if (prob) prob[i] = p->P[i];
        if (mh_id) mh_id[i] = p->MH[i].id;
        if (mh_val) mh_val[i] = p->MH[i].val;
    }
}

void orc_planner_get_flags(orc_planner* p, int32_t* is_done, int32_t* connected,
                           int32_t* dm_updated, int32_t* ds_parent)
{
    for (int k = 0; k < p->K; ++k)
    {
        if (is_done) is_done[k] = p->trees[k].is_done ? 1 : 0;
        if (ds_parent) ds_parent[k] = p->ds_parent[k];
    }
    if (connected) *connected = p->connected ? 1 : 0;
    if (dm_updated) *dm_updated = p->dm_updated ? 1 : 0;
}

uint64_t orc_planner_frontier(orc_planner* p, int k, uint64_t* out, uint64_t cap)
{
    const Tree& t = p->trees[k];
    for (uint64_t i = 0; i < t.fr_items.size() && i < cap; ++i) out[i] = t.fr_items[i];
    return t.fr_items.size();
}

int orc_planner_nearest(orc_planner* p, int k, double lat, double lon, uint64_t* idx,
                        double* dist, int32_t* n_ties, double* second)
{
    uint32_t arg;
    double best;
    int tc;
    double sec;
    bool ok = p->nearest(p->trees[k], lat, lon, arg, best, &tc, &sec);
    *idx = arg == NONE ? (uint64_t)-1 : arg;
    *dist = best;
    if (n_ties) *n_ties = tc;
    if (second) *second = sec;
    return ok ? 1 : 0;
}

uint64_t orc_planner_tie_count(orc_planner* p) { return p->ties; }

uint64_t orc_planner_connection_log(orc_planner* p, uint32_t* set_idx, uint64_t* vertex,
                                    uint64_t cap)
{
    for (uint64_t i = 0; i < p->conn_log.size() && i < cap; ++i)
    {
        set_idx[i] = p->conn_log[i].first;
        vertex[i] = p->conn_log[i].second;
    }
    return p->conn_log.size();
}

}  // extern "C"

// ============================================================== ECI-Gen =====

struct orc_solver
{
    bool swapping, genetic;
    int ga_mutation_iter, ga_population, ga_generation;
    Mt19937 rng{0};   // persists across solves (src/eci_gen_tsp_solver.cpp:39-53)

    int K = 0;
    const double* D = nullptr;
    int source = 0, target = 0;
    double min_path_cost = 0;
    std::vector<int> seq_rtsp, seq_ins, seq_dij;
    std::vector<double> fitness;
    std::vector<std::vector<int>> chromosomes;

    double d(int i, int j) const { return D[(size_t)i * K + j]; }

    // :442-450
    double path_cost(const std::vector<int>& s) const
    {
        double distance = 0;
        for (size_t i = 0; i + 1 < s.size(); ++i) distance += d(s[i], s[i + 1]);
        return distance;
    }

    // :207-260 (priority_queue of (-cost, id), no closed set)
    double dijkstra()
    {
        std::vector<int> parent(K, -1);
        std::vector<double> cost(K, 0.0);
        std::vector<uint8_t> has(K, 0);
        parent[source] = source;
        cost[target] = INFINITY;   // {target, inf} is dropped by the map when target == source
        has[target] = 1;
        cost[source] = 0;
        has[source] = 1;
        std::priority_queue<std::pair<double, int>> open;
        open.push({0, source});
        while (!open.empty())
        {
            std::pair<double, int> state = open.top();
            open.pop();
            if (state.second == target) break;
            for (int i = 0; i < K; ++i)
            {
                if (state.second == i) continue;
                double new_cost = cost[state.second] + d(state.second, i);
                if (!has[i])
                {
                    cost[i] = INFINITY;
                    has[i] = 1;
                }
                if (new_cost < cost[i])
                {
                    cost[i] = new_cost;
                    parent[i] = state.second;
                    open.push({-new_cost, i});
                }
            }
        }
        // extractSequence_ :428-440
        seq_dij.clear();
        seq_dij.push_back(target);
        int state = target;
        while (state != source)
        {
            state = parent[state];
            seq_dij.push_back(state);
        }
        std::reverse(seq_dij.begin(), seq_dij.end());
        return cost[target];
    }

    enum Ins { REVISIT, SEQUENCE, SWAP_LEFT, SWAP_RIGHT, SWAP_BOTH };

    // :262-374
    void best_insertion(int ins, double& min_cost, int& best_left, Ins& method) const
    {
        min_cost = INFINITY;
        best_left = -1;
        method = REVISIT;
        int L = (int)seq_ins.size();
        for (int it = 0; it < L; ++it)
        {
            double revisit = 2 * d(ins, seq_ins[it]);
            double seqc = INFINITY, left = INFINITY, right = INFINITY, both = INFINITY;
            if (L - it == 1)
            {
                seqc = revisit;
            }
            else
            {
                int nx = it + 1;
                seqc = d(seq_ins[it], ins) + d(ins, seq_ins[nx]) - d(seq_ins[it], seq_ins[nx]);
                if (swapping)
                {
                    if (it > 1)
                    {
                        int pv = it - 1, ppv = it - 2;
                        left = d(ins, seq_ins[nx]) - d(seq_ins[it], seq_ins[nx]) +
                               d(seq_ins[pv], ins) + d(seq_ins[ppv], seq_ins[it]) -
                               d(seq_ins[ppv], seq_ins[pv]);
                    }
                    if (L - nx > 2)
                    {
                        int nnx = nx + 1, nnnx = nx + 2;
                        right = d(seq_ins[it], ins) - d(seq_ins[it], seq_ins[nx]) +
                                d(seq_ins[nnx], ins) + d(seq_ins[nnnx], seq_ins[nx]) -
                                d(seq_ins[nnnx], seq_ins[nnx]);
                    }
                    if (it > 1 && L - nx > 2)
                    {
                        int pv = it - 1, ppv = it - 2, nnx = nx + 1, nnnx = nx + 2;
                        both = d(seq_ins[pv], ins) + d(seq_ins[ppv], seq_ins[it]) -
                               d(seq_ins[ppv], seq_ins[pv]) - d(seq_ins[it], seq_ins[nx]) +
                               d(seq_ins[nnx], ins) + d(seq_ins[nnnx], seq_ins[nx]) -
                               d(seq_ins[nnnx], seq_ins[nnx]);
                    }
                }
            }
            if (revisit < min_cost)
            {
                best_left = it;
                min_cost = revisit;
                method = REVISIT;
            }
            if (seqc < min_cost)
            {
                best_left = it;
                min_cost = seqc;
                method = SEQUENCE;
            }
            if (swapping)
            {
                if (left < min_cost)
                {
                    best_left = it;
                    min_cost = left;
                    method = SWAP_LEFT;
                }
                if (right < min_cost)
                {
                    best_left = it;
                    min_cost = right;
                    method = SWAP_RIGHT;
                }
                if (both < min_cost)
                {
                    best_left = it;
                    min_cost = both;
                    method = SWAP_BOTH;
                }
            }
        }
    }

    // :103-205 with the five insert helpers (:376-426) restated on a vector
    void solve_best_insertion()
    {
        double path = dijkstra();
        seq_ins = seq_dij;
        std::vector<uint8_t> not_visited(K, 1);
        int remaining = K;
        for (int v : seq_dij)
        {
            if (not_visited[v])
            {
                not_visited[v] = 0;
                --remaining;
            }
        }
        // Iteration order of upstream's unordered_set<int> filled with 0..K-1 (:112-116):
        // identity hash, every key lands in an empty bucket and goes to the list head, the
        // list is reversed by each rehash (13 -> 29 -> 59 buckets).  Erasing keeps the order.
        std::vector<int> order;
        if (K <= 13)
        {
            for (int v = K - 1; v >= 0; --v) order.push_back(v);
        }
        else if (K <= 29)
        {
            for (int v = K - 1; v >= 13; --v) order.push_back(v);
            for (int v = 0; v <= 12; ++v) order.push_back(v);
        }
        else
        {
            for (int v = K - 1; v >= 29; --v) order.push_back(v);
            for (int v = 12; v >= 0; --v) order.push_back(v);
            for (int v = 13; v <= 28; ++v) order.push_back(v);
        }
        while (remaining > 0)
        {
            double min_ins = INFINITY;
            int left_idx = -1, id_insert = -1;
            Ins method = REVISIT;
            // exact cost ties between different destinations go to the first one met
            for (int cand : order)
            {
                if (!not_visited[cand]) continue;
                double c;
                int l;
                Ins m;
                best_insertion(cand, c, l, m);
                if (c < min_ins)
                {
                    id_insert = cand;
                    left_idx = l;
                    method = m;
                    min_ins = c;
                }
            }
            if (id_insert < 0)
            {
                std::fprintf(stderr, "oracle: solveRTSP on a disconnected matrix (UB upstream)\n");
                std::abort();
            }
            std::vector<int>& s = seq_ins;
            switch (method)
            {
                case REVISIT:   // A -> A, value, A
                {
                    int a = s[left_idx];
                    s.insert(s.begin() + left_idx + 1, {id_insert, a});
                    break;
                }
                case SEQUENCE:
                    s.insert(s.begin() + left_idx + 1, id_insert);
                    break;
                case SWAP_LEFT:
                    std::swap(s[left_idx], s[left_idx - 1]);
                    s.insert(s.begin() + left_idx + 1, id_insert);
                    break;
                case SWAP_RIGHT:
                    std::swap(s[left_idx + 1], s[left_idx + 2]);
                    s.insert(s.begin() + left_idx + 1, id_insert);
                    break;
                case SWAP_BOTH:
                    std::swap(s[left_idx], s[left_idx - 1]);
                    std::swap(s[left_idx + 1], s[left_idx + 2]);
                    s.insert(s.begin() + left_idx + 1, id_insert);
                    break;
            }
            path += min_ins;
            not_visited[id_insert] = 0;
            --remaining;
        }
        min_path_cost = path;
        seq_rtsp = seq_ins;
        if ((int)seq_rtsp.size() > K) refine(seq_rtsp);
    }

    // :827-882
    void refine(std::vector<int>& sequence) const
    {
        int n = (int)sequence.size();
        std::vector<int> last_idx(K, 0);
        for (int i = 0; i < n; ++i) last_idx[sequence[i]] = i;
        std::vector<uint8_t> visited(K, 0);
        std::vector<int> refined;
        for (int it = 0; it != n - 1; ++it)
        {
            if (visited[sequence[it]])
            {
                bool shortcut = false;
                int nx = it + 1;
                while (std::isinf(d(refined.back(), sequence[nx])) && nx != n - 1 &&
                       (visited[sequence[nx]] || nx < last_idx[sequence[nx]]))
                {
                    nx += 1;
                    if (!visited[sequence[nx]] && !std::isinf(d(refined.back(), sequence[nx])))
                    {
                        refined.push_back(sequence[nx]);
                        visited[sequence[nx]] = 1;
                        shortcut = true;
                        it = --nx;
                    }
                }
                if (!shortcut && (refined.back() != sequence[it]))
                {
                    refined.push_back(sequence[it]);
                    visited[sequence[it]] = 1;
                }
            }
            else
            {
                refined.push_back(sequence[it]);
                visited[sequence[it]] = 1;
            }
        }
        if (refined.back() != sequence.back()) refined.push_back(sequence.back());
        sequence = refined;
    }

    // :801-810 (`i + 0x9e3779b9` is evaluated in unsigned int)
    static uint64_t hash_chromosome(const std::vector<int>& c)
    {
        uint64_t h = c.size();
        for (int i : c)
        {
            uint32_t t = (uint32_t)i + 0x9e3779b9u;
            h ^= (uint64_t)t + (h << 6) + (h >> 2);
        }
        return h;
    }

    // :660-788
    void generate_chromosomes()
    {
        const std::vector<int> original = seq_rtsp;
        int count = 0, min_idx = 0;
        double tmp_min = min_path_cost;
        std::unordered_set<uint64_t> seen;   // membership only, order never observed
        int hi = (int)seq_rtsp.size() - 1;
        for (int i = 0; i < ga_mutation_iter; ++i)
        {
            int slice[4];
            for (int j = 0; j < 4; ++j) slice[j] = uniform_int(rng, 1, hi);
            std::sort(slice, slice + 4);
            std::vector<int> part1(original.begin(), original.begin() + slice[0]);
            std::vector<int> part2(original.begin() + slice[0], original.begin() + slice[1]);
            std::vector<int> part3(original.begin() + slice[1], original.begin() + slice[2]);
            std::vector<int> part4(original.begin() + slice[2], original.begin() + slice[3]);
            std::vector<int> part5(original.begin() + slice[3], original.end());
            int rev = uniform_int(rng, 0, 7);
            bool r2 = rev == 1 || rev == 4 || rev == 6 || rev == 7;
            bool r3 = rev == 2 || rev == 4 || rev == 5 || rev == 7;
            bool r4 = rev == 3 || rev == 5 || rev == 6 || rev == 7;
            if (r2) std::reverse(part2.begin(), part2.end());
            if (r3) std::reverse(part3.begin(), part3.end());
            if (r4) std::reverse(part4.begin(), part4.end());
            int order = uniform_int(rng, 0, 4);
            const std::vector<int>* seqp[5] = {&part1, nullptr, nullptr, nullptr, &part5};
            switch (order)
            {
                case 0: seqp[1] = &part2; seqp[2] = &part4; seqp[3] = &part3; break;
                case 1: seqp[1] = &part3; seqp[2] = &part2; seqp[3] = &part4; break;
                case 2: seqp[1] = &part3; seqp[2] = &part4; seqp[3] = &part2; break;
                case 3: seqp[1] = &part4; seqp[2] = &part2; seqp[3] = &part3; break;
                default: seqp[1] = &part4; seqp[2] = &part3; seqp[3] = &part2; break;
            }
            std::vector<int> chrom;
            for (int q = 0; q < 5; ++q) chrom.insert(chrom.end(), seqp[q]->begin(), seqp[q]->end());
            if ((int)chrom.size() > K) refine(chrom);
            uint64_t h = hash_chromosome(chrom);
            if (seen.count(h)) continue;
            seen.insert(h);
            double pc = path_cost(chrom);
            if (pc <= min_path_cost * 1.1)
            {
                fitness.push_back(1 / pc);
                chromosomes.push_back(chrom);
                if (pc < tmp_min)
                {
                    tmp_min = pc;
                    min_idx = count;
                }
                ++count;
            }
        }
        if (!chromosomes.empty())
        {
            min_path_cost = tmp_min;
            seq_rtsp = chromosomes[min_idx];
        }
    }

    // :502-658
    void solve_genetic()
    {
        generate_chromosomes();
        if (chromosomes.size() < 2) return;
        for (int gen = 0; gen < ga_generation && chromosomes.size() >= 2; ++gen)
        {
            int offspring_count = 0, min_idx = 0;
            double tmp_min = min_path_cost;
            std::unordered_set<uint64_t> seen;
            std::vector<double> tmp_fitness;
            std::vector<std::vector<int>> offsprings;
            double sum_fit = 0.0;
            for (double f : fitness) sum_fit += f;
            for (int i = 0; i < ga_population; ++i)
            {
                double rnd_a = uniform_real(rng, 0.0, sum_fit);
                int a = 0;
                while (rnd_a >= 0 && a < (int)fitness.size()) rnd_a -= fitness[a++];
                const std::vector<int>& pa = chromosomes[a - 1];
                double rnd_b = uniform_real(rng, 0.0, sum_fit);
                int b = 0;
                while (rnd_b >= 0 && b < (int)fitness.size()) rnd_b -= fitness[b++];
                const std::vector<int>& pb = chromosomes[b - 1];

                int s1 = uniform_int(rng, 1, (int)pa.size() - 1);
                int s2 = uniform_int(rng, 1, (int)pa.size() - 1);
                if (s2 < s1) std::swap(s1, s2);
                std::vector<uint8_t> in_sub(K, 0);
                for (int q = s1; q < s2; ++q) in_sub[pa[q]] = 1;
                int offset = uniform_int(rng, 1 - s1,
                                         (int)std::max(pa.size(), pb.size()) - s2 - 1);
                std::vector<int> off;
                off.push_back(source);
                int idx = 1, b_idx = 1;
                while (idx < offset + s1 && b_idx < (int)pb.size() - 1)
                {
                    if (!in_sub[pb[b_idx]])
                    {
                        off.push_back(pb[b_idx]);
                        idx++;
                    }
                    b_idx++;
                }
                if (uniform_int(rng, 0, 1))
                {
                    for (idx = offset + s1; idx < offset + s2; ++idx) off.push_back(pa[idx - offset]);
                }
                else
                {
                    for (idx = offset + s2 - 1; idx > offset + s1 - 1; --idx)
                    {
                        off.push_back(pa[idx - offset]);
                    }
                }
                idx = offset + s2;
                while (idx >= offset + s2 && b_idx < (int)pb.size() - 1)
                {
                    if (!in_sub[pb[b_idx]])
                    {
                        off.push_back(pb[b_idx]);
                        idx++;
                    }
                    b_idx++;
                }
                if (off.back() != target) off.push_back(target);
                if ((int)off.size() > K) refine(off);
                uint64_t h = hash_chromosome(off);
                if (seen.count(h)) continue;
                seen.insert(h);
                double pc = path_cost(off);
                if (pc <= min_path_cost * 1.1)
                {
                    tmp_fitness.push_back(1 / pc);
                    offsprings.push_back(off);
                    if (pc < tmp_min)
                    {
                        tmp_min = pc;
                        min_idx = offspring_count;
                    }
                    offspring_count++;
                }
            }
            fitness = tmp_fitness;
            chromosomes = offsprings;
            if (!chromosomes.empty())
            {
                min_path_cost = tmp_min;
                seq_rtsp = chromosomes[min_idx];
            }
        }
    }

    // :76-101
    double solve(int K_, const double* D_, int src, int tgt)
    {
        K = K_;
        D = D_;
        source = src;
        target = tgt;
        min_path_cost = 0;
        seq_rtsp.clear();
        seq_ins.clear();
        seq_dij.clear();
        chromosomes.clear();
        fitness.clear();
        solve_best_insertion();
        if (genetic) solve_genetic();
        return min_path_cost;
    }
};

extern "C" {

orc_solver* orc_solver_create(int swapping, int genetic, int ga_mutation_iter, int ga_population,
                              int ga_generation)
{
    orc_solver* s = new orc_solver;
    s->swapping = swapping != 0;
    s->genetic = genetic != 0;
    s->ga_mutation_iter = ga_mutation_iter;
    s->ga_population = ga_population;
    s->ga_generation = ga_generation;
    return s;
}

void orc_solver_destroy(orc_solver* s) { delete s; }

int orc_solver_solve(orc_solver* s, int K, const double* D, int source, int target, double* cost,
                     int32_t* seq, int cap)
{
    *cost = s->solve(K, D, source, target);
    int n = (int)s->seq_rtsp.size();
    for (int i = 0; i < n && i < cap; ++i) seq[i] = s->seq_rtsp[i];
    return n;
}

void orc_path_costs(int K, const double* D, int n_tours, int stride, const int32_t* tours,
                    const int32_t* lens, double* out)
{
    for (int t = 0; t < n_tours; ++t)
    {
        double distance = 0;
        const int32_t* s = tours + (size_t)t * stride;
        for (int i = 0; i + 1 < lens[t]; ++i) distance += D[(size_t)s[i] * K + s[i + 1]];
        out[t] = distance;
    }
}

}  // extern "C"
