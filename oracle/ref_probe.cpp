// oracle/ref_probe.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Command-line driver over oracle/ref_shim.cpp (i.e. over the unmodified
// reference).  A separate process per run is needed because
// ImomdRRT::selectRandomVertex_ keeps function-local static distributions whose
// range is frozen by the first planner of the process
// (/root/reference/src/imomd_rrt_star.cpp:360-362).
//
//   ref_probe (--graph file.bin | --grid W | --osm path | --fake T)
//             (--dest s,o1,..,t | --pick seed:count)
//             [--iters N] [--goal-bias g] [--pseudo] [--mode expand|run]
//             [--max-time s] [--max-iter n] [--dump out.bin]
//   ref_probe --mode rtsp --matrix D.bin [--solves n] [--ga mut,pop,gen]
//
// Prints one JSON line.  `--mode expand` times `iters` calls of
// expandTreeLayers_() (the expansion hot path, RTSP excluded); `--mode run`
// executes the reference's anytime loop and reports the log rows.

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

extern "C" {
void* ref_graph_from_edges(uint64_t n, const double* lat, const double* lon, uint64_t m,
                           const uint64_t* eu, const uint64_t* ev, const double* ew);
void* ref_graph_grid(int W);
void* ref_graph_from_osm(const char* path);
void* ref_graph_fake(int type);
uint64_t ref_graph_num_nodes(void* g);
uint64_t ref_graph_num_arcs(void* g);
void ref_pick_destinations(void* g, uint32_t seed, int count, uint64_t* out);
void* ref_planner_create(void* g, uint64_t source, uint64_t target, const uint64_t* objectives,
                         int n_obj, double goal_bias, int pseudo_mode, int max_iter,
                         int max_time, int swapping, int genetic, int ga_mutation_iter,
                         int ga_population, int ga_generation);
uint64_t ref_planner_iterate(void* p, int iters);
uint64_t ref_planner_tree_size(void* p, int k);
uint64_t ref_planner_treehash(void* p);
void ref_planner_get_tree(void* p, int k, uint32_t* parent, double* cost);
void ref_planner_get_matrices(void* p, double* D, uint64_t* conn_node, double* prob,
                              uint64_t* mh_id, double* mh_val);
struct ref_log_row_t
{
    double time_s;
    double path_cost;
    int64_t tree_size;
    int64_t iteration;
};
int ref_planner_run(void* p, ref_log_row_t* rows, int cap, int64_t* iterations,
                    double* elapsed_s, int64_t* expansions);
double ref_planner_path_cost(void* p);
uint64_t ref_planner_path(void* p, uint64_t* out, uint64_t cap);
const char* ref_toolchain(void);
void* ref_solver_create(int swapping, int genetic, int ga_mutation_iter, int ga_population,
                        int ga_generation);
int ref_solver_solve(void* sp, int K, const double* D, int source, int target, double* cost,
                     int32_t* seq, int cap);
}

// --mode rtsp: times EciGenSolver::solveRTSP (src/eci_gen_tsp_solver.cpp:76-101) of the unmodified
// reference on a K x K matrix read from a file {u64 K, f64 D[K*K]}: `solves` consecutive calls on
// one solver (its mt19937 persists across calls, as in the planner).
static int rtsp_mode(const char* matrix_file, int solves, int ga_mut, int ga_pop, int ga_gen)
{
    FILE* f = std::fopen(matrix_file, "rb");
    if (!f)
    {
        std::perror(matrix_file);
        return 2;
    }
    uint64_t K = 0;
    if (std::fread(&K, 8, 1, f) != 1 || K < 2 || K > 64)
    {
        std::fprintf(stderr, "bad matrix file\n");
        return 2;
    }
    std::vector<double> D(K * K);
    if (std::fread(D.data(), 8, K * K, f) != K * K)
    {
        std::fprintf(stderr, "short matrix file\n");
        return 2;
    }
    std::fclose(f);
    void* s = ref_solver_create(1, 1, ga_mut, ga_pop, ga_gen);
    std::vector<int32_t> seq(4096);
    std::vector<double> costs, ms;
    int len = 0;
    for (int i = 0; i < solves; ++i)
    {
        double cost = 0;
        auto t0 = std::chrono::steady_clock::now();
        len = ref_solver_solve(s, (int)K, D.data(), 0, (int)K - 1, &cost, seq.data(), (int)seq.size());
        ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        costs.push_back(cost);
    }
    std::printf("{\"toolchain\": \"%s\", \"mode\": \"rtsp\", \"K\": %llu, \"solves\": %d, \"ms\": [", ref_toolchain(),
                (unsigned long long)K, solves);
    for (size_t i = 0; i < ms.size(); ++i) std::printf("%s%.4f", i ? ", " : "", ms[i]);
    std::printf("], \"costs\": [");
    for (size_t i = 0; i < costs.size(); ++i) std::printf("%s%.17g", i ? ", " : "", costs[i]);
    std::printf("], \"last_sequence\": [");
    for (int i = 0; i < len; ++i) std::printf("%s%d", i ? ", " : "", seq[i]);
    std::printf("]}\n");
    std::fflush(stdout);
    _exit(0);
}

static void* load_graph_file(const char* path)
{
    FILE* f = std::fopen(path, "rb");
    if (!f)
    {
        std::perror(path);
        std::exit(2);
    }
    uint64_t hdr[4];
    if (std::fread(hdr, 8, 4, f) != 4 || hdr[0] != 0x315247444D4F4D49ull /* "IMOMDGR1" little-endian */)
    {
        std::fprintf(stderr, "bad graph file %s\n", path);
        std::exit(2);
    }
    uint64_t n = hdr[1], m = hdr[2], has_w = hdr[3];
    std::vector<double> lat(n), lon(n), ew(has_w ? m : 0);
    std::vector<uint64_t> eu(m), ev(m);
    bool ok = std::fread(lat.data(), 8, n, f) == n && std::fread(lon.data(), 8, n, f) == n &&
              std::fread(eu.data(), 8, m, f) == m && std::fread(ev.data(), 8, m, f) == m;
    if (ok && has_w) ok = std::fread(ew.data(), 8, m, f) == m;
    std::fclose(f);
    if (!ok)
    {
        std::fprintf(stderr, "short graph file %s\n", path);
        std::exit(2);
    }
    return ref_graph_from_edges(n, lat.data(), lon.data(), m, eu.data(), ev.data(),
                                has_w ? ew.data() : nullptr);
}

int main(int argc, char** argv)
{
    void* g = nullptr;
    std::vector<uint64_t> dest;
    int iters = 1000, pseudo = 0, max_time = 60, max_iter = 1000000;
    double goal_bias = 1.0;
    std::string mode = "expand", dump, matrix;
    int solves = 5;
    int ga_mut = 10000, ga_pop = 1000, ga_gen = 5;
    uint32_t pick_seed = 0;
    int pick_count = 0;
    for (int i = 1; i < argc; ++i)
    {
        std::string a = argv[i];
        auto next = [&]() -> const char* {
            if (i + 1 >= argc)
            {
                std::fprintf(stderr, "missing value for %s\n", a.c_str());
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "--graph") g = load_graph_file(next());
        else if (a == "--grid") g = ref_graph_grid(std::atoi(next()));
        else if (a == "--osm") g = ref_graph_from_osm(next());
        else if (a == "--fake") g = ref_graph_fake(std::atoi(next()));
        else if (a == "--dest")
        {
            std::string s = next();
            size_t pos = 0;
            while (pos < s.size())
            {
                size_t c = s.find(',', pos);
                if (c == std::string::npos) c = s.size();
                dest.push_back(std::strtoull(s.substr(pos, c - pos).c_str(), nullptr, 10));
                pos = c + 1;
            }
        }
        else if (a == "--pick")
        {
            std::string s = next();
            size_t c = s.find(':');
            pick_seed = (uint32_t)std::strtoul(s.substr(0, c).c_str(), nullptr, 10);
            pick_count = std::atoi(s.substr(c + 1).c_str());
        }
        else if (a == "--iters") iters = std::atoi(next());
        else if (a == "--goal-bias") goal_bias = std::atof(next());
        else if (a == "--pseudo") pseudo = 1;
        else if (a == "--mode") mode = next();
        else if (a == "--max-time") max_time = std::atoi(next());
        else if (a == "--max-iter") max_iter = std::atoi(next());
        else if (a == "--ga")
        {
            std::string s = next();
            std::sscanf(s.c_str(), "%d,%d,%d", &ga_mut, &ga_pop, &ga_gen);
        }
        else if (a == "--dump") dump = next();
        else if (a == "--matrix") matrix = next();
        else if (a == "--solves") solves = std::atoi(next());
        else
        {
            std::fprintf(stderr, "unknown argument %s\n", a.c_str());
            return 2;
        }
    }
    if (mode == "rtsp") return rtsp_mode(matrix.c_str(), solves, ga_mut, ga_pop, ga_gen);
    if (!g)
    {
        std::fprintf(stderr, "no graph\n");
        return 2;
    }
    if (pick_count > 0)
    {
        dest.resize(pick_count);
        ref_pick_destinations(g, pick_seed, pick_count, dest.data());
    }
    if (dest.size() < 2)
    {
        std::fprintf(stderr, "need at least source and target\n");
        return 2;
    }
    int K = (int)dest.size();
    uint64_t n = ref_graph_num_nodes(g);
    void* p = ref_planner_create(g, dest.front(), dest.back(), dest.data() + 1, K - 2, goal_bias,
                                 pseudo, max_iter, max_time, 1, 1, ga_mut, ga_pop, ga_gen);
    std::printf("{\"toolchain\": \"%s\", \"n\": %llu, \"arcs\": %llu, \"K\": %d, \"dest\": [",
                ref_toolchain(), (unsigned long long)n, (unsigned long long)ref_graph_num_arcs(g),
                K);
    for (int k = 0; k < K; ++k) std::printf("%s%llu", k ? ", " : "", (unsigned long long)dest[k]);
    std::printf("], \"goal_bias\": %.17g, \"mode\": \"%s\", ", goal_bias, mode.c_str());
    if (mode == "expand")
    {
        auto t0 = std::chrono::steady_clock::now();
        uint64_t expansions = ref_planner_iterate(p, iters);
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        uint64_t total = 0;
        for (int k = 0; k < K; ++k) total += ref_planner_tree_size(p, k);
        std::printf("\"iters\": %d, \"expansions\": %llu, \"seconds\": %.6f, "
                    "\"expansions_per_s\": %.3f, \"tree_vertices\": %llu, "
                    "\"treehash\": \"%016llx\"",
                    iters, (unsigned long long)expansions, dt, expansions / dt,
                    (unsigned long long)total, (unsigned long long)ref_planner_treehash(p));
    }
    else
    {
        std::vector<ref_log_row_t> rows(4096);
        int64_t it = 0, nexp = 0;
        double el = 0;
        int nr = ref_planner_run(p, rows.data(), (int)rows.size(), &it, &el, &nexp);
        uint64_t total = 0;
        for (int k = 0; k < K; ++k) total += ref_planner_tree_size(p, k);
        std::printf("\"iterations\": %lld, \"expansions\": %lld, \"seconds\": %.6f, "
                    "\"tree_vertices\": %llu, \"final_cost\": %.17g, \"treehash\": \"%016llx\", "
                    "\"rows\": [",
                    (long long)it, (long long)nexp, el, (unsigned long long)total,
                    ref_planner_path_cost(p), (unsigned long long)ref_planner_treehash(p));
        for (int r = 0; r < nr; ++r)
        {
            std::printf("%s[%.6f, %.17g, %lld, %lld]", r ? ", " : "", rows[r].time_s,
                        rows[r].path_cost, (long long)rows[r].tree_size,
                        (long long)rows[r].iteration);
        }
        std::printf("], \"path\": [");
        std::vector<uint64_t> path(1 << 22);
        uint64_t pl = ref_planner_path(p, path.data(), path.size());
        for (uint64_t i = 0; i < pl && i < path.size(); ++i)
        {
            std::printf("%s%llu", i ? ", " : "", (unsigned long long)path[i]);
        }
        std::printf("]");
    }
    std::printf("}\n");
    if (!dump.empty())
    {
        FILE* f = std::fopen(dump.c_str(), "wb");
        std::vector<uint32_t> par(n);
        std::vector<double> cost(n);
        uint64_t hdr[2] = {n, (uint64_t)K};
        std::fwrite(hdr, 8, 2, f);
        for (int k = 0; k < K; ++k)
        {
            ref_planner_get_tree(p, k, par.data(), cost.data());
            std::fwrite(par.data(), 4, n, f);
            std::fwrite(cost.data(), 8, n, f);
        }
        std::vector<double> D(K * K);
        std::vector<uint64_t> cn(K * K);
        ref_planner_get_matrices(p, D.data(), cn.data(), nullptr, nullptr, nullptr);
        std::fwrite(D.data(), 8, K * K, f);
        std::fwrite(cn.data(), 8, K * K, f);
        std::fclose(f);
    }
    std::fflush(stdout);
    _exit(0);   // skip destructors (the reference's ~CSVFile aborts)
}
