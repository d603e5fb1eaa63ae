// host/src/host_capi.cpp -- C entry points over the C++ facade for the Python harness.
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include <pthread.h>

#include "host_api.h"
#include "imomd_rrt_star/graph_cache.h"
#include "imomd_rrt_star/imomd_rrt_star.h"

namespace {

void put_err(char* err, int errlen, const std::string& msg)
{
    if (err && errlen > 0)
    {
        std::strncpy(err, msg.c_str(), static_cast<size_t>(errlen) - 1);
        err[errlen - 1] = 0;
    }
}

}  // namespace

extern "C" int imomd_host_run(int device, uint64_t n, const double* lat, const double* lon, uint64_t m,
                              const uint64_t* eu, const uint64_t* ev, const double* ew, int K,
                              const uint64_t* dest, double goal_bias, int pseudo_mode, int max_iter,
                              int max_time,
                              int ga_mutation_iter, int ga_population, int ga_generation,
                              imomd_host_row* rows, int row_cap, int* n_rows, uint64_t* path,
                              uint64_t path_cap, uint64_t* path_len, int32_t* seq, int seq_cap,
                              int* seq_len, double* final_cost, long long* iterations,
                              long long* expansions, char* err, int errlen)
{
    try
    {
        auto map = std::make_shared<std::vector<location_t>>(n);
        auto conn = std::make_shared<std::vector<std::unordered_map<size_t, double>>>(n);
        for (uint64_t i = 0; i < n; ++i) (*map)[i] = location_t(i, lat[i], lon[i]);
        for (uint64_t e = 0; e < m; ++e)
        {
            (*conn)[eu[e]].insert({ev[e], ew[e]});
            (*conn)[ev[e]].insert({eu[e], ew[e]});
        }
        imomd_setting_t st{};
        st.goal_bias = goal_bias;
        st.max_iter = max_iter;
        st.max_time = max_time;
        st.random_seed = false;
        st.pseudo_mode = pseudo_mode != 0;
        st.log_data = 0;
        st.rtsp_setting.swapping = true;
        st.rtsp_setting.genetic = true;
        st.rtsp_setting.ga_mutation_iter = ga_mutation_iter;
        st.rtsp_setting.ga_population = ga_population;
        st.rtsp_setting.ga_generation = ga_generation;
        st.rtsp_setting.ga_random_seed = false;
        std::vector<size_t> objectives(dest + 1, dest + K - 1);
        ImomdRRT::setDefaultDevice(device);
        ImomdRRT planner(dest[0], dest[K - 1], objectives, map, conn, st);
        pthread_t tid;
        pthread_create(&tid, NULL, ImomdRRT::findShortestPath, &planner);
        pthread_join(tid, NULL);
        planner.rethrowFailure();
        const auto& log = planner.getLog();
        *n_rows = static_cast<int>(log.size());
        for (int i = 0; i < *n_rows && i < row_cap; ++i)
            rows[i] = {log[i].time_s, log[i].path_cost, log[i].tree_size, log[i].iteration};
        std::vector<size_t> p = planner.getShortestPath();
        *path_len = p.size();
        for (uint64_t i = 0; i < p.size() && i < path_cap; ++i) path[i] = p[i];
        std::vector<int> s = planner.getVisitingSequence();
        *seq_len = static_cast<int>(s.size());
        for (int i = 0; i < *seq_len && i < seq_cap; ++i) seq[i] = s[i];
        *final_cost = planner.getShortestPathCost();
        *iterations = planner.iterations();
        *expansions = planner.expansions();
        return 0;
    }
    catch (const std::exception& ex)
    {
        put_err(err, errlen, ex.what());
        return -1;
    }
}

// stepwise facade handle (tests: state after the pseudo-tree merge against the reference's)
extern "C" void* imomd_host_planner_create(int device, uint64_t n, const double* lat, const double* lon,
                                           uint64_t m, const uint64_t* eu, const uint64_t* ev,
                                           const double* ew, int K, const uint64_t* dest, double goal_bias,
                                           int pseudo_mode, int ga_mutation_iter, int ga_population,
                                           int ga_generation, char* err, int errlen)
{
    try
    {
        auto map = std::make_shared<std::vector<location_t>>(n);
        auto conn = std::make_shared<std::vector<std::unordered_map<size_t, double>>>(n);
        for (uint64_t i = 0; i < n; ++i) (*map)[i] = location_t(i, lat[i], lon[i]);
        for (uint64_t e = 0; e < m; ++e)
        {
            (*conn)[eu[e]].insert({ev[e], ew[e]});
            (*conn)[ev[e]].insert({eu[e], ew[e]});
        }
        imomd_setting_t st{};
        st.goal_bias = goal_bias;
        st.max_iter = 1 << 30;
        st.max_time = 1 << 30;
        st.random_seed = false;
        st.pseudo_mode = pseudo_mode != 0;
        st.log_data = 0;
        st.rtsp_setting.swapping = true;
        st.rtsp_setting.genetic = true;
        st.rtsp_setting.ga_mutation_iter = ga_mutation_iter;
        st.rtsp_setting.ga_population = ga_population;
        st.rtsp_setting.ga_generation = ga_generation;
        st.rtsp_setting.ga_random_seed = false;
        std::vector<size_t> objectives(dest + 1, dest + K - 1);
        ImomdRRT::setDefaultDevice(device);
        return new ImomdRRT(dest[0], dest[K - 1], objectives, map, conn, st);
    }
    catch (const std::exception& ex)
    {
        put_err(err, errlen, ex.what());
        return nullptr;
    }
}

extern "C" void imomd_host_planner_destroy(void* h) { delete static_cast<ImomdRRT*>(h); }

extern "C" int imomd_host_planner_step(void* h, int passes, int solve_rtsp, char* err, int errlen)
{
    try
    {
        ImomdRRT* p = static_cast<ImomdRRT*>(h);
        if (passes > 0) p->stepPasses(passes);
        if (solve_rtsp) p->stepRTSP();
        return 0;
    }
    catch (const std::exception& ex)
    {
        put_err(err, errlen, ex.what());
        return -1;
    }
}

extern "C" void* imomd_host_planner_native(void* h) { return static_cast<ImomdRRT*>(h)->nativeHandle(); }

extern "C" double imomd_host_planner_cost(void* h) { return static_cast<ImomdRRT*>(h)->getShortestPathCost(); }

extern "C" void* imomd_host_solver_create(int device, int swapping, int genetic, int ga_mutation_iter,
                                          int ga_population, int ga_generation)
{
    rtsp_setting_t st{};
    st.swapping = swapping != 0;
    st.genetic = genetic != 0;
    st.ga_mutation_iter = ga_mutation_iter;
    st.ga_population = ga_population;
    st.ga_generation = ga_generation;
    st.ga_random_seed = false;
    return new EciGenSolver(st, device);
}

extern "C" void imomd_host_solver_destroy(void* s) { delete static_cast<EciGenSolver*>(s); }

extern "C" int imomd_host_solver_solve(void* s, int K, const double* D, int source, int target,
                                       double* cost, int32_t* seq, int cap, uint64_t* evaluated,
                                       char* err, int errlen)
{
    try
    {
        auto m = std::make_shared<std::vector<std::vector<double>>>(K, std::vector<double>(K));
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) (*m)[i][j] = D[static_cast<size_t>(i) * K + j];
        EciGenSolver* solver = static_cast<EciGenSolver*>(s);
        auto res = solver->solveRTSP(m, source, target);
        *cost = res.first;
        int n = static_cast<int>(res.second->size());
        for (int i = 0; i < n && i < cap; ++i) seq[i] = (*res.second)[i];
        if (evaluated) *evaluated = solver->lastEvaluatedTours();
        return n;
    }
    catch (const std::exception& ex)
    {
        put_err(err, errlen, ex.what());
        return -1;
    }
}

extern "C" int imomd_host_island_begin(void* s, int K, const double* D, int source, int target,
                                       unsigned seed, char* err, int errlen)
{
    try
    {
        auto m = std::make_shared<std::vector<std::vector<double>>>(K, std::vector<double>(K));
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) (*m)[i][j] = D[static_cast<size_t>(i) * K + j];
        static_cast<EciGenSolver*>(s)->islandBegin(m, source, target, seed);
        return 0;
    }
    catch (const std::exception& ex)
    {
        put_err(err, errlen, ex.what());
        return -1;
    }
}

extern "C" int imomd_host_island_step(void* s, char* err, int errlen)
{
    try
    {
        return static_cast<EciGenSolver*>(s)->islandStep() ? 1 : 0;
    }
    catch (const std::exception& ex)
    {
        put_err(err, errlen, ex.what());
        return -1;
    }
}

extern "C" int imomd_host_island_best(void* s, double* cost, int32_t* seq, int cap)
{
    auto best = static_cast<EciGenSolver*>(s)->islandBest();
    *cost = best.first;
    int n = static_cast<int>(best.second.size());
    for (int i = 0; i < n && i < cap; ++i) seq[i] = best.second[i];
    return n;
}

extern "C" void imomd_host_island_accept(void* s, double cost, const int32_t* seq, int len)
{
    static_cast<EciGenSolver*>(s)->islandAccept(std::vector<int>(seq, seq + len), cost);
}

extern "C" uint64_t imomd_host_island_evaluated(void* s)
{
    return static_cast<EciGenSolver*>(s)->lastEvaluatedTours();
}

// GraphCache round trip for the tests: containers built from the edge list (insertion order =
// list order, as every other wrapper here), saved, loaded; the loaded containers' iteration
// order is written out as CSR.  Returns 0, or -1 with a message.
extern "C" int imomd_host_cache_roundtrip(const char* path, uint64_t n, const double* lat, const double* lon,
                                          uint64_t m, const uint64_t* eu, const uint64_t* ev,
                                          const double* ew, uint32_t* row_ptr, uint32_t* col, double* w,
                                          double* lat_out, double* lon_out, char* err, int errlen)
{
    try
    {
        imomd_b200::GraphCache::Map map(n);
        imomd_b200::GraphCache::Connection conn(n);
        for (uint64_t i = 0; i < n; ++i) map[i] = location_t(i, lat[i], lon[i]);
        for (uint64_t e = 0; e < m; ++e)
        {
            conn[eu[e]].insert({ev[e], ew[e]});
            conn[ev[e]].insert({eu[e], ew[e]});
        }
        imomd_b200::GraphCache::save(path, map, conn);
        std::shared_ptr<imomd_b200::GraphCache::Map> m2;
        std::shared_ptr<imomd_b200::GraphCache::Connection> c2;
        if (!imomd_b200::GraphCache::load(path, m2, c2)) throw std::runtime_error("cache file vanished");
        if (m2->size() != n || c2->size() != n) throw std::runtime_error("size mismatch after load");
        uint32_t at = 0;
        for (uint64_t v = 0; v < n; ++v)
        {
            lat_out[v] = (*m2)[v].latitude;
            lon_out[v] = (*m2)[v].longitude;
            row_ptr[v] = at;
            // must equal the ORIGINAL containers' order, element by element
            auto it = conn[v].begin();
            for (const auto& kv : (*c2)[v])
            {
                if (it == conn[v].end() || it->first != kv.first || it->second != kv.second)
                    throw std::runtime_error("iteration order differs after load");
                ++it;
                col[at] = static_cast<uint32_t>(kv.first);
                w[at] = kv.second;
                ++at;
            }
            if (it != conn[v].end()) throw std::runtime_error("row shorter after load");
        }
        row_ptr[n] = at;
        return 0;
    }
    catch (const std::exception& ex)
    {
        put_err(err, errlen, ex.what());
        return -1;
    }
}

// 0 = loaded and verified, 1 = no such file, -1 = rejected (message in err)
extern "C" int imomd_host_cache_load_check(const char* path, char* err, int errlen)
{
    try
    {
        std::shared_ptr<imomd_b200::GraphCache::Map> m;
        std::shared_ptr<imomd_b200::GraphCache::Connection> c;
        return imomd_b200::GraphCache::load(path, m, c) ? 0 : 1;
    }
    catch (const std::exception& ex)
    {
        put_err(err, errlen, ex.what());
        return -1;
    }
}
