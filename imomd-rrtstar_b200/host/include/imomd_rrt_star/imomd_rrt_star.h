// host/include/imomd_rrt_star/imomd_rrt_star.h -- the planner the application links
// against.  Drop-in for the reference's library API
// (include/imomd_rrt_star/imomd_rrt_star.h:126-138 upstream): same constructor signature,
// same pthread entry points `findShortestPath` / `printPath`, same `getDistanceMatrix()`,
// same road-graph containers and imomd_setting_t going in, same stdout / CSV rows coming
// out.  The multi-tree expansion itself runs on the GPU behind the C ABI of
// include/imomd_b200.h; this class only keeps what the reference keeps on the host side
// of that line: the RTSP trigger every 100 passes, path extraction and logging.
//
// pseudo_mode: the trees rooted at pseudo destinations are merged into the source / target
// trees once the graph of trees is connected (mergePseudoTrees_, :744-916 upstream).  The merge
// is driven from here -- it iterates std containers whose order is part of the reference's
// behaviour -- and performs every tree update on the device (imomd_attach).
#ifndef IMOMD_B200_IMOMD_RRT_STAR_H
#define IMOMD_B200_IMOMD_RRT_STAR_H

#include <chrono>
#include <cstddef>
#include <fstream>
#include <memory>
#include <random>
#include <string>
#include <exception>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <pthread.h>

#include "data/location_t.h"
#include "setting/imomd_setting_t.h"

#include "imomd_rrt_star/eci_gen_tsp_solver.h"

struct imomd_graph;
struct imomd_planner;

class ImomdRRT
{
public:
    struct LogRow
    {
        double time_s;
        double path_cost;
        long long tree_size;
        long long iteration;
    };

    ImomdRRT() = default;
    ImomdRRT(size_t source, size_t target, std::vector<size_t> objectives,
             const std::shared_ptr<std::vector<location_t>> map,
             const std::shared_ptr<std::vector<std::unordered_map<size_t, double>>> connection,
             const imomd_setting_t& setting);
    ~ImomdRRT();
    ImomdRRT(const ImomdRRT&) = delete;
    ImomdRRT& operator=(const ImomdRRT&) = delete;

    // pthread entry points; both end in pthread_exit like the reference's
    static void* findShortestPath(void* arg);
    static void* printPath(void* arg);

    std::shared_ptr<std::vector<std::vector<double>>> getDistanceMatrix();

    // additions (the reference exposes its results only through stdout and the CSV file)
    std::vector<size_t> getShortestPath();
    double getShortestPathCost();
    std::vector<int> getVisitingSequence();
    const std::vector<LogRow>& getLog() const { return log_; }
    long long iterations() const { return iteration_count_; }
    long long expansions() const { return expansions_; }
    // findShortestPath runs on a pthread and cannot throw: a failure (device capacity status,
    // unmergeable pseudo tree) ends the run, is printed to stderr and kept here
    bool failed() const { return static_cast<bool>(failure_); }
    void rethrowFailure() const { if (failure_) std::rethrow_exception(failure_); }
    // stepwise driving (what findShortestPath's loop does): n x expandTreeLayers_, one
    // solveRTSP_ ; the C-ABI planner handle for state readback
    void stepPasses(int passes) { expandBatch(passes); }
    void stepRTSP() { solveRTSP(); }
    imomd_planner* nativeHandle() { return planner_; }
    // device of the planners constructed afterwards (graph, trees and solver live there)
    static void setDefaultDevice(int device);

private:
    void run();                 // body of findShortestPath
    bool expandBatch(int passes);
    void solveRTSP();
    void updatePath();
    void logData();
    void feedWords(size_t n);
    void drainConnectionLog();
    void mergePseudoTrees();
    void mergeTwoTrees(int parent_tree, int child_tree);
    void attachList(int tree, const std::vector<uint32_t>& vertices);
    int setIndex(int a, int b) const { return a > b ? a * (a - 1) / 2 + b : b * (b - 1) / 2 + a; }

    imomd_setting_t setting_{};
    std::shared_ptr<std::vector<location_t>> map_;
    std::shared_ptr<std::vector<std::unordered_map<size_t, double>>> connection_;
    std::vector<size_t> destinations_;
    int K_ = 0;
    int device_ = 0;

    imomd_graph* graph_ = nullptr;
    imomd_planner* planner_ = nullptr;
    std::mt19937 rng_{0};
    size_t words_fed_ = 0;
    size_t words_consumed_ = 0;

    long long iteration_count_ = 0;
    long long expansions_ = 0;
    long long tree_vertices_ = 0;
    int active_trees_ = 1;
    bool connected_ = false;
    bool dm_updated_ = false;

    EciGenSolver solver_;
    std::shared_ptr<std::vector<int>> sequence_;
    double shortest_path_cost_ = 0;
    std::vector<size_t> shortest_path_;

    bool path_updated_ = false;
    bool finished_ = false;
    pthread_cond_t cond_ = PTHREAD_COND_INITIALIZER;
    pthread_mutex_t lock_ = PTHREAD_MUTEX_INITIALIZER;

    // pseudo mode: host mirrors of connection_nodes_set_ (real unordered_sets fed in the
    // device's insertion order, so that they iterate like the reference's)
    std::vector<std::unordered_set<size_t>> connection_sets_;
    uint32_t log_consumed_ = 0;
    bool merge_done_ = false;
    std::exception_ptr failure_;

    std::chrono::steady_clock::time_point start_;
    std::ofstream csv_;
    std::vector<LogRow> log_;
};

#endif
