// host/include/imomd_rrt_star/eci_gen_tsp_solver.h -- ECI-Gen relaxed-TSP solver of the
// host facade.  Same public API as the reference's EciGenSolver
// (include/imomd_rrt_star/eci_gen_tsp_solver.h:63-74 upstream): constructed from
// rtsp_setting_t, `solveRTSP(matrix, source, target)` returns {cost, visiting sequence}.
//
// What differs is where the work runs: Dijkstra, the enhanced cheapest insertion and the
// duplicate-removal pass stay on the host, while every fitness
// evaluation of the genetic stage (calculatePathCost_, eci_gen_tsp_solver.cpp:442-450)
// is batched onto the GPU through imomd_ga_eval: candidates of a whole seeding pass /
// generation are produced first with exactly the reference's random-number draws, ranked
// on the device in one launch, then filtered in their original order.
#ifndef IMOMD_B200_ECI_GEN_TSP_H
#define IMOMD_B200_ECI_GEN_TSP_H

#include <cstdint>
#include <memory>
#include <random>
#include <utility>
#include <vector>

#include "setting/rtsp_setting_t.h"

struct imomd_ctx;

class EciGenSolver
{
public:
    EciGenSolver() = default;
    explicit EciGenSolver(const rtsp_setting_t& setting, int device = 0);
    ~EciGenSolver();
    EciGenSolver(const EciGenSolver&) = delete;
    EciGenSolver& operator=(const EciGenSolver&) = delete;
    EciGenSolver(EciGenSolver&& other) noexcept;
    EciGenSolver& operator=(EciGenSolver&& other) noexcept;

    std::pair<double, std::shared_ptr<std::vector<int>>> solveRTSP(
        std::shared_ptr<std::vector<std::vector<double>>> adjacency_matrix_ptr, int source_id,
        int target_id);

    // shortest chain source -> target only (used by the reference's baseline planners)
    std::pair<double, std::shared_ptr<std::vector<int>>> solveDijkstra(
        std::shared_ptr<std::vector<std::vector<double>>> adjacency_matrix_ptr, int source_id,
        int target_id);

    // ---- GA islands (BASELINE config 4; not in the reference, which has one population) ----
    // Several solvers ("islands", one per GPU) evolve their own population from their own seed
    // and exchange their best tour between generations; the exchange itself is the caller's
    // (imomd_b200/sharding.py: one NCCL all-gather of 272 bytes per rank and generation).
    // islandBegin = insertion + seeding pass with mt19937{seed}; seed 0 reproduces upstream.
    void islandBegin(std::shared_ptr<std::vector<std::vector<double>>> adjacency_matrix_ptr,
                     int source_id, int target_id, unsigned seed);
    bool islandStep();                                   // one generation; false: population < 2
    std::pair<double, std::vector<int>> islandBest() const { return {best_cost_, best_}; }
    // a migrant replaces the least fit chromosome when it is fitter; keeps the incumbent
    void islandAccept(const std::vector<int>& tour, double cost);
    size_t islandPopulation() const { return pool_.size(); }

    // number of tours the GPU ranked during the last solve (for tests / reports)
    size_t lastEvaluatedTours() const { return evaluated_; }

private:
    using Tour = std::vector<int>;

    rtsp_setting_t setting_{};
    std::mt19937 rng_{0};
    imomd_ctx* ctx_ = nullptr;
    int device_ = 0;

    int K_ = 0;
    std::vector<double> D_;     // row-major K x K
    int source_ = 0, target_ = 0;
    double best_cost_ = 0;
    Tour best_;
    std::vector<Tour> pool_;    // chromosomes
    std::vector<double> fitness_;
    size_t evaluated_ = 0;

    double d(int i, int j) const { return D_[static_cast<size_t>(i) * K_ + j]; }
    double shortestChain(Tour& chain) const;
    void cheapestInsertion();
    void dropRepeats(Tour& tour) const;
    void evaluate(const std::vector<Tour>& tours, std::vector<double>& costs);
    void seedPopulation();
    void evolve();
    bool evolveOnce();
    int drawInt(int lo, int hi);
    double drawReal(double lo, double hi);
};

#endif
