// host/src/host_api.h -- C entry points of libimomd_host.so used by the Python harness
// (tests, bench) to reach the C++ host facade.  The facade itself is the C++ class in
// host/include/imomd_rrt_star/imomd_rrt_star.h.
#ifndef IMOMD_HOST_API_H
#define IMOMD_HOST_API_H

#include <stdint.h>

namespace imomd_host {
void selection_probabilities(int K, const double* lat, const double* lon, double* P);
}

extern "C" {
/* K x K row-major selection-probability matrix for destinations at (lat[i], lon[i]) */
void imomd_host_selection_probabilities(int K, const double* lat, const double* lon, double* P);
double imomd_host_great_circle_m(double lat_a, double lon_a, double lat_b, double lon_b);

/* One complete planner run through the C++ facade (ImomdRRT constructor + findShortestPath
 * on a pthread, exactly like the reference's main.cpp:219-229).  The road graph is handed
 * over as an edge list and rebuilt into the reference's two containers with the parser's
 * two inserts per edge.  Results: log rows (time, cost, tree size, iteration), the final
 * vertex path, the visiting sequence.  Returns 0, or -1 with a message in err. */
typedef struct imomd_host_row
{
    double time_s;
    double path_cost;
    long long tree_size;
    long long iteration;
} imomd_host_row;

int imomd_host_run(int device, uint64_t n, const double* lat, const double* lon, uint64_t m,
                   const uint64_t* eu, const uint64_t* ev, const double* ew, int K, const uint64_t* dest,
                   double goal_bias, int pseudo_mode, int max_iter, int max_time, int ga_mutation_iter, int ga_population,
                   int ga_generation, imomd_host_row* rows, int row_cap, int* n_rows, uint64_t* path,
                   uint64_t path_cap, uint64_t* path_len, int32_t* seq, int seq_cap, int* seq_len,
                   double* final_cost, long long* iterations, long long* expansions, char* err,
                   int errlen);

/* the facade driven stepwise: passes x expandTreeLayers_, optionally one solveRTSP_ (in pseudo mode
 * that is where the tree merge happens); _native returns the imomd_planner* for state readback */
void* imomd_host_planner_create(int device, uint64_t n, const double* lat, const double* lon, uint64_t m,
                                const uint64_t* eu, const uint64_t* ev, const double* ew, int K,
                                const uint64_t* dest, double goal_bias, int pseudo_mode,
                                int ga_mutation_iter, int ga_population, int ga_generation, char* err,
                                int errlen);
void imomd_host_planner_destroy(void* h);
int imomd_host_planner_step(void* h, int passes, int solve_rtsp, char* err, int errlen);
void* imomd_host_planner_native(void* h);
double imomd_host_planner_cost(void* h);

/* GraphCache (imomd_rrt_star/graph_cache.h) round trip: save + load + order check; CSR of the
 * loaded containers out */
int imomd_host_cache_roundtrip(const char* path, uint64_t n, const double* lat, const double* lon, uint64_t m,
                               const uint64_t* eu, const uint64_t* ev, const double* ew, uint32_t* row_ptr,
                               uint32_t* col, double* w, double* lat_out, double* lon_out, char* err,
                               int errlen);

int imomd_host_cache_load_check(const char* path, char* err, int errlen);

/* EciGenSolver of the facade on a row-major K x K matrix (GA fitness on the GPU) */
void* imomd_host_solver_create(int device, int swapping, int genetic, int ga_mutation_iter,
                               int ga_population, int ga_generation);
void imomd_host_solver_destroy(void* s);
int imomd_host_solver_solve(void* s, int K, const double* D, int source, int target, double* cost,
                            int32_t* seq, int cap, uint64_t* evaluated, char* err, int errlen);
/* GA islands: begin (insertion + seeding with mt19937{seed}), one generation, best tour, migrant */
int imomd_host_island_begin(void* s, int K, const double* D, int source, int target, unsigned seed,
                            char* err, int errlen);
int imomd_host_island_step(void* s, char* err, int errlen);   /* 1 = stepped, 0 = population < 2 */
int imomd_host_island_best(void* s, double* cost, int32_t* seq, int cap);
void imomd_host_island_accept(void* s, double cost, const int32_t* seq, int len);
uint64_t imomd_host_island_evaluated(void* s);
}
#endif
