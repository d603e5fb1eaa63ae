/* oracle/imomd_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C API of the CPU restatement of the reference's multi-tree expansion hot path
 * and ECI-Gen solver (oracle/imomd_oracle.cpp).  The restatement is container
 * free (dense arrays, CSR, explicit child-order rule) so that its results do not
 * depend on libstdc++'s hash-container iteration order; it is pinned against
 * the unmodified reference (oracle/_ref) by tests/test_oracle_vs_ref.py and by
 * the golden fixtures under tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this library.  The product (imomd-rrtstar_b200/) never links or calls it.
 */
#ifndef IMOMD_ORACLE_H
#define IMOMD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_graph orc_graph;
typedef struct orc_planner orc_planner;
typedef struct orc_solver orc_solver;

typedef struct orc_trace
{
    int32_t tree;
    int32_t skipped;
    uint64_t rand_id;
    double rand_lat;
    double rand_lon;
    uint64_t x_nearest;
    double nearest_dist;
    uint64_t x_new;
    uint64_t parent_of_new;
    double cost_of_new;
    uint64_t tree_size_after;
    uint64_t frontier_after;
} orc_trace;

/* geometry (include/osm_converter/compute_haversine.hpp:41-58, compute_bearing.hpp:45-61) */
double orc_haversine(double lat1, double lon1, double lat2, double lon2);
double orc_bearing(double lat1, double lon1, double lat2, double lon2);

/* libstdc++ random number plumbing restated */
void orc_mt19937_words(uint32_t seed, uint64_t n, uint32_t* out);
void orc_uniform_real(uint32_t seed, double a, double b, uint64_t n, double* out);
void orc_uniform_size_t(uint32_t seed, uint64_t lo, uint64_t hi, uint64_t n, uint64_t* out);
void orc_uniform_int(uint32_t seed, int lo, int hi, uint64_t n, int32_t* out);

/* road graph: CSR in the reference containers' native iteration order */
orc_graph* orc_graph_create(uint32_t n, const uint32_t* row_ptr, const uint32_t* col,
                            const double* w, const double* lat, const double* lon);
void orc_graph_destroy(orc_graph* g);
/* adjacency iteration order libstdc++ gives for `unordered_map<size_t,double>` rows
 * filled by the two inserts of src/osm_parser.cpp:179-180 per edge, in edge order */
void orc_native_csr_from_edges(uint32_t n, uint64_t m, const uint64_t* eu, const uint64_t* ev,
                               const double* ew, uint32_t* row_ptr, uint32_t* col, double* w);

/* planner: dest = {source, objectives..., target} (src/imomd_rrt_star.cpp:50-55) */
orc_planner* orc_planner_create(orc_graph* g, int K, const uint64_t* dest, double goal_bias,
                                int pseudo_mode);
void orc_planner_destroy(orc_planner* p);
uint64_t orc_planner_iterate(orc_planner* p, int iters);
/* batched mode (imomd_expand_batch): `steps` steps of B expansions per tree */
uint64_t orc_planner_iterate_batched(orc_planner* p, int steps, int B);
int orc_planner_expand_traced(orc_planner* p, int k, orc_trace* tr);
void orc_planner_get_tree(orc_planner* p, int k, uint32_t* parent, double* cost);
uint64_t orc_planner_tree_size(orc_planner* p, int k);
int orc_planner_children(orc_planner* p, int k, uint64_t v, uint64_t* out, int cap);
uint64_t orc_planner_treehash(orc_planner* p);
void orc_planner_get_matrices(orc_planner* p, double* D, uint64_t* conn_node, double* prob,
                              uint64_t* mh_id, double* mh_val);
void orc_planner_get_flags(orc_planner* p, int32_t* is_done, int32_t* connected,
                           int32_t* dm_updated, int32_t* ds_parent);
uint64_t orc_planner_frontier(orc_planner* p, int k, uint64_t* out, uint64_t cap);
int orc_planner_nearest(orc_planner* p, int k, double lat, double lon, uint64_t* idx,
                        double* dist, int32_t* n_ties, double* second);
/* number of exact distance ties seen by the arg-min loops so far */
uint64_t orc_planner_tie_count(orc_planner* p);
/* pseudo mode: ordered log of (set index, vertex) insertions into the connection-node sets */
uint64_t orc_planner_connection_log(orc_planner* p, uint32_t* set_idx, uint64_t* vertex,
                                    uint64_t cap);

/* ECI-Gen (src/eci_gen_tsp_solver.cpp) */
orc_solver* orc_solver_create(int swapping, int genetic, int ga_mutation_iter, int ga_population,
                              int ga_generation);
void orc_solver_destroy(orc_solver* s);
int orc_solver_solve(orc_solver* s, int K, const double* D, int source, int target, double* cost,
                     int32_t* seq, int cap);
void orc_path_costs(int K, const double* D, int n_tours, int stride, const int32_t* tours,
                    const int32_t* lens, double* out);

#ifdef __cplusplus
}
#endif
#endif /* IMOMD_ORACLE_H */
