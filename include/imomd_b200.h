/* include/imomd_b200.h -- the C ABI of the B200-native IMOMD-RRT* expansion path.
 *
 * This is the only boundary between the C++ host (the reference's library API,
 * imomd-rrtstar_b200/host/) and CUDA (imomd-rrtstar_b200/csrc/ -> libimomd_b200.so).
 * The reference has no FFI of its own: its API is the C++
 * class ImomdRRT.  Each entry point below therefore cites the reference member
 * whose work it takes over (paths under the reference checkout).
 *
 * Conventions: every function returns an int status (IMOMD_OK == 0), nothing
 * throws, no C++ or torch types cross the line, the caller owns every host buffer
 * and the library owns all device memory.  One handle lives on one device and one
 * stream; a handle is not thread-safe, distinct handles may be used from distinct
 * host threads.  imomd_last_error() returns a thread-local message for the last
 * non-zero status.  There is NO CPU fallback: without a CUDA device every call
 * that needs one fails with IMOMD_ERR_CUDA.
 */
#ifndef IMOMD_B200_H
#define IMOMD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMOMD_KMAX 32u              /* trees per query: source + target + <= 30 objectives */
#define IMOMD_NONE 0xFFFFFFFFu      /* "no vertex" / "not visited" */
#define IMOMD_MAX_DEGREE 13u        /* children order is defined for <= 13 */

enum imomd_status
{
    IMOMD_OK = 0,
    IMOMD_ERR_ARG = 1,          /* bad argument */
    IMOMD_ERR_CUDA = 2,         /* CUDA runtime error or no device */
    IMOMD_ERR_GRAPH = 3,        /* asymmetric adjacency, self loop, degree > 13, ... */
    IMOMD_ERR_NOMEM = 4
};

/* per-query run status written by the device (imomd_query_report.status) */
enum imomd_run_status
{
    IMOMD_RUN_OK = 0,
    IMOMD_RUN_NEED_WORDS = 1,   /* RNG word stream exhausted: feed more, call again */
    IMOMD_RUN_FRONTIER_FULL = 2,/* frontier chunk pool exhausted (raise frontier_chunks) */
    IMOMD_RUN_ARENA_FULL = 3,   /* rewire scratch exhausted (raise arena_entries) */
    IMOMD_RUN_CDF_OVERRUN = 4,  /* goal-bias CDF walk ran past the row (UB upstream) */
    IMOMD_RUN_BAD_STATE = 5
};

typedef struct imomd_graph imomd_graph;       /* road graph resident in HBM */
typedef struct imomd_planner imomd_planner;   /* Q independent queries (K trees each) on one graph */
typedef struct imomd_ctx imomd_ctx;           /* device context for the GA evaluation */

/* Mirrors imomd_setting_t (include/setting/imomd_setting_t.h:39-48) for the fields
 * the expansion path reads, plus sizing knobs of the device data structures. */
typedef struct imomd_params
{
    double goal_bias;           /* setting.goal_bias */
    int32_t pseudo_mode;        /* setting.pseudo_mode */
    int32_t max_resident_queries; /* 0: every query keeps its trees in HBM (resident planner).
                                   * n > 0: at most n sets of per-query state; with more queries than
                                   * that the planner STREAMS: imomd_expand runs every query from its
                                   * constructed state to `iters` passes, queries take a state slot from
                                   * a device work queue as slots fall free, and only the K x K state,
                                   * the report and the tree hash of a query outlive its run.
                                   * -1: as many slots as run concurrently and fit in device memory. */
    uint32_t frontier_chunks;   /* 32-record chunks per tree, 0 = auto */
    uint32_t arena_entries;     /* rewire scratch (uint32 entries) per query, 0 = auto */
    uint32_t log_entries;       /* pseudo-mode connection log entries per query, 0 = auto */
    uint32_t reserved;
    double tie_band;            /* relative band inside which a comparison of haversine values is
                                 * decided with the host's libm instead of CUDA's; 0 = 1e-12 (tests
                                 * widen it to exercise the service) */
} imomd_params;

/* what one imomd_expand call did for one query */
typedef struct imomd_query_report
{
    int32_t status;             /* imomd_run_status */
    int32_t active_trees;       /* trees that can still expand */
    int64_t iterations;         /* expandTreeLayers_ passes executed by this call */
    int64_t expansions;         /* non-early-returning expandTree_ calls, this call */
    int64_t tree_vertices;      /* sum over trees of visited vertices (logData_ tree_size) */
    int64_t words_consumed;     /* absolute cursor into the injected word stream */
    int32_t connected;          /* is_connected_graph_ */
    int32_t dm_updated;         /* is_distance_matrix_updated_ */
    int64_t near_ties;          /* arg-min decisions certified on the slow exact path */
    int64_t unresolved_ties;    /* exact ties on the reference's own values, or near-ties the host
                                 * could not be asked about (lowest id / device choice taken) */
    int64_t log_entries;        /* pseudo-mode connection log length */
    int64_t certified_ties;     /* near-ties decided on the host's glibc values (cumulative) */
    uint64_t tree_hash;         /* streaming planners: TREEHASH of the finished query (imomd_get_treehash) */
    uint64_t busy_cycles;       /* SM-clock cycles the query occupied its CTA in this call */
} imomd_query_report;

typedef struct imomd_nn_query
{
    int32_t query;              /* which query of the planner */
    int32_t tree;               /* which tree of that query */
    double lat;
    double lon;
} imomd_nn_query;

/* ------------------------------------------------------------------ graph ---
 * Replaces the two shared_ptr containers the reference planner is constructed
 * with (include/imomd_rrt_star/imomd_rrt_star.h:130-133, :147-149): uploaded once
 * as CSR in the containers' native iteration order + SoA coordinates.  Validates
 * symmetry, absence of self loops and max degree <= 13. */
int imomd_graph_create(int device, uint32_t n, const uint32_t* row_ptr, const uint32_t* col,
                       const double* w, const double* lat, const double* lon,
                       imomd_graph** out);
/* same, with the number of hash cells per axis forced (0 = automatic) */
int imomd_graph_create_ex(int device, uint32_t n, const uint32_t* row_ptr, const uint32_t* col,
                          const double* w, const double* lat, const double* lon, int grid_dim,
                          imomd_graph** out);
int imomd_graph_destroy(imomd_graph* g);
int imomd_graph_info(const imomd_graph* g, uint32_t* n, uint32_t* arcs, int32_t* grid_dim,
                     int32_t* device);

/* ---------------------------------------------------------------- planner ---
 * ImomdRRT::ImomdRRT (src/imomd_rrt_star.cpp:39-223): `dest` holds, per query,
 * {source, objectives..., target}; `prob` the K x K selection-probability matrix
 * the host facade computes with the reference's formula (:84-185, glibc math),
 * row-major.  Tree construction incl. the initial frontier / min-heuristic rows
 * (:73-78) runs on the device. */
int imomd_planner_create(imomd_graph* g, int n_queries, int K, const uint32_t* dest,
                         const double* prob, const imomd_params* params, imomd_planner** out);
int imomd_planner_destroy(imomd_planner* p);
/* new destinations / probability matrices for every query (same Q and K, host buffers
 * laid out as for imomd_planner_create), followed by a reset */
int imomd_planner_set_queries(imomd_planner* p, const uint32_t* dest, const double* prob);
/* back to the freshly constructed state (same dest / prob), word cursors to 0 */
int imomd_planner_reset(imomd_planner* p);

/* RNG sample stream injected from the host: raw std::mt19937 words, appended to the
 * planner's stream (all queries read the same stream, each at its own cursor --
 * the reference seeds every planner with mt19937{0}, :198-201).  The device turns
 * words into samples with libstdc++'s formulas (selectRandomVertex_, :358-400). */
int imomd_planner_feed_words(imomd_planner* p, const uint32_t* words, size_t n);
int imomd_planner_words_available(const imomd_planner* p, uint64_t* total_fed);
/* forget the injected stream (the next feed starts again at word 0) */
int imomd_planner_clear_words(imomd_planner* p);

/* `iters` passes of expandTreeLayers_ (:329-356) for every query of the planner, in
 * one persistent kernel.  reports[n_queries] may be NULL.  Asynchronous variant:
 * imomd_expand_async enqueues on the planner's stream, imomd_sync waits and fills
 * the reports.  A query whose word cursor reaches the end of the injected stream stops
 * after its last complete expansion with status IMOMD_RUN_NEED_WORDS and
 * `iterations` < iters: feed more words and call again with the REMAINING passes.
 * That resume is per call, not per query -- a planner with several queries must be
 * fed up front (an expansion draws at most 8 words: iters * K * 8 + a margin covers
 * every query), because a second call runs its `iters` for all of them. */
int imomd_expand(imomd_planner* p, int iters, imomd_query_report* reports);
int imomd_expand_async(imomd_planner* p, int iters);
int imomd_sync(imomd_planner* p, imomd_query_report* reports);
/* BATCHED expansion (north_star (2) "batched random sampling", (4) "atomic cost/parent updates resolved
 * deterministically"): `steps` steps, each of them B expansions of EVERY tree of every query.
 *   1. attach: the K trees of a query concurrently, one warp per tree: sample, nearest frontier
 *      vertex, jump-point hops, choose-parent, frontier / min-heuristic maintenance (the code of the
 *      exact mode on the tree's own records).  Samples come from the injected stream in fixed
 *      slots of 8 words per (pass, tree).
 *   2. rewire: rewireTree_ (:561-600) for all new vertices of all trees at once, as a
 *      label-correcting relaxation over the road graph by the whole CTA: a vertex whose cost went
 *      down offers cost + w to its visited neighbours, a neighbour takes the offer when its
 *      (cost, parent id) gets lexicographically smaller -- one 128-bit compare-and-swap on the
 *      record head.  The fixed point (cheapest cost through visited vertices, lowest parent id
 *      among equal offers) is independent of the order of the offers.
 *   3. connect: updateConnectionTree_ (:645-706) of every vertex attached or changed in the step
 *      with the step's final costs; per pair of trees the smallest sum (lowest vertex on ties)
 *      replaces the distance when it beats it by more than 0.1 (:680).
 *   4. the updateSelectionProbability tests (:602-643) in tree order and program order.
 * Deterministic (independent of warp timing); NOT the reference's sequence: the exact,
 * batch-size-1 parity mode is imomd_expand.  Costs are never worse than what the reference's
 * cascade leaves for the same visited vertices.  A report's `iterations` counts passes
 * (steps * B).  Not available in pseudo mode; a resident planner must be reset before it goes
 * back to imomd_expand (child rows are not maintained by the relaxation). */
int imomd_expand_batch(imomd_planner* p, int steps, int B, imomd_query_report* reports);
int imomd_expand_batch_async(imomd_planner* p, int steps, int B);
/* ImomdRRT::findShortestPath's max_time test runs after EVERY pass (:258-262): with a budget
 * set, an imomd_expand launch stops after the first pass that ends more than `seconds` after
 * the launch began (imomd_query_report.iterations tells how many ran).  0 = none. */
int imomd_set_time_budget(imomd_planner* p, double seconds);
/* The per-request cycle counters below cost three clock reads per state-machine step, so the
 * kernels keep them only after imomd_set_profiling(p, 1) (default: off, or on when the planner is
 * created with IMOMD_PROFILE=1 in the environment).  Takes effect with the next launch. */
int imomd_set_profiling(imomd_planner* p, int on);
/* cumulative SM-clock cycles and call counts per kind of work of one query: index 1..6 =
 * nearest search, heuristic rescan, heuristic insert, connection gather, cost propagation,
 * rewire batch check; index 7 = the serial steps in between (development / profiling aid) */
int imomd_get_profile(imomd_planner* p, int query, uint64_t* cycles, uint64_t* counts);
/* exact mode: cycles / counts of the serial steps by the phase they started in (next tree, after
 * nearest, jump-point search, insert, finish, rewire, last connection check); batched mode: cycles
 * the per-tree groups spent per kind of work (index as in imomd_get_profile), summed over groups */
int imomd_get_phase_cycles(imomd_planner* p, int query, uint64_t* cycles, uint64_t* counts);
/* the last comparison the device could not decide within 1e-12 relative (see
 * imomd_query_report.unresolved_ties): {kind 1 nearest / 2 heuristic rescan, tree, other tree,
 * best value, runner-up value, vertex taken, sample lat, sample lon} */
int imomd_get_tie_info(imomd_planner* p, int query, double* out);
/* cumulative work counters of every query, out[Q][8]: hash cells bounded, frontier records
 * ranked by nearest searches, rewireTree_ calls, adjacency slots read by serial scans,
 * re-parentings, cost-propagated descendants, connection-checked vertices, frontier
 * records ranked by heuristic rescans (the inputs of the algorithmic-bytes figure) */
int imomd_get_counters(imomd_planner* p, uint64_t* out);
/* device time of the last imomd_expand kernel in milliseconds (CUDA events) */
int imomd_last_kernel_ms(imomd_planner* p, float* ms);

/* K x K state of one query (any pointer may be NULL):
 * distance_matrix_, connection_node_matrix_, probability_matrix_,
 * expandables_min_heuristic_matrix_ (ids, values), tree is_done flags,
 * disjoint_set_parent_ (include/imomd_rrt_star/imomd_rrt_star.h:159-178) */
int imomd_get_state(imomd_planner* p, int query, double* D, uint32_t* conn_node, double* prob,
                    uint32_t* mh_id, double* mh_val, int32_t* is_done, int32_t* ds_parent);
/* the K x K matrices (and is_done flags) of queries [first, first + count) in one device->host
 * copy; every output holds `count` consecutive blocks, any of them may be NULL */
int imomd_get_states(imomd_planner* p, int first, int count, double* D, uint32_t* conn_node, double* prob,
                     int32_t* is_done);
/* Fingerprint of a query's K trees, computed on the device: tree_hash = FNV-style fold over the
 * trees in id order, each tree's visited vertices in ascending id, of (vertex, parent, bits of
 * cost) -- what a walk over the reference's tree.parent / tree.cost maps (imomd_rrt_star.h:66-76)
 * in sorted order produces; tree_sum = order-free 64-bit checksum of the same entries. */
int imomd_get_treehash(imomd_planner* p, int query, uint64_t* tree_hash, uint64_t* tree_sum);
/* layout of the planner: queries, state slots, 1 when it streams, requests answered by the
 * exact-value service (near-ties decided with the host's libm, see DESIGN.md "Exactness") */
int imomd_planner_info(imomd_planner* p, int32_t* n_queries, int32_t* slots, int32_t* streaming,
                       uint64_t* certified_requests);
/* is_connected_graph_ / is_distance_matrix_updated_ (imomd_rrt_star.h:178, :190) */
int imomd_get_flags(imomd_planner* p, int query, int32_t* connected, int32_t* dm_updated);
/* the facade clears is_distance_matrix_updated_ after updatePath_ (:960) */
int imomd_clear_dm_updated(imomd_planner* p, int query);

/* dense tree arrays: parent[N] (IMOMD_NONE = not visited) and cost[N] (+inf there) */
int imomd_get_tree(imomd_planner* p, int query, int tree, uint32_t* parent, double* cost);
/* frontier ("expandables") of one tree, unordered; *count receives its size */
int imomd_get_frontier(imomd_planner* p, int query, int tree, uint32_t* out, uint32_t cap,
                       uint32_t* count);
/* sparse readback: parent (IMOMD_NONE = not visited) and cost (INFINITY then) of the listed
 * vertices of one tree; either output may be NULL */
int imomd_gather_tree(imomd_planner* p, int query, int tree, const uint32_t* vertices, uint32_t count,
                      uint32_t* parent, double* cost);
/* parent-pointer walk vertex -> root of `tree` (updatePath_, :936-954); out[0] = vertex */
int imomd_walk_to_root(imomd_planner* p, int query, int tree, uint32_t vertex, uint32_t* out,
                       uint32_t cap, uint32_t* len);
/* pseudo mode: ordered (set index, vertex) insertions into connection_nodes_set_ (:692-703) */
int imomd_get_connection_log(imomd_planner* p, int query, uint32_t* set_idx, uint32_t* vertex,
                             uint32_t cap, uint32_t* count);
/* forget the logged entries (a caller that has consumed them all keeps the log from filling up) */
int imomd_clear_connection_log(imomd_planner* p, int query);
/* same, starting at entry `first` (*count still receives the total length) */
int imomd_get_connection_log_from(imomd_planner* p, int query, uint32_t first, uint32_t* set_idx,
                                  uint32_t* vertex, uint32_t cap, uint32_t* count);

/* Pseudo-tree merge support (mergePseudoTrees_ / mergeTwoTree_, :744-916; the merge itself is
 * driven by the host facade because it iterates std containers in libstdc++ order):
 *   imomd_attach          connectNewNode_ + updateExpandables + rewireTree_ +
 *                         updateConnectionTree_ of `vertex` in `tree` (:848-851, :883-886);
 *                         parent_out receives the parent chosen (IMOMD_NONE: not connectable)
 *   imomd_attach_list     the same for `count` vertices one after the other in ONE launch
 *                         (each sees the tree as the previous ones left it); parents_out[i]
 *                         = parent of vertices[i] after the whole list
 *   imomd_get_visit_order vertices of a tree in the order they were visited = insertion order
 *                         of the reference's tree.parent map (pseudo mode only)
 *   imomd_set_distance    distance / connection-node entry (i,j) and (j,i), raises dm_updated
 *   imomd_set_tree_done, imomd_set_contact   is_done flag; "connection set non-empty" flag */
int imomd_attach(imomd_planner* p, int query, int tree, uint32_t vertex, uint32_t* parent_out,
                 imomd_query_report* report);
int imomd_attach_list(imomd_planner* p, int query, int tree, const uint32_t* vertices, uint32_t count,
                      uint32_t* parents_out, imomd_query_report* report);
int imomd_get_visit_order(imomd_planner* p, int query, int tree, uint32_t* out, uint32_t cap,
                          uint32_t* count);
int imomd_set_distance(imomd_planner* p, int query, int i, int j, double d, uint32_t conn_node);
int imomd_set_tree_done(imomd_planner* p, int query, int tree, int done);
int imomd_set_contact(imomd_planner* p, int query, int set_idx);

/* Parity hook for the nearest-frontier-vertex search (steerNewNode_ arg-min,
 * :406-418): n independent (query, tree, lat, lon) lookups against the current
 * frontiers; idx = IMOMD_NONE when the frontier is empty; dist = haversine metres. */
int imomd_nearest(imomd_planner* p, int n, const imomd_nn_query* q, uint32_t* idx, double* dist);

/* Same answers for LARGE batches: lookups are grouped into tiles of one frontier x 4 samples
 * and every frontier is streamed once per tile through a shared-memory ring filled by bulk
 * asynchronous copies of the occupied records (the bandwidth-shaped brute-force variant).
 * Optional outputs: device time of the tile kernel, algorithmic bytes it streamed (32 B per
 * occupied frontier record per tile + 8 B per chunk-list entry + 36 B per lookup), number of
 * lookups re-done on the exact path. */
int imomd_nearest_batch(imomd_planner* p, int n, const imomd_nn_query* q, uint32_t* idx, double* dist,
                        float* kernel_ms, uint64_t* bytes_streamed, uint32_t* n_refined);
/* Device times of the stages of the last imomd_nearest_batch call on this planner (chunk lists +
 * seeds + segment records, tile kernel, merge of the segment partials) and its number of segments. */
int imomd_nearest_batch_last_timing(imomd_planner* p, float* prep_ms, float* tiles_ms, float* merge_ms,
                                    uint32_t* n_segments);

/* ---------------------------------------------------------------- ECI-Gen ---
 * EciGenSolver::calculatePathCost_ (src/eci_gen_tsp_solver.cpp:442-450) for a batch
 * of tours: one warp per tour, sequential left-to-right sum (bit-exact). */
int imomd_ctx_create(int device, imomd_ctx** out);
int imomd_ctx_destroy(imomd_ctx* c);
int imomd_ga_eval(imomd_ctx* c, int K, const double* D, int n_tours, int stride,
                  const int32_t* tours, const int32_t* lens, double* cost_out);

const char* imomd_last_error(void);
const char* imomd_version(void);
/* bytes imomd_get_state moves device->host per query */
size_t imomd_sizeof_query_state(void);

#ifdef __cplusplus
}
#endif
#endif /* IMOMD_B200_H */
