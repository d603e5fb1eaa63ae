// csrc/imomd_layout.h -- how the expansion path's state is laid out in HBM.
//
// Plain structs shared by the CUDA kernels (imomd_kernels.cu), the sequential /
// CTA-parallel device logic (imomd_core.cuh) and the host-side C ABI glue.
//
// Road graph (uploaded once, read-only, shared by every query of every planner):
//   row_ptr u32[N+1], col u32[2E], w f64[2E]   CSR, native container order
//   lat f64[N], lon f64[N]                     SoA coordinates (degrees)
//   vrec VRec[N]                               32-byte vertex records: unit vector
//                                              (x,y,z), id, frontier-hash cell
// Per query (one CTA drives one query):
//   QueryDev header                            K x K matrices, flags, counters
//   lbd f64[K][G*G]                            haversine lower bound root -> cell
//   arena u32[arena_entries]                   rewire scratch (updateCost lists, frames)
//   clist u32x2[chunks_per_tree]               chunk work list of the pruned searches
// Per (query, tree):
//   trec[N]: {parent u32 (NONE = not visited), frontier slot u32, cost f64,
//             child[width] u32 (NONE-terminated, container order)}  -- 32 B when width = 4
//   cell_head u32[G*G], cell_cnt u32[G*G]      frontier hash: per-cell chunk chain
//   occ u32[2*G*G]                             occupied cells (list + inverse)
//   rec VRec[chunks*32], chunk_next u32[chunks], free_stack u32[chunks]
#ifndef IMOMD_LAYOUT_H
#define IMOMD_LAYOUT_H

#include <stddef.h>
#include <stdint.h>

#ifndef IMOMD_KMAX
#define IMOMD_KMAX 32u
#endif
#ifndef IMOMD_NONE
#define IMOMD_NONE 0xFFFFFFFFu
#endif
#define IMOMD_CHUNK 32               // frontier records per chunk (one warp-wide load)
#define IMOMD_MAXDEG 13
#define IMOMD_WORD_MARGIN 64         // words that must be available to start an expansion

// 32-byte record: two 128-bit loads.  Doubles as the frontier entry.
struct __attribute__((aligned(16))) VRec
{
    double x, y, z;      // unit vector of (lat, lon)
    uint32_t id;         // vertex
    uint32_t cell;       // frontier-hash cell (row * G + col)
};

struct GraphDev
{
    uint32_t n;
    uint32_t arcs;
    // Adjacency in PADDED rows of `width` slots (width = power of two >= the maximum degree):
    // the neighbours of v are col[v*width .. v*width + deg[v]) in the reference containers'
    // iteration order, the remaining slots hold IMOMD_NONE / 0.  A row's address follows from
    // the vertex id alone -- no row_ptr load in front of every neighbourhood access.
    const uint8_t* deg;
    const uint32_t* col;
    const double* w;
    const double* lat;
    const double* lon;
    const VRec* vrec;
    const double* coslat;  // cos(lat * pi/180) of every vertex, evaluated once at upload with the
                           // expression the haversine uses (bit-identical, two libm calls less per call)
    int32_t G;            // cells per axis
    int32_t prune;        // 0 = map too large for the cell bound: scan every cell
    int32_t width;        // adjacency slots per vertex in the batched checks: pow2 >= max degree
    int32_t pad_;
    double lat0, lon0;    // south-west corner of the hash (degrees)
    double hlat, hlon;    // cell size (degrees)
    const double* cosrow; // [2G]: min cos(lat) over the padded latitude span of each cell row,
                          //       then the max over the same span
};

// K x K matrices are stored with stride K.
struct QueryDev
{
    int32_t K;
    int32_t status;
    int32_t connected;
    int32_t dm_updated;
    uint32_t dest[IMOMD_KMAX];            // destinations_
    int32_t view_id[IMOMD_KMAX];          // tree_layers_[i].id   (0 until constructed)
    uint32_t view_root[IMOMD_KMAX];       // tree_layers_[i].root (0 until constructed)
    int32_t is_done[IMOMD_KMAX];
    int32_t ds_parent[IMOMD_KMAX];
    uint32_t frontier_count[IMOMD_KMAX];
    uint32_t free_top[IMOMD_KMAX];        // free chunks left on the tree's stack
    uint32_t n_occ[IMOMD_KMAX];           // occupied cells of the tree's frontier hash
    uint64_t tree_size[IMOMD_KMAX];       // parent.size()
    uint8_t contact[IMOMD_KMAX * (IMOMD_KMAX - 1) / 2 + 8];   // connection_nodes_set_[i] non-empty
    uint64_t word_cursor;                 // next word of the injected stream
    int64_t iterations;                   // cumulative
    int64_t expansions;                   // cumulative
    int64_t near_ties;
    int64_t unresolved_ties;
    uint32_t log_count;                   // pseudo-mode connection log entries
    int32_t resume_k;                     // tree to resume at inside an interrupted pass
    // what the last launch did
    int64_t last_iterations;
    int64_t last_expansions;
    int32_t active_trees;
    int32_t pad0;
    // SM-clock cycles and call counts per request type (index = ReqType, 7 = serial steps)
    uint64_t prof_cycles[8];
    uint64_t prof_count[8];
    // work counters behind the algorithmic-bytes figure ("roofline"):
    // 0 hash cells bounded, 1 frontier records ranked (nearest), 2 rewireTree_ calls,
    // 3 adjacency slots read by serial scans, 4 re-parentings, 5 cost-propagated
    // descendants, 6 connection-checked vertices, 7 frontier records ranked (rescans)
    uint64_t stat[8];
    double tie_info[8];         // last undecidable comparison: kind (1 nearest, 2 rescan), tree,
                                // other, best, runner-up, vertex, sample lat, sample lon
    uint64_t phase_cycles[8];   // serial-step cycles by the phase the step started in
    uint64_t phase_count[8];
    int64_t certified_ties;     // near-ties decided with the host's glibc values (Cta::certify)
    uint64_t tree_hash;         // SURVEY App. A TREEHASH of the K trees (streaming planners: written when
                                // the query finishes; resident planners: by imomd_get_treehash)
    uint64_t tree_sum;          // order-free companion: sum of mix64(tree, vertex, parent, cost bits)
    uint64_t busy_cycles;       // SM-clock cycles this query kept its CTA busy (last launch)
    // ---- everything above is staged in shared memory while a launch runs (hot fields);
    // ---- the K x K matrices below are staged separately, compacted to stride K
    double P[IMOMD_KMAX * IMOMD_KMAX];    // probability_matrix_
    double D[IMOMD_KMAX * IMOMD_KMAX];    // distance_matrix_
    double MHv[IMOMD_KMAX * IMOMD_KMAX];  // expandables_min_heuristic_matrix_ .second
    uint32_t MHi[IMOMD_KMAX * IMOMD_KMAX];// ... .first
    uint32_t CN[IMOMD_KMAX * IMOMD_KMAX]; // connection_node_matrix_
};
#define IMOMD_QUERY_HOT_BYTES (offsetof(QueryDev, P))

// mirrors imomd_query_report (include/imomd_b200.h)
struct ReportDev
{
    int32_t status;
    int32_t active_trees;
    int64_t iterations;
    int64_t expansions;
    int64_t tree_vertices;
    int64_t words_consumed;
    int32_t connected;
    int32_t dm_updated;
    int64_t near_ties;
    int64_t unresolved_ties;
    int64_t log_entries;
    int64_t certified_ties;
    uint64_t tree_hash;
    uint64_t busy_cycles;
};

// Mailbox between one CTA and the host thread that waits for the launch (mapped pinned host memory):
// the device posts coordinate pairs of a comparison it cannot decide in its own arithmetic, the host
// answers with the reference's haversine evaluated by glibc (include/osm_converter/
// compute_haversine.hpp:41-58).  state: 0 idle, 1 request posted, 2 answer posted.
#define IMOMD_TIE_PAIRS 16
#define IMOMD_TIE_IDS 64     /* candidates of one certified arg-min (asked in rounds of IMOMD_TIE_PAIRS) */
struct TieBox
{
    uint32_t state;
    uint32_t n;
    uint32_t pad_[2];
    double in[IMOMD_TIE_PAIRS][4];   // lat1, lon1, lat2, lon2 (degrees)
    double out[IMOMD_TIE_PAIRS];     // metres
};

// Batched mode: one deferred cross-tree effect of an expansion.
//   kind 0: updateConnectionTree_(tree, x) -- val = cost of x in the tree at that moment
//   kind 1: updateSelectionProbability for the pair (tree, aux) -- val / extra = the
//           min-heuristic value / holder of that pair at that moment
struct BEvent
{
    uint32_t x;
    uint32_t aux;
    double val;
    uint32_t extra;
    uint32_t kind;
};

struct PlannerDev
{
    GraphDev g;
    int32_t Q;                 // queries
    int32_t K;
    int32_t slots;             // sets of per-query device state (trees, frontier hashes, scratch):
                               // == Q for a resident planner; < Q for a streaming planner, whose
                               // queries take a slot from a work queue, run to completion and leave
                               // only their header (matrices, counters, tree hash) behind
    int32_t streaming;         // slots < Q
    double goal_bias;
    double tie_rel;            // comparisons closer than this (relative) are decided on the host's values
    uint64_t budget_ns;        // > 0: a launch stops after the first pass that ends later than this many
                               // nanoseconds after its start (max_time, src/imomd_rrt_star.cpp:258-262)
    int32_t pseudo;
    int32_t chunks;            // chunks per tree
    uint32_t arena_entries;
    uint32_t log_entries;
    QueryDev* query;           // [Q]
    const QueryDev* query_init;// [Q] freshly constructed headers (streaming planners restart from them)
    uint32_t* queue;           // [2] streaming planners: next query to hand out, queries finished
    TieBox* tiebox;            // [max grid] host mailboxes (device address of mapped pinned memory)
    // per (slot, tree), index qt = slot*K + k.  A resident planner has slot == query.
    unsigned char* trec;       // [Q*K][n] vertex records of trec_stride bytes:
                               //   {u32 parent, u32 frontier slot, f64 cost, u32 child[width]}
    uint32_t trec_stride;      // 16 + 4 * width
    uint32_t dirty_words;      // streaming planners: words of the per-tree "record touched" bitmap that
                               // follows each tree's occ lists (slot recycling and the tree hash walk it
                               // instead of all n records); 0 = no bitmap (resident planners)
    uint32_t* cell_head;       // [Q*K][G*G]
    uint32_t* cell_cnt;        // [Q*K][G*G]
    uint32_t* occ;             // [Q*K][2*G*G + dirty_words] list of occupied cells, cell -> list position,
                               // then the dirty bitmap of the tree's records (bit v: record v was written)
    VRec* rec;                 // [Q*K][chunks*32]
    uint32_t* chunk_next;      // [Q*K][chunks]
    uint32_t* free_stack;      // [Q*K][chunks]
    // per slot
    double* lbd;               // [Q][K][G*G] lower bound root -> cell (metres)
    double* ubd;               // [Q][K][G*G] upper bound root -> cell (metres)
    uint32_t* arena;           // [Q][arena_entries]
    uint32_t* clist;           // [Q][2*chunks]   (chunk id, valid count)
    uint32_t* log;             // [Q][2*log_entries] (set idx, vertex)
    uint32_t* visit_log;       // pseudo mode: [Q*K][n] vertices in the order they were visited
    uint32_t* attach_list;     // pseudo mode: [n] vertices imomd_attach_list works through; each entry is
                               // overwritten with the parent the vertex received
    struct ReportDev* reports; // [Q] compact per-launch report (copied back to the host)
    // batched mode (imomd_expand_batch): per (slot, tree) event logs and chunk work lists
    BEvent* events;            // [slots*K][ev_cap]
    uint32_t ev_cap;
    uint32_t profile;          // 1: the kernels keep the per-request cycle counters of imomd_get_profile /
                               // imomd_get_phase_cycles (three clock reads per state-machine step; off by default)
    uint32_t* clist_tree;      // [slots*K][2*chunks]
    // injected random words
    const uint32_t* words;
    uint64_t words_avail;
};

#endif  // IMOMD_LAYOUT_H
