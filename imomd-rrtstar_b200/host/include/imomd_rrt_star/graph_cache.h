// imomd_rrt_star/graph_cache.h -- binary cache of a parsed road graph (an addition: SURVEY.md 8f-4).
//
// The reference parses its .osm input with a single-threaded XML DOM walk on every start
// (src/osm_parser.cpp:48-184 upstream), which dominates the wall time of a run on a
// million-vertex map.  GraphCache stores what the parser produced -- the `location_t` vector
// and the per-vertex `unordered_map<size_t, double>` adjacency (main.cpp:114-131 upstream) --
// as one flat file: coordinates as SoA doubles and the adjacency as CSR **in the containers'
// own iteration order**, which is part of the planner's behaviour (tie-breaks of choose-parent,
// the rewire sequence).  load() rebuilds containers that iterate in exactly that order and
// verifies it; a file written with another standard library that cannot be reproduced is
// rejected instead of silently changing the planner's results.
#ifndef IMOMD_B200_GRAPH_CACHE_H
#define IMOMD_B200_GRAPH_CACHE_H

#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "data/location_t.h"

namespace imomd_b200 {

struct GraphCache
{
    using Map = std::vector<location_t>;
    using Connection = std::vector<std::unordered_map<size_t, double>>;

    // Throws std::invalid_argument for ids that differ from their index or a degree above 13
    // (beyond it the containers rehash and their order cannot be replayed), std::runtime_error
    // on I/O errors.
    static void save(const std::string& path, const Map& map, const Connection& connection);

    // Returns false when the file does not exist; throws std::runtime_error for a damaged
    // file or when the rebuilt containers do not iterate in the recorded order.
    static bool load(const std::string& path, std::shared_ptr<Map>& map,
                     std::shared_ptr<Connection>& connection);
};

}  // namespace imomd_b200

#endif  // IMOMD_B200_GRAPH_CACHE_H
