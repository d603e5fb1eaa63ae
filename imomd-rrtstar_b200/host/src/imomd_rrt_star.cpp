// host/src/imomd_rrt_star.cpp -- host side of the drop-in planner.
//
// What the reference's ImomdRRT does per pass on the CPU (expandTreeLayers_ and everything
// below it, src/imomd_rrt_star.cpp:329-742 upstream) happens on the GPU behind
// include/imomd_b200.h.  This file keeps the parts that stay on the host:
//   * construction: containers -> CSR (native iteration order) + SoA upload, the
//     destination-selection probabilities (:84-185, see probability.cpp), RNG seeding
//     (:193-201), CSV log file (:203-222);
//   * the anytime loop of findShortestPath (:225-291): expansion passes in batches that end
//     exactly where the reference calls solveRTSP_ (every 100 passes), max_iter / max_time;
//   * solveRTSP_ (:966-1003), updatePath_ (:918-964), logData_ (:1005-1029), printPath
//     (:293-322).
// max_time is enforced after every pass, as upstream, by a time budget the device kernel checks.
#include "imomd_rrt_star/imomd_rrt_star.h"

#include <algorithm>
#include <cmath>
#include <ctime>
#include <iomanip>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <stdexcept>

#include <unistd.h>

#include "host_api.h"
#include "imomd_b200.h"

namespace {

int g_default_device = 0;

std::string fixed4(double v)
{
    std::ostringstream out;
    out.precision(4);
    out << std::fixed << v;
    return out.str();
}

// IMOMD_FACADE_TIMING=1: wall time of the facade's stages on stderr (development aid)
struct StageTimer
{
    const char* name;
    std::chrono::steady_clock::time_point t0;
    explicit StageTimer(const char* n) : name(n), t0(std::chrono::steady_clock::now()) {}
    ~StageTimer()
    {
        static const bool on = std::getenv("IMOMD_FACADE_TIMING") != nullptr;
        if (on)
            std::cerr << "[facade] " << name << " "
                      << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()
                      << " ms" << std::endl;
    }
};

void check(int rc, const char* what)
{
    if (rc != IMOMD_OK)
        throw std::runtime_error(std::string("ImomdRRT: ") + what + ": " + imomd_last_error());
}

}  // namespace

void ImomdRRT::setDefaultDevice(int device) { g_default_device = device; }

ImomdRRT::ImomdRRT(size_t source, size_t target, std::vector<size_t> objectives,
                   const std::shared_ptr<std::vector<location_t>> map,
                   const std::shared_ptr<std::vector<std::unordered_map<size_t, double>>> connection,
                   const imomd_setting_t& setting)
    : setting_(setting), map_(map), connection_(connection), device_(g_default_device)
{
    destinations_.push_back(source);
    for (size_t o : objectives) destinations_.push_back(o);
    destinations_.push_back(target);
    K_ = static_cast<int>(destinations_.size());
    if (K_ > static_cast<int>(IMOMD_KMAX))
        throw std::invalid_argument("ImomdRRT (B200): more than 30 objectives");
    const size_t n = map_->size();
    if (n == 0 || n > 0xFFFFFFF0ull || connection_->size() != n)
        throw std::invalid_argument("ImomdRRT (B200): bad map / connection sizes");
    for (size_t d : destinations_)
        if (d >= n) throw std::invalid_argument("ImomdRRT (B200): destination id out of range");

    // road graph -> CSR in the containers' own iteration order, coordinates -> SoA
    std::vector<uint32_t> row_ptr(n + 1), col;
    std::vector<double> w, lat(n), lon(n);
    size_t arcs = 0;
    for (const auto& row : *connection_) arcs += row.size();
    col.reserve(arcs);
    w.reserve(arcs);
    for (size_t v = 0; v < n; ++v)
    {
        const location_t& loc = (*map_)[v];
        if (loc.id != v)
            throw std::invalid_argument(
                "ImomdRRT (B200): location ids must equal their index (as OSMParser / FakeMap produce)");
        lat[v] = loc.latitude;
        lon[v] = loc.longitude;
        row_ptr[v] = static_cast<uint32_t>(col.size());
        for (const auto& kv : (*connection_)[v])
        {
            col.push_back(static_cast<uint32_t>(kv.first));
            w.push_back(kv.second);
        }
    }
    row_ptr[n] = static_cast<uint32_t>(col.size());
    check(imomd_graph_create(device_, static_cast<uint32_t>(n), row_ptr.data(), col.data(), w.data(),
                             lat.data(), lon.data(), &graph_),
          "graph upload");

    std::vector<uint32_t> dest(K_);
    std::vector<double> dlat(K_), dlon(K_), prob(static_cast<size_t>(K_) * K_);
    for (int i = 0; i < K_; ++i)
    {
        dest[i] = static_cast<uint32_t>(destinations_[i]);
        dlat[i] = lat[destinations_[i]];
        dlon[i] = lon[destinations_[i]];
    }
    imomd_host::selection_probabilities(K_, dlat.data(), dlon.data(), prob.data());
    imomd_params params{};
    params.goal_bias = setting_.goal_bias;
    params.pseudo_mode = setting_.pseudo_mode ? 1 : 0;
    // Device pools of the one query this planner drives.  The library's automatic sizes are meant
    // for hundreds of queries per GPU (a frontier of ~N/8 vertices per tree); a single anytime run
    // may grow tendril-like trees whose frontier is a multiple of that, and the reference has no
    // capacity limit at all -- so be generous here (a frontier of up to N vertices per tree while
    // that stays below ~4 GB, nested rewire lists of 4 N entries), and let the environment override:
    // IMOMD_FRONTIER_CHUNKS (32-record chunks per tree), IMOMD_ARENA_ENTRIES (uint32 entries).
    {
        const double per_tree = static_cast<double>(n) / 32.0 + 8192.0;
        if (per_tree * 1024.0 * K_ < 4e9) params.frontier_chunks = static_cast<uint32_t>(per_tree);
        params.arena_entries = static_cast<uint32_t>(std::min<size_t>(std::max<size_t>(1u << 20, 4 * n), 1u << 28));
        if (const char* e = std::getenv("IMOMD_FRONTIER_CHUNKS")) params.frontier_chunks = static_cast<uint32_t>(std::strtoul(e, nullptr, 10));
        if (const char* e = std::getenv("IMOMD_ARENA_ENTRIES")) params.arena_entries = static_cast<uint32_t>(std::strtoul(e, nullptr, 10));
    }
    if (setting_.pseudo_mode) connection_sets_.resize(static_cast<size_t>(K_) * (K_ - 1) / 2);
    check(imomd_planner_create(graph_, 1, K_, dest.data(), prob.data(), &params, &planner_),
          "planner creation");

    solver_ = EciGenSolver(setting_.rtsp_setting, device_);
    shortest_path_cost_ = INFINITY;
    if (setting_.random_seed)
    {
        std::random_device rd;
        rng_ = std::mt19937{rd()};
    }
    else
    {
        rng_ = std::mt19937{0};
    }
    start_ = std::chrono::steady_clock::now();
    if (setting_.log_data)
    {
        std::time_t t = std::time(nullptr);
        std::tm tm = *std::localtime(&t);
        std::ostringstream stamp;
        stamp << std::put_time(&tm, "%Y-%d-%m-%H-%M-%S");
        char cwd[4096];
        std::string dir = getcwd(cwd, sizeof(cwd)) ? cwd : ".";
        csv_.open(dir + "/experiments/IMOMD_" + stamp.str() + "_command_history.csv");
        if (csv_.good())
            csv_ << "\"CPU_time\";\"path_cost\";\"tree_size\";" << std::endl;
        else
            std::cout << "Exception was thrown: cannot open the log file" << std::endl;
    }
}

ImomdRRT::~ImomdRRT()
{
    if (planner_) imomd_planner_destroy(planner_);
    if (graph_) imomd_graph_destroy(graph_);
}

void ImomdRRT::feedWords(size_t n)
{
    std::vector<uint32_t> words(n);
    for (uint32_t& x : words) x = static_cast<uint32_t>(rng_());
    check(imomd_planner_feed_words(planner_, words.data(), words.size()), "sample stream");
    words_fed_ += n;
}

// `passes` calls of expandTreeLayers_ on the device
bool ImomdRRT::expandBatch(int passes)
{
    StageTimer timer("expandBatch");
    int left = passes;
    while (left > 0)
    {
        // every expansion draws 4 or 6 words (a few more for the rare Lemire rejection): top the
        // injected stream up to what `left` passes can consume, counting what is still unread
        const size_t want = static_cast<size_t>(left) * K_ * 8 + 256;
        const size_t unread = words_fed_ - words_consumed_;
        if (unread < want) feedWords(want - unread);
        // max_time is tested after every pass upstream (:258-262): the launch gets what is left
        // of the budget and stops by itself after the pass that crosses it
        const double remaining = static_cast<double>(setting_.max_time) -
                                 std::chrono::duration<double>(std::chrono::steady_clock::now() - start_).count();
        check(imomd_set_time_budget(planner_, std::max(remaining, 1e-6)), "time budget");
        imomd_query_report rep{};
        check(imomd_expand(planner_, left, &rep), "expansion");
        left -= static_cast<int>(rep.iterations);
        iteration_count_ += rep.iterations;
        expansions_ += rep.expansions;
        tree_vertices_ = rep.tree_vertices;
        active_trees_ = rep.active_trees;
        connected_ = rep.connected != 0;
        dm_updated_ = rep.dm_updated != 0;
        words_consumed_ = static_cast<size_t>(rep.words_consumed);
        if (rep.status == IMOMD_RUN_FRONTIER_FULL || rep.status == IMOMD_RUN_ARENA_FULL)
            throw std::runtime_error(std::string("ImomdRRT (B200): device ") +
                                     (rep.status == IMOMD_RUN_FRONTIER_FULL ? "frontier pool" : "rewire scratch") +
                                     " exhausted; raise IMOMD_FRONTIER_CHUNKS / IMOMD_ARENA_ENTRIES");
        if (rep.status != IMOMD_RUN_OK && rep.status != IMOMD_RUN_NEED_WORDS)
            throw std::runtime_error("ImomdRRT (B200): device run status " + std::to_string(rep.status));
        if (rep.status == IMOMD_RUN_OK && left > 0) break;   // the time budget ended the launch
    }
    if (setting_.pseudo_mode) drainConnectionLog();
    return true;
}

// connection_nodes_set_ insertions the device logged since the last call, applied in order
void ImomdRRT::drainConnectionLog()
{
    if (merge_done_) return;   // the sets are only read by the merge (as upstream)
    for (;;)
    {
        std::vector<uint32_t> idx(1 << 16), vx(1 << 16);
        uint32_t total = 0;
        check(imomd_get_connection_log_from(planner_, 0, log_consumed_, idx.data(), vx.data(),
                                            static_cast<uint32_t>(idx.size()), &total),
              "connection log");
        const uint32_t got = std::min<uint32_t>(total - log_consumed_, static_cast<uint32_t>(idx.size()));
        for (uint32_t i = 0; i < got; ++i) connection_sets_[idx[i]].insert(vx[i]);
        log_consumed_ += got;
        if (log_consumed_ >= total) break;
    }
    if (log_consumed_ > 0)
    {
        check(imomd_clear_connection_log(planner_, 0), "connection log");
        log_consumed_ = 0;
    }
}

// connectNewNode_ + updateExpandables + rewireTree_ + updateConnectionTree_ (:848-851 /
// :883-886 upstream) of the listed vertices, one after the other, in one device launch; then the
// connection-set insertions they made are applied to the host mirrors.
void ImomdRRT::attachList(int tree, const std::vector<uint32_t>& vertices)
{
    if (vertices.empty()) return;
    std::vector<uint32_t> got(vertices.size());
    imomd_query_report rep{};
    check(imomd_attach_list(planner_, 0, tree, vertices.data(), static_cast<uint32_t>(vertices.size()),
                            got.data(), &rep),
          "tree merge");
    if (rep.status != IMOMD_RUN_OK)
        throw std::runtime_error("ImomdRRT (B200): device status " + std::to_string(rep.status) +
                                 " during the tree merge");
    for (uint32_t parent : got)
        if (parent == IMOMD_NONE)
            throw std::runtime_error("ImomdRRT (B200): a vertex did not connect during the tree merge");
    tree_vertices_ = rep.tree_vertices;
    if (rep.dm_updated) dm_updated_ = true;
    drainConnectionLog();
}

// mergeTwoTree_, src/imomd_rrt_star.cpp:819-916 upstream
void ImomdRRT::mergeTwoTrees(int parent_tree, int child_tree)
{
    StageTimer timer("mergeTwoTrees");
    const size_t n = map_->size();
    // mirrors of the parent tree: visited set (= its visit order) and frontier ("expandables");
    // all readbacks are sparse, nothing of size n crosses the bus
    std::vector<char> visited(n, 0), frontier(n, 0);
    std::vector<uint32_t> order(n);
    uint32_t visited_count = 0;
    check(imomd_get_visit_order(planner_, 0, parent_tree, order.data(), static_cast<uint32_t>(n),
                                &visited_count),
          "visit order readback");
    for (uint32_t i = 0; i < visited_count; ++i) visited[order[i]] = 1;
    {
        std::vector<uint32_t> fr(n);
        uint32_t cnt = 0;
        check(imomd_get_frontier(planner_, 0, parent_tree, fr.data(), static_cast<uint32_t>(n), &cnt),
              "frontier readback");
        for (uint32_t i = 0; i < cnt; ++i) frontier[fr[i]] = 1;
    }
    // the child tree's parent map, rebuilt with the reference's insertion history so that it
    // iterates in the same order (:860): {root, root} first, then every vertex as visited
    check(imomd_get_visit_order(planner_, 0, child_tree, order.data(), static_cast<uint32_t>(n),
                                &visited_count),
          "visit order readback");
    std::vector<uint32_t> cpar(visited_count);
    check(imomd_gather_tree(planner_, 0, child_tree, order.data(), visited_count, cpar.data(), nullptr),
          "tree readback");
    const size_t child_root = destinations_[child_tree];
    std::unordered_map<size_t, size_t> child_parent = {{child_root, child_root}};
    for (uint32_t i = 1; i < visited_count; ++i) child_parent[order[i]] = cpar[i];

    const int idx = setIndex(parent_tree, child_tree);
    std::unordered_set<int> merged;
    StageTimer timer2("  mergeTwoTrees after readback");
    // main branches: connection node ----> root of the child tree
    // Which vertices a chain attaches follows from the child's parent map and the visited
    // mirror alone, so each chain is ONE device launch; the set insertions its attaches made
    // are applied, in order, before the set iterator advances (the state it then sees is the
    // one upstream's in-loop insertions produce).
    std::vector<uint32_t> chain;
    for (size_t connection_node : connection_sets_[idx])
    {
        size_t current = connection_node;
        bool optimal = true;
        chain.clear();
        while (current != child_root && optimal)
        {
            const size_t x_new = child_parent[current];
            if (visited[x_new])
            {
                if (merged.count(static_cast<int>(x_new))) optimal = false;
            }
            else
            {
                chain.push_back(static_cast<uint32_t>(x_new));
                visited[x_new] = 1;
                frontier[x_new] = 0;
                for (const auto& kv : (*connection_)[x_new])
                    if (!visited[kv.first]) frontier[kv.first] = 1;
                merged.insert(static_cast<int>(x_new));
            }
            current = x_new;
        }
        attachList(parent_tree, chain);
    }
    // side branches, in the iteration order of the child's parent map.  Nothing in this loop
    // reads what the attaches compute (each vertex is adjacent to a visited one, so it always
    // connects; the visited set and the frontier follow from the adjacency alone), so the
    // whole sequence is worked out here and performed by ONE device launch.
    std::vector<uint32_t> todo;
    for (const auto& node : child_parent)
    {
        if (visited[node.first]) continue;
        size_t current = node.first;
        std::vector<size_t> branch = {current};
        while (!frontier[current])
        {
            current = child_parent[current];
            branch.push_back(current);
        }
        for (auto it = branch.rbegin(); it != branch.rend(); ++it)
        {
            const size_t x = *it;
            if (visited[x]) continue;
            todo.push_back(static_cast<uint32_t>(x));
            visited[x] = 1;
            frontier[x] = 0;
            for (const auto& kv : (*connection_)[x])
                if (!visited[kv.first]) frontier[kv.first] = 1;
        }
    }
    attachList(parent_tree, todo);
    // connection nodes of the child towards the other trees now belong to the parent
    for (int t = 0; t < K_; ++t)
    {
        if (t == child_tree || t == parent_tree) continue;
        const int from = setIndex(child_tree, t), to = setIndex(parent_tree, t);
        for (size_t node : connection_sets_[from])
        {
            if (!visited[node])
            {
                connection_sets_[to].insert(node);
                check(imomd_set_contact(planner_, 0, to), "contact flag");
            }
        }
    }
    check(imomd_set_tree_done(planner_, 0, child_tree, 1), "tree flag");
}

// mergePseudoTrees_, src/imomd_rrt_star.cpp:744-817 upstream
void ImomdRRT::mergePseudoTrees()
{
    StageTimer timer("mergePseudoTrees");
    const int source = 0, target = K_ - 1;
    std::unordered_set<int> unmerged;
    for (int i = 1; i < K_ - 1; ++i) unmerged.insert(i);
    int pick = -1;
    while (!unmerged.empty())
    {
        bool progress = false;
        size_t most = 0;
        for (int id : unmerged)
        {
            const size_t c = connection_sets_[id * (id - 1) / 2].size();
            if (c > most)
            {
                most = c;
                pick = id;
            }
        }
        if (most > 0)
        {
            mergeTwoTrees(source, pick);
            unmerged.erase(pick);
            progress = true;
        }
        most = 0;
        for (int id : unmerged)
        {
            const size_t c = connection_sets_[target * (target - 1) / 2 + id].size();
            if (c > most)
            {
                most = c;
                pick = id;
            }
        }
        if (most > 0)
        {
            mergeTwoTrees(target, pick);
            unmerged.erase(pick);
            progress = true;
        }
        if (!progress)
            throw std::runtime_error("ImomdRRT (B200): a pseudo tree touches neither the source nor the "
                                     "target tree (the reference loops forever here)");
    }
    // best direct connection between the merged source and target trees (:797-815; the
    // threshold stays the initial distance, as upstream)
    const auto& direct = connection_sets_[target * (target - 1) / 2];
    std::vector<uint32_t> nodes;
    for (size_t node : direct) nodes.push_back(static_cast<uint32_t>(node));
    std::vector<double> cost_s(nodes.size()), cost_t(nodes.size());
    check(imomd_gather_tree(planner_, 0, source, nodes.data(), static_cast<uint32_t>(nodes.size()), nullptr,
                            cost_s.data()),
          "tree readback");
    check(imomd_gather_tree(planner_, 0, target, nodes.data(), static_cast<uint32_t>(nodes.size()), nullptr,
                            cost_t.data()),
          "tree readback");
    auto D = getDistanceMatrix();
    const double threshold = (*D)[source][target];
    for (size_t i = 0; i < nodes.size(); ++i)
    {
        const size_t node = nodes[i];
        const double distance = cost_s[i] + cost_t[i];
        if (distance < threshold)
        {
            check(imomd_set_distance(planner_, 0, source, target, distance, static_cast<uint32_t>(node)),
                  "distance update");
            dm_updated_ = true;
        }
    }
    merge_done_ = true;
}

void ImomdRRT::run()
{
    bool done_iter = false, done_time = false, done_expansion = false;
    while (!done_iter && !done_time && !done_expansion)
    {
        // the reference solves the RTSP after the pass that starts with
        // iteration_count_ % 100 == 0 and stops after the pass that starts with
        // iteration_count_ > max_iter: batch up to whichever comes first
        long long to_rtsp = (iteration_count_ % 100 == 0) ? 1 : 101 - (iteration_count_ % 100);
        long long to_stop = static_cast<long long>(setting_.max_iter) + 2 - iteration_count_;
        long long passes = std::max(1LL, std::min(to_rtsp, to_stop));
        const long long before = iteration_count_;
        expandBatch(static_cast<int>(passes));
        if ((iteration_count_ - 1) % 100 == 0)
        {
            solveRTSP();
            std::vector<int32_t> done(K_);
            check(imomd_get_state(planner_, 0, nullptr, nullptr, nullptr, nullptr, nullptr, done.data(),
                                  nullptr),
                  "state readback");
            done_expansion = std::all_of(done.begin(), done.end(), [](int32_t d) { return d != 0; });
        }
        if (iteration_count_ - 1 > setting_.max_iter)
            done_iter = true;
        else if (std::chrono::duration<double>(std::chrono::steady_clock::now() - start_).count() >
                 setting_.max_time)
            done_time = true;   // (the device stopped the batch after the pass that crossed max_time)
        if (iteration_count_ == before) break;   // defensive: the device made no progress
    }
    if (done_iter) std::cout << "[IMOMD] Reached Max Iteration" << std::endl;
    else if (done_time) std::cout << "[IMOMD] Reached Max Time" << std::endl;
    if (done_expansion) std::cout << "[IMOMD] Tree Exploration Done" << std::endl;
    for (int i = 0; i < 10; ++i) solveRTSP();

    pthread_mutex_lock(&lock_);
    finished_ = true;
    pthread_mutex_unlock(&lock_);
    pthread_cond_signal(&cond_);
}

void* ImomdRRT::findShortestPath(void* arg)
{
    // An exception cannot leave a pthread entry point: keep it for failure() / rethrowFailure(),
    // report it and release a printPath thread that may be waiting.
    ImomdRRT* self = static_cast<ImomdRRT*>(arg);
    try
    {
        self->run();
    }
    catch (const std::exception& ex)
    {
        self->failure_ = std::current_exception();
        std::cerr << "[IMOMD] planner stopped: " << ex.what() << std::endl;
        pthread_mutex_lock(&self->lock_);
        self->finished_ = true;
        pthread_mutex_unlock(&self->lock_);
        pthread_cond_signal(&self->cond_);
    }
    pthread_exit(NULL);
}

void* ImomdRRT::printPath(void* arg)
{
    ImomdRRT* self = static_cast<ImomdRRT*>(arg);
    int count = 0;
    while (!self->finished_)
    {
        pthread_mutex_lock(&self->lock_);
        while (!self->path_updated_ && !self->finished_) pthread_cond_wait(&self->cond_, &self->lock_);
        std::cout << "[main] print path: " << count++ << std::endl;
        for (size_t node : self->shortest_path_) std::cout << node << " -> ";
        std::cout << "#" << std::endl;
        self->path_updated_ = false;
        pthread_mutex_unlock(&self->lock_);
    }
    pthread_exit(NULL);
}

std::shared_ptr<std::vector<std::vector<double>>> ImomdRRT::getDistanceMatrix()
{
    std::vector<double> flat(static_cast<size_t>(K_) * K_);
    check(imomd_get_state(planner_, 0, flat.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr),
          "distance matrix readback");
    auto m = std::make_shared<std::vector<std::vector<double>>>(K_, std::vector<double>(K_));
    for (int i = 0; i < K_; ++i)
        for (int j = 0; j < K_; ++j) (*m)[i][j] = flat[static_cast<size_t>(i) * K_ + j];
    return m;
}

void ImomdRRT::solveRTSP()
{
    if (!(connected_ && dm_updated_)) return;
    if (setting_.pseudo_mode)
    {
        // :990-1001 upstream: merge once, then the path is source -> target through the merged trees
        if (!merge_done_)
        {
            mergePseudoTrees();
            sequence_ = std::make_shared<std::vector<int>>(std::vector<int>{0, K_ - 1});
        }
        shortest_path_cost_ = (*getDistanceMatrix())[0][K_ - 1];
        updatePath();
        logData();
        return;
    }
    auto result = solver_.solveRTSP(getDistanceMatrix(), 0, K_ - 1);
    if (result.first < shortest_path_cost_)
    {
        sequence_ = result.second;
        shortest_path_cost_ = result.first;
        updatePath();
        logData();
    }
}

// Stitches the tree branches along the visiting sequence into one vertex path.
void ImomdRRT::updatePath()
{
    StageTimer timer("updatePath");
    std::vector<uint32_t> cn(static_cast<size_t>(K_) * K_);
    check(imomd_get_state(planner_, 0, nullptr, cn.data(), nullptr, nullptr, nullptr, nullptr, nullptr),
          "connection nodes readback");
    std::vector<size_t> path;
    std::vector<uint32_t> buf(1 << 16);
    auto walk = [&](int tree, uint32_t from) {
        uint32_t len = 0;
        for (;;)
        {
            check(imomd_walk_to_root(planner_, 0, tree, from, buf.data(), static_cast<uint32_t>(buf.size()),
                                     &len),
                  "path walk");
            if (len <= buf.size()) break;
            buf.resize(len);
        }
        return len;
    };
    for (size_t s = 0; s + 1 < sequence_->size(); ++s)
    {
        const int cur = (*sequence_)[s], nxt = (*sequence_)[s + 1];
        const uint32_t meet = cn[static_cast<size_t>(cur) * K_ + nxt];
        // root of `cur` ... up to (excluding) the meeting vertex
        uint32_t len = walk(cur, meet);
        for (uint32_t i = len; i-- > 1;) path.push_back((*map_)[buf[i]].id);
        // the meeting vertex ... up to (excluding) the root of `nxt`
        len = walk(nxt, meet);
        for (uint32_t i = 0; i + 1 < len; ++i) path.push_back((*map_)[buf[i]].id);
    }
    path.push_back(destinations_.back());

    pthread_mutex_lock(&lock_);
    shortest_path_.swap(path);
    path_updated_ = true;
    dm_updated_ = false;
    pthread_mutex_unlock(&lock_);
    check(imomd_clear_dm_updated(planner_, 0), "flag reset");
    pthread_cond_signal(&cond_);
}

void ImomdRRT::logData()
{
    const double elapsed =
        std::chrono::duration<double>(std::chrono::steady_clock::now() - start_).count();
    std::cout << "Elapsed time[s]: " << std::setw(10) << elapsed << " | "
              << "Path Cost[m]: " << std::setw(10) << shortest_path_cost_ << " | "
              << "Tree Size: " << std::setw(10) << tree_vertices_ << std::endl;
    log_.push_back({elapsed, shortest_path_cost_, tree_vertices_, iteration_count_ - 1});
    if (setting_.log_data && csv_.good())
    {
        // three quoted fields like the reference's CSVFile (utils/csv.h:85-95); tree_size is
        // an int there, so it prints without decimals
        csv_ << '"' << fixed4(elapsed) << "\";\"" << fixed4(shortest_path_cost_) << "\";\""
             << tree_vertices_ << "\";" << std::endl;
    }
}

std::vector<size_t> ImomdRRT::getShortestPath()
{
    pthread_mutex_lock(&lock_);
    std::vector<size_t> copy = shortest_path_;
    pthread_mutex_unlock(&lock_);
    return copy;
}

double ImomdRRT::getShortestPathCost() { return shortest_path_cost_; }

std::vector<int> ImomdRRT::getVisitingSequence()
{
    return sequence_ ? *sequence_ : std::vector<int>{};
}
