// host/src/eci_gen_solver.cpp -- ECI-Gen relaxed-TSP solver of the host facade.
//
// Reference behaviour (src/eci_gen_tsp_solver.cpp upstream): shortest chain source->target
// over the K x K pseudo-complete matrix (:207-260), enhanced cheapest insertion of the
// remaining destinations with five insertion shapes (:103-205, :262-426), removal of
// repeated visits (:827-882), then a genetic stage: `ga_mutation_iter` four-cut mutations of
// the tour seed a population (:660-788) and `ga_generation` rounds of `ga_population`
// roulette-selected two-cut crossovers refine it (:502-658); a candidate is kept when its
// cost is within 10 % of the incumbent and its hash is new.
//
// Here the genetic stage is batch-shaped: a whole seeding pass / generation of candidates
// is produced on the host (all random draws happen in the reference's order and do not
// depend on the costs), their costs come back from ONE device launch (imomd_ga_eval, one
// warp per tour, left-to-right sum = bit-identical to calculatePathCost_ :442-450), and the
// acceptance rule is then applied in candidate order.  Same tours, same costs, same
// incumbent as the sequential loop.
#include "imomd_rrt_star/eci_gen_tsp_solver.h"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <queue>
#include <stdexcept>
#include <string>
#include <unordered_set>

#include "imomd_b200.h"

namespace {

// size_t-wide hash of a tour (:801-810): `i + 0x9e3779b9` wraps in 32 bits first
size_t tourHash(const std::vector<int>& tour)
{
    size_t h = tour.size();
    for (int v : tour)
    {
        const unsigned salted = static_cast<unsigned>(v) + 0x9e3779b9u;
        h ^= salted + (h << 6) + (h >> 2);
    }
    return h;
}

enum class Shape { Revisit, Between, SwapLeft, SwapRight, SwapBoth };

struct Placement
{
    double extra = INFINITY;
    int at = -1;
    Shape shape = Shape::Revisit;
};

}  // namespace

EciGenSolver::EciGenSolver(const rtsp_setting_t& setting, int device)
    : setting_(setting), device_(device)
{
    if (setting_.ga_random_seed)
    {
        std::random_device rd;
        rng_ = std::mt19937{rd()};
    }
    else
    {
        rng_ = std::mt19937{0};
    }
}

EciGenSolver::~EciGenSolver()
{
    if (ctx_) imomd_ctx_destroy(ctx_);
}

EciGenSolver::EciGenSolver(EciGenSolver&& o) noexcept { *this = std::move(o); }

EciGenSolver& EciGenSolver::operator=(EciGenSolver&& o) noexcept
{
    if (this != &o)
    {
        if (ctx_) imomd_ctx_destroy(ctx_);
        setting_ = o.setting_;
        rng_ = o.rng_;
        ctx_ = o.ctx_;
        device_ = o.device_;
        o.ctx_ = nullptr;
    }
    return *this;
}

int EciGenSolver::drawInt(int lo, int hi) { return std::uniform_int_distribution<>(lo, hi)(rng_); }

double EciGenSolver::drawReal(double lo, double hi)
{
    return std::uniform_real_distribution<>(lo, hi)(rng_);
}

// Costs of a batch of tours, computed on the GPU.
void EciGenSolver::evaluate(const std::vector<Tour>& tours, std::vector<double>& costs)
{
    costs.assign(tours.size(), 0.0);
    if (tours.empty()) return;
    if (!ctx_)
    {
        if (imomd_ctx_create(device_, &ctx_) != IMOMD_OK)
            throw std::runtime_error(std::string("EciGenSolver: no CUDA context: ") +
                                     imomd_last_error());
    }
    size_t longest = 1;
    for (const Tour& t : tours) longest = std::max(longest, t.size());
    const int stride = static_cast<int>(longest);
    std::vector<int32_t> flat(tours.size() * longest, 0), lens(tours.size());
    for (size_t i = 0; i < tours.size(); ++i)
    {
        lens[i] = static_cast<int32_t>(tours[i].size());
        std::copy(tours[i].begin(), tours[i].end(), flat.begin() + i * longest);
    }
    if (imomd_ga_eval(ctx_, K_, D_.data(), static_cast<int>(tours.size()), stride, flat.data(),
                      lens.data(), costs.data()) != IMOMD_OK)
        throw std::runtime_error(std::string("EciGenSolver: imomd_ga_eval failed: ") +
                                 imomd_last_error());
    evaluated_ += tours.size();
}

// Label-correcting search with a max-heap of (-cost, id) and no closed set, like upstream;
// returns the chain cost and the chain source..target.
double EciGenSolver::shortestChain(Tour& chain) const
{
    std::vector<int> via(K_, -1);
    std::vector<double> reach(K_, INFINITY);
    via[source_] = source_;
    reach[source_] = 0;
    std::priority_queue<std::pair<double, int>> open;
    open.push({0, source_});
    while (!open.empty())
    {
        const int at = open.top().second;
        open.pop();
        if (at == target_) break;
        for (int next = 0; next < K_; ++next)
        {
            if (next == at) continue;
            const double through = reach[at] + d(at, next);
            if (through < reach[next])
            {
                reach[next] = through;
                via[next] = at;
                open.push({-through, next});
            }
        }
    }
    chain.assign(1, target_);
    for (int at = target_; at != source_;)
    {
        at = via[at];
        chain.push_back(at);
    }
    std::reverse(chain.begin(), chain.end());
    return reach[target_];
}

void EciGenSolver::cheapestInsertion()
{
    Tour tour;
    double cost = shortestChain(tour);
    // Candidates are visited in the order upstream's `std::unordered_set<int> not_visited`
    // iterates (:112-121): exact cost ties between different destinations go to the first one
    // met, and symmetric maps do produce such ties.  Same libstdc++ here, so the real container
    // is used (filled 0..K-1, then the chain's members erased).
    std::unordered_set<int> waiting;
    for (int v = 0; v < K_; ++v) waiting.insert(v);
    for (int v : tour) waiting.erase(v);
    auto cheapest = [&](int v) {
        Placement best;
        const int n = static_cast<int>(tour.size());
        for (int i = 0; i < n; ++i)
        {
            const double revisit = 2 * d(v, tour[i]);
            double between = INFINITY, left = INFINITY, right = INFINITY, both = INFINITY;
            if (i == n - 1)
            {
                between = revisit;
            }
            else
            {
                const int a = tour[i], b = tour[i + 1];
                between = d(a, v) + d(v, b) - d(a, b);
                if (setting_.swapping)
                {
                    const bool room_left = i > 1, room_right = n - (i + 1) > 2;
                    if (room_left)
                    {
                        const int p = tour[i - 1], pp = tour[i - 2];
                        left = d(v, b) - d(a, b) + d(p, v) + d(pp, a) - d(pp, p);
                    }
                    if (room_right)
                    {
                        const int c = tour[i + 2], cc = tour[i + 3];
                        right = d(a, v) - d(a, b) + d(c, v) + d(cc, b) - d(cc, c);
                    }
                    if (room_left && room_right)
                    {
                        const int p = tour[i - 1], pp = tour[i - 2];
                        const int c = tour[i + 2], cc = tour[i + 3];
                        both = d(p, v) + d(pp, a) - d(pp, p) - d(a, b) + d(c, v) + d(cc, b) - d(cc, c);
                    }
                }
            }
            auto consider = [&](double extra, Shape s) {
                if (extra < best.extra)
                {
                    best.extra = extra;
                    best.at = i;
                    best.shape = s;
                }
            };
            consider(revisit, Shape::Revisit);
            consider(between, Shape::Between);
            if (setting_.swapping)
            {
                consider(left, Shape::SwapLeft);
                consider(right, Shape::SwapRight);
                consider(both, Shape::SwapBoth);
            }
        }
        return best;
    };
    while (!waiting.empty())
    {
        Placement pick;
        int who = -1;
        for (int v : waiting)
        {
            const Placement p = cheapest(v);
            if (p.extra < pick.extra)
            {
                pick = p;
                who = v;
            }
        }
        if (who < 0) throw std::runtime_error("EciGenSolver: distance matrix is not connected");
        const int i = pick.at;
        switch (pick.shape)
        {
            case Shape::Revisit:
            {
                const int back = tour[i];
                tour.insert(tour.begin() + i + 1, {who, back});
                break;
            }
            case Shape::Between:
                tour.insert(tour.begin() + i + 1, who);
                break;
            case Shape::SwapLeft:
                std::swap(tour[i], tour[i - 1]);
                tour.insert(tour.begin() + i + 1, who);
                break;
            case Shape::SwapRight:
                std::swap(tour[i + 1], tour[i + 2]);
                tour.insert(tour.begin() + i + 1, who);
                break;
            case Shape::SwapBoth:
                std::swap(tour[i], tour[i - 1]);
                std::swap(tour[i + 1], tour[i + 2]);
                tour.insert(tour.begin() + i + 1, who);
                break;
        }
        cost += pick.extra;
        waiting.erase(who);
    }
    best_cost_ = cost;
    best_ = tour;
    if (static_cast<int>(best_.size()) > K_) dropRepeats(best_);
}

// Removes repeated visits where the matrix allows a shortcut (refineSequence_ upstream).
void EciGenSolver::dropRepeats(Tour& tour) const
{
    const int n = static_cast<int>(tour.size());
    std::vector<int> last_at(K_, 0);
    for (int i = 0; i < n; ++i) last_at[tour[i]] = i;
    std::vector<char> seen(K_, 0);
    Tour out;
    out.reserve(tour.size());
    for (int i = 0; i != n - 1; ++i)
    {
        const int v = tour[i];
        if (!seen[v])
        {
            out.push_back(v);
            seen[v] = 1;
            continue;
        }
        bool jumped = false;
        int j = i + 1;
        while (std::isinf(d(out.back(), tour[j])) && j != n - 1 &&
               (seen[tour[j]] || j < last_at[tour[j]]))
        {
            ++j;
            if (!seen[tour[j]] && !std::isinf(d(out.back(), tour[j])))
            {
                out.push_back(tour[j]);
                seen[tour[j]] = 1;
                jumped = true;
                i = --j;
            }
        }
        if (!jumped && out.back() != tour[i])
        {
            out.push_back(tour[i]);
            seen[tour[i]] = 1;
        }
    }
    if (out.back() != tour.back()) out.push_back(tour.back());
    tour.swap(out);
}

// Four-cut mutations of the incumbent (generateChromosomes_ upstream), costs on the GPU.
void EciGenSolver::seedPopulation()
{
    const Tour base = best_;
    const int hi = static_cast<int>(base.size()) - 1;
    std::unordered_set<size_t> known;
    std::vector<Tour> fresh;
    for (int it = 0; it < setting_.ga_mutation_iter; ++it)
    {
        int cut[4];
        for (int& c : cut) c = drawInt(1, hi);
        std::sort(cut, cut + 4);
        const int flip = drawInt(0, 7);
        const int order = drawInt(0, 4);
        // segments: [0,c0) [c0,c1) [c1,c2) [c2,c3) [c3,end); the middle three may be reversed
        // and are re-ordered, the outer two stay
        struct Seg
        {
            int lo, hi;
            bool rev;
        };
        Seg s2{cut[0], cut[1], flip == 1 || flip == 4 || flip == 6 || flip == 7};
        Seg s3{cut[1], cut[2], flip == 2 || flip == 4 || flip == 5 || flip == 7};
        Seg s4{cut[2], cut[3], flip == 3 || flip == 5 || flip == 6 || flip == 7};
        Seg mid[3];
        switch (order)
        {
            case 0: mid[0] = s2; mid[1] = s4; mid[2] = s3; break;
            case 1: mid[0] = s3; mid[1] = s2; mid[2] = s4; break;
            case 2: mid[0] = s3; mid[1] = s4; mid[2] = s2; break;
            case 3: mid[0] = s4; mid[1] = s2; mid[2] = s3; break;
            default: mid[0] = s4; mid[1] = s3; mid[2] = s2; break;
        }
        Tour t;
        t.reserve(base.size());
        t.insert(t.end(), base.begin(), base.begin() + cut[0]);
        for (const Seg& s : mid)
        {
            if (s.rev)
                t.insert(t.end(), base.rend() - s.hi, base.rend() - s.lo);
            else
                t.insert(t.end(), base.begin() + s.lo, base.begin() + s.hi);
        }
        t.insert(t.end(), base.begin() + cut[3], base.end());
        if (static_cast<int>(t.size()) > K_) dropRepeats(t);
        if (!known.insert(tourHash(t)).second) continue;
        fresh.push_back(std::move(t));
    }
    std::vector<double> costs;
    evaluate(fresh, costs);
    const double admit = best_cost_ * 1.1;
    double lowest = best_cost_;
    int lowest_at = 0, kept = 0;
    for (size_t i = 0; i < fresh.size(); ++i)
    {
        if (costs[i] <= admit)
        {
            fitness_.push_back(1 / costs[i]);
            pool_.push_back(fresh[i]);
            if (costs[i] < lowest)
            {
                lowest = costs[i];
                lowest_at = kept;
            }
            ++kept;
        }
    }
    if (!pool_.empty())
    {
        best_cost_ = lowest;
        best_ = pool_[lowest_at];
    }
}

// Roulette-selected two-cut crossovers (solveGeneticAlgorithm_ upstream), one device
// launch per generation.
void EciGenSolver::evolve()
{
    for (int gen = 0; gen < setting_.ga_generation && pool_.size() >= 2; ++gen) evolveOnce();
}

bool EciGenSolver::evolveOnce()
{
    if (pool_.size() < 2) return false;
    {
        const double total = std::accumulate(fitness_.begin(), fitness_.end(), 0.0);
        auto spin = [&]() -> const Tour& {
            double r = drawReal(0.0, total);
            size_t i = 0;
            while (r >= 0 && i < fitness_.size()) r -= fitness_[i++];
            return pool_[i - 1];
        };
        std::unordered_set<size_t> known;
        std::vector<Tour> fresh;
        std::vector<char> in_slice(K_);
        for (int n = 0; n < setting_.ga_population; ++n)
        {
            const Tour& a = spin();
            const Tour& b = spin();
            int s1 = drawInt(1, static_cast<int>(a.size()) - 1);
            int s2 = drawInt(1, static_cast<int>(a.size()) - 1);
            if (s2 < s1) std::swap(s1, s2);
            std::fill(in_slice.begin(), in_slice.end(), 0);
            for (int i = s1; i < s2; ++i) in_slice[a[i]] = 1;
            const int shift = drawInt(1 - s1, static_cast<int>(std::max(a.size(), b.size())) - s2 - 1);
            // child = head of b (without a's slice), a's slice (maybe reversed), tail of b
            Tour child;
            child.push_back(source_);
            const int b_end = static_cast<int>(b.size()) - 1;
            int pos = 1, bi = 1;
            while (pos < shift + s1 && bi < b_end)
            {
                if (!in_slice[b[bi]])
                {
                    child.push_back(b[bi]);
                    ++pos;
                }
                ++bi;
            }
            if (drawInt(0, 1))
                child.insert(child.end(), a.begin() + s1, a.begin() + s2);
            else
                child.insert(child.end(), a.rend() - s2, a.rend() - s1);
            while (bi < b_end)
            {
                if (!in_slice[b[bi]]) child.push_back(b[bi]);
                ++bi;
            }
            if (child.back() != target_) child.push_back(target_);
            if (static_cast<int>(child.size()) > K_) dropRepeats(child);
            if (!known.insert(tourHash(child)).second) continue;
            fresh.push_back(std::move(child));
        }
        std::vector<double> costs;
        evaluate(fresh, costs);
        const double admit = best_cost_ * 1.1;
        double lowest = best_cost_;
        int lowest_at = 0, kept = 0;
        std::vector<Tour> next;
        std::vector<double> next_fit;
        for (size_t i = 0; i < fresh.size(); ++i)
        {
            if (costs[i] <= admit)
            {
                next_fit.push_back(1 / costs[i]);
                next.push_back(std::move(fresh[i]));
                if (costs[i] < lowest)
                {
                    lowest = costs[i];
                    lowest_at = kept;
                }
                ++kept;
            }
        }
        fitness_.swap(next_fit);
        pool_.swap(next);
        if (!pool_.empty())
        {
            best_cost_ = lowest;
            best_ = pool_[lowest_at];
        }
    }
    return true;
}

void EciGenSolver::islandBegin(std::shared_ptr<std::vector<std::vector<double>>> matrix,
                               int source_id, int target_id, unsigned seed)
{
    K_ = static_cast<int>(matrix->size());
    if (K_ < 2 || K_ > static_cast<int>(IMOMD_KMAX))
        throw std::invalid_argument("EciGenSolver: matrix size out of range");
    D_.resize(static_cast<size_t>(K_) * K_);
    for (int i = 0; i < K_; ++i)
        for (int j = 0; j < K_; ++j) D_[static_cast<size_t>(i) * K_ + j] = (*matrix)[i][j];
    source_ = source_id;
    target_ = target_id;
    rng_ = std::mt19937{seed};
    best_cost_ = 0;
    best_.clear();
    pool_.clear();
    fitness_.clear();
    evaluated_ = 0;
    cheapestInsertion();
    seedPopulation();
}

bool EciGenSolver::islandStep() { return evolveOnce(); }

void EciGenSolver::islandAccept(const std::vector<int>& tour, double cost)
{
    if (pool_.empty())
    {
        pool_.push_back(tour);
        fitness_.push_back(1 / cost);
    }
    else
    {
        size_t worst = 0;
        for (size_t i = 1; i < fitness_.size(); ++i)
            if (fitness_[i] < fitness_[worst]) worst = i;
        if (1 / cost > fitness_[worst])
        {
            pool_[worst] = tour;
            fitness_[worst] = 1 / cost;
        }
    }
    if (cost < best_cost_)
    {
        best_cost_ = cost;
        best_ = tour;
    }
}

std::pair<double, std::shared_ptr<std::vector<int>>> EciGenSolver::solveRTSP(
    std::shared_ptr<std::vector<std::vector<double>>> matrix, int source_id, int target_id)
{
    K_ = static_cast<int>(matrix->size());
    if (K_ < 2 || K_ > static_cast<int>(IMOMD_KMAX))
        throw std::invalid_argument("EciGenSolver: matrix size out of range");
    D_.resize(static_cast<size_t>(K_) * K_);
    for (int i = 0; i < K_; ++i)
        for (int j = 0; j < K_; ++j) D_[static_cast<size_t>(i) * K_ + j] = (*matrix)[i][j];
    source_ = source_id;
    target_ = target_id;
    best_cost_ = 0;
    best_.clear();
    pool_.clear();
    fitness_.clear();
    evaluated_ = 0;
    cheapestInsertion();
    if (setting_.genetic)
    {
        seedPopulation();
        if (pool_.size() >= 2) evolve();
    }
    return {best_cost_, std::make_shared<std::vector<int>>(best_)};
}

// solveDijkstra upstream (:55-74): the chain alone, no insertion, no GA
std::pair<double, std::shared_ptr<std::vector<int>>> EciGenSolver::solveDijkstra(
    std::shared_ptr<std::vector<std::vector<double>>> matrix, int source_id, int target_id)
{
    K_ = static_cast<int>(matrix->size());
    D_.resize(static_cast<size_t>(K_) * K_);
    for (int i = 0; i < K_; ++i)
        for (int j = 0; j < K_; ++j) D_[static_cast<size_t>(i) * K_ + j] = (*matrix)[i][j];
    source_ = source_id;
    target_ = target_id;
    best_.clear();
    pool_.clear();
    fitness_.clear();
    Tour chain;
    best_cost_ = shortestChain(chain);
    return {best_cost_, std::make_shared<std::vector<int>>(chain)};
}
