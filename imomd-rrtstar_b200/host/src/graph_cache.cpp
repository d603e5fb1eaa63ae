// host/src/graph_cache.cpp -- see imomd_rrt_star/graph_cache.h
#include "imomd_rrt_star/graph_cache.h"

#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>

namespace imomd_b200 {

namespace {

constexpr char kMagic[8] = {'I', 'M', 'O', 'M', 'D', 'C', 'S', 'R'};
constexpr uint32_t kVersion = 1;
constexpr size_t kMaxDegree = 13;   // libstdc++: 13 buckets until the 14th element

struct Header
{
    char magic[8];
    uint32_t version;
    uint32_t stdlib;      // __GLIBCXX__ of the writer (informational)
    uint64_t n;
    uint64_t arcs;
};

template <class T>
void put(std::ofstream& f, const std::vector<T>& v)
{
    f.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(T)));
}

template <class T>
void get(std::ifstream& f, std::vector<T>& v)
{
    f.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(T)));
    if (!f) throw std::runtime_error("GraphCache: truncated file");
}

}  // namespace

void GraphCache::save(const std::string& path, const Map& map, const Connection& connection)
{
    const size_t n = map.size();
    if (connection.size() != n) throw std::invalid_argument("GraphCache: map / connection sizes differ");
    std::vector<double> lat(n), lon(n), w;
    std::vector<uint64_t> row_ptr(n + 1);
    std::vector<uint32_t> col;
    for (size_t v = 0; v < n; ++v)
    {
        if (map[v].id != v) throw std::invalid_argument("GraphCache: location ids must equal their index");
        if (connection[v].size() > kMaxDegree)
            throw std::invalid_argument("GraphCache: degree above 13 cannot be replayed");
        lat[v] = map[v].latitude;
        lon[v] = map[v].longitude;
        row_ptr[v] = col.size();
        for (const auto& kv : connection[v])   // the container's own order
        {
            if (kv.first >= n) throw std::invalid_argument("GraphCache: neighbour id out of range");
            col.push_back(static_cast<uint32_t>(kv.first));
            w.push_back(kv.second);
        }
    }
    row_ptr[n] = col.size();
    Header h{};
    std::memcpy(h.magic, kMagic, 8);
    h.version = kVersion;
#ifdef __GLIBCXX__
    h.stdlib = static_cast<uint32_t>(__GLIBCXX__);
#endif
    h.n = n;
    h.arcs = col.size();
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f) throw std::runtime_error("GraphCache: cannot open " + path + " for writing");
    f.write(reinterpret_cast<const char*>(&h), sizeof(h));
    put(f, lat);
    put(f, lon);
    put(f, row_ptr);
    put(f, col);
    put(f, w);
    if (!f) throw std::runtime_error("GraphCache: write to " + path + " failed");
}

bool GraphCache::load(const std::string& path, std::shared_ptr<Map>& map,
                      std::shared_ptr<Connection>& connection)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    Header h{};
    f.read(reinterpret_cast<char*>(&h), sizeof(h));
    if (!f || std::memcmp(h.magic, kMagic, 8) != 0 || h.version != kVersion)
        throw std::runtime_error("GraphCache: " + path + " is not a graph cache of this version");
    if (h.n > 0xFFFFFFF0ull || h.arcs > h.n * kMaxDegree)
        throw std::runtime_error("GraphCache: implausible sizes in " + path);
    const size_t n = h.n;
    std::vector<double> lat(n), lon(n), w(h.arcs);
    std::vector<uint64_t> row_ptr(n + 1);
    std::vector<uint32_t> col(h.arcs);
    get(f, lat);
    get(f, lon);
    get(f, row_ptr);
    get(f, col);
    get(f, w);
    if (row_ptr[0] != 0 || row_ptr[n] != h.arcs) throw std::runtime_error("GraphCache: damaged row table");
    auto m = std::make_shared<Map>(n);
    auto c = std::make_shared<Connection>(n);
    for (size_t v = 0; v < n; ++v)
    {
        (*m)[v] = location_t(v, lat[v], lon[v]);
        const uint64_t e0 = row_ptr[v], e1 = row_ptr[v + 1];
        if (e1 < e0 || e1 - e0 > kMaxDegree || e1 > h.arcs)
            throw std::runtime_error("GraphCache: damaged row table");
        auto& row = (*c)[v];
        // libstdc++ puts a new key at the front of its bucket's run, or at the head of the
        // whole list when the bucket is empty: inserting in REVERSE iteration order rebuilds
        // the recorded order as long as no rehash happens (<= 13 elements).
        for (uint64_t e = e1; e > e0; --e)
        {
            if (col[e - 1] >= n) throw std::runtime_error("GraphCache: neighbour id out of range");
            row.insert({static_cast<size_t>(col[e - 1]), w[e - 1]});
        }
        uint64_t e = e0;
        for (const auto& kv : row)
        {
            if (e >= e1 || kv.first != col[e] || std::memcmp(&kv.second, &w[e], sizeof(double)) != 0)
                throw std::runtime_error(
                    "GraphCache: this standard library does not reproduce the recorded adjacency order; "
                    "parse the map again and rewrite the cache");
            ++e;
        }
        if (e != e1) throw std::runtime_error("GraphCache: duplicate neighbour in " + path);
    }
    map = m;
    connection = c;
    return true;
}

}  // namespace imomd_b200
