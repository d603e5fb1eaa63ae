// csrc/imomd_setup.h -- host-side sizing and validation shared by the C ABI glue
// (imomd_kernels.cu) and the test-only host simulation (tests/hostsim).
#ifndef IMOMD_SETUP_H
#define IMOMD_SETUP_H

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "imomd_layout.h"

namespace imomd {

// Adjacency must be a simple symmetric graph with degree <= 13: the container-order
// rule of the children sets is only defined up to 13 elements, and
// the reference's `(*connection_ptr_)[a][b]` lookups assume both directions exist.
inline bool validate_graph(uint32_t n, const uint32_t* row_ptr, const uint32_t* col,
                           uint32_t* max_degree, std::string* err)
{
    char buf[160];
    uint32_t md = 0;
    for (uint32_t v = 0; v < n; ++v)
    {
        if (row_ptr[v + 1] < row_ptr[v])
        {
            std::snprintf(buf, sizeof(buf), "row_ptr not monotone at %u", v);
            *err = buf;
            return false;
        }
        md = std::max(md, row_ptr[v + 1] - row_ptr[v]);
        for (uint32_t e = row_ptr[v]; e < row_ptr[v + 1]; ++e)
        {
            if (col[e] >= n)
            {
                std::snprintf(buf, sizeof(buf), "neighbour id out of range at arc %u", e);
                *err = buf;
                return false;
            }
            if (col[e] == v)
            {
                std::snprintf(buf, sizeof(buf), "self loop at vertex %u", v);
                *err = buf;
                return false;
            }
        }
    }
    if (md > IMOMD_MAXDEG)
    {
        std::snprintf(buf, sizeof(buf), "max degree %u exceeds %d: children order undefined", md,
                      IMOMD_MAXDEG);
        *err = buf;
        return false;
    }
    for (uint32_t v = 0; v < n; ++v)
    {
        for (uint32_t e = row_ptr[v]; e < row_ptr[v + 1]; ++e)
        {
            uint32_t u = col[e];
            bool found = false;
            for (uint32_t f = row_ptr[u]; f < row_ptr[u + 1]; ++f)
            {
                if (col[f] == v)
                {
                    found = true;
                    break;
                }
            }
            if (!found)
            {
                std::snprintf(buf, sizeof(buf), "arc %u->%u has no reverse arc", v, u);
                *err = buf;
                return false;
            }
        }
    }
    *max_degree = md;
    return true;
}

// Geometry of the frontier hash: G x G cells over the bounding box, about 256
// vertices per cell, G a power of two in [4, 64] unless forced.
inline void setup_grid(GraphDev& d, uint32_t n, const double* lat, const double* lon, int forced_G,
                       uint32_t max_degree, std::vector<double>* cosrow)
{
    d.width = 4;
    while ((uint32_t)d.width < max_degree) d.width *= 2;
    double lat_min = lat[0], lat_max = lat[0], lon_min = lon[0], lon_max = lon[0];
    for (uint32_t v = 1; v < n; ++v)
    {
        lat_min = std::min(lat_min, lat[v]);
        lat_max = std::max(lat_max, lat[v]);
        lon_min = std::min(lon_min, lon[v]);
        lon_max = std::max(lon_max, lon[v]);
    }
    int G = 4;
    while (G < 64 && (double)G * G * 256.0 < (double)n) G *= 2;
    if (forced_G > 0) G = std::min(forced_G, 256);
    double span_lat = std::max(lat_max - lat_min, 1e-6);
    double span_lon = std::max(lon_max - lon_min, 1e-6);
    d.G = G;
    d.lat0 = lat_min;
    d.lon0 = lon_min;
    d.hlat = span_lat / G * (1.0 + 1e-9);
    d.hlon = span_lon / G * (1.0 + 1e-9);
    // the cell bound needs |lat| < 90 and longitude differences below 180 degrees
    d.prune = (lat_min > -89.0 && lat_max < 89.0 && span_lon < 170.0) ? 1 : 0;
    const double d2r = 3.14159265358979323846 / 180;
    cosrow->resize(2 * (size_t)G);
    for (int r = 0; r < G; ++r)
    {
        double lo = (d.lat0 + r * d.hlat - 1e-9) * d2r;
        double hi = (d.lat0 + (r + 1) * d.hlat + 1e-9) * d2r;
        (*cosrow)[r] = std::max(0.0, std::min(std::cos(lo), std::cos(hi)) * (1.0 - 1e-12));
        double mx = (lo <= 0.0 && hi >= 0.0) ? 1.0 : std::max(std::cos(lo), std::cos(hi));
        (*cosrow)[G + r] = std::min(1.0, mx * (1.0 + 1e-12));
    }
}

// CSR -> padded rows (GraphDev::col / w / deg)
inline void build_padded_rows(uint32_t n, const uint32_t* row_ptr, const uint32_t* col, const double* w,
                              int width, std::vector<uint32_t>* pcol, std::vector<double>* pw,
                              std::vector<uint8_t>* deg)
{
    pcol->assign((size_t)n * width, IMOMD_NONE);
    pw->assign((size_t)n * width, 0.0);
    deg->assign(n, 0);
    for (uint32_t v = 0; v < n; ++v)
    {
        const uint32_t d = row_ptr[v + 1] - row_ptr[v];
        (*deg)[v] = (uint8_t)d;
        for (uint32_t j = 0; j < d; ++j)
        {
            (*pcol)[(size_t)v * width + j] = col[row_ptr[v] + j];
            (*pw)[(size_t)v * width + j] = w[row_ptr[v] + j];
        }
    }
}

inline void default_sizes(PlannerDev& d, uint32_t frontier_chunks, uint32_t arena_entries,
                          uint32_t log_entries)
{
    const size_t n = d.g.n;
    const size_t cells = (size_t)d.g.G * d.g.G;
    // a tree's frontier needs at most (occupied cells + entries / 32) chunks; the default
    // covers frontiers of up to ~N/8 vertices (the run stops with FRONTIER_FULL beyond it)
    size_t chunks = frontier_chunks ? frontier_chunks : cells + n / 256 + 64;
    d.chunks = (int32_t)chunks;
    // rewire scratch: update lists are bounded by the tree size, frames by the depth
    d.arena_entries = arena_entries
                          ? arena_entries
                          : (uint32_t)std::min<size_t>(std::max<size_t>(1u << 16, n), 1u << 28);
    d.log_entries = log_entries ? log_entries : (d.pseudo ? (uint32_t)std::max<size_t>(4 * n, 1024) : 16u);
}

// header of a freshly constructed query (src/imomd_rrt_star.cpp:57-71, :103-108)
inline void fill_initial_header(QueryDev& h, int K, const uint32_t* dest, const double* prob)
{
    std::memset(&h, 0, sizeof(h));
    h.K = K;
    for (int i = 0; i < K; ++i) h.dest[i] = dest[i];
    for (int i = 0; i < K; ++i)
    {
        for (int j = 0; j < K; ++j)
        {
            h.P[i * K + j] = prob[(size_t)i * K + j];
            h.D[i * K + j] = (i == j) ? 0.0 : INFINITY;
            h.MHv[i * K + j] = INFINITY;
            h.MHi[i * K + j] = IMOMD_NONE;
            h.CN[i * K + j] = IMOMD_NONE;
        }
    }
}

}  // namespace imomd

#endif  // IMOMD_SETUP_H
