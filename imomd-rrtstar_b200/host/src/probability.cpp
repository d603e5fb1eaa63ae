// host/src/probability.cpp -- destination-selection probabilities of the planner
// constructor (reference: src/imomd_rrt_star.cpp:84-185).
//
// P[i][j] = chance that tree i steers towards destination j.  Two terms: an angular
// term (destinations with an empty sector around them, seen from i, are favoured; only
// with more than three destinations, :124) and an inverse-distance term; the matrix is
// then made symmetric and row-normalised in the reference's own (row-sequential) order,
// which is why the loops below must not be reordered.
#include <algorithm>
#include <utility>
#include <vector>

#include "geometry.h"
#include "host_api.h"

namespace imomd_host {

void selection_probabilities(int K, const double* lat, const double* lon, double* P)
{
    std::vector<std::vector<std::pair<double, int>>> by_bearing(K);
    std::vector<double> inv_dist((size_t)K * K, 0.0), inv_sum(K, 0.0);
    for (int i = 0; i < K * K; ++i) P[i] = 0.0;
    for (int i = 0; i < K; ++i)
    {
        for (int j = i + 1; j < K; ++j)
        {
            by_bearing[i].push_back({initial_bearing(lat[i], lon[i], lat[j], lon[j]), j});
            by_bearing[j].push_back({initial_bearing(lat[j], lon[j], lat[i], lon[i]), i});
            const double inv = 1.0 / great_circle_m(lat[i], lon[i], lat[j], lon[j]);
            inv_dist[(size_t)i * K + j] = inv;
            inv_dist[(size_t)j * K + i] = inv;
            inv_sum[i] += inv;
            inv_sum[j] += inv;
        }
    }
    if (K > 3)
    {
        for (int i = 0; i < K; ++i)
        {
            auto& ring = by_bearing[i];
            std::sort(ring.begin(), ring.end());
            const int m = K - 1;
            double total = 0.0;
            for (int s = 0; s < m; ++s)
            {
                const double here = ring[s].first;
                const double before = ring[(s - 2 + K) % m].first;
                const double after = ring[(s + 1) % m].first;
                double gap_before = here - before;
                gap_before += gap_before < 0 ? 2 * M_PI : 0;
                double gap_after = after - here;
                gap_after += gap_after < 0 ? 2 * M_PI : 0;
                const int j = ring[s].second;
                P[(size_t)i * K + j] = gap_before + gap_after;
                total += P[(size_t)i * K + j];
            }
            for (int j = 0; j < K; ++j) P[(size_t)i * K + j] /= total;
        }
    }
    for (int i = 0; i < K; ++i)
    {
        for (int j = 0; j < K; ++j) P[(size_t)i * K + j] += inv_dist[(size_t)i * K + j] / inv_sum[i];
    }
    std::vector<double> row_sum(K, 0.0);
    for (int i = 0; i < K; ++i)
    {
        for (int j = i + 1; j < K; ++j)
        {
            P[(size_t)i * K + j] += P[(size_t)j * K + i];
            P[(size_t)j * K + i] = P[(size_t)i * K + j];
            row_sum[i] += P[(size_t)i * K + j];
            row_sum[j] += P[(size_t)j * K + i];
        }
        for (int j = 0; j < K; ++j) P[(size_t)i * K + j] /= row_sum[i];
    }
}

}  // namespace imomd_host

extern "C" void imomd_host_selection_probabilities(int K, const double* lat, const double* lon,
                                                   double* P)
{
    imomd_host::selection_probabilities(K, lat, lon, P);
}

extern "C" double imomd_host_great_circle_m(double lat_a, double lon_a, double lat_b, double lon_b)
{
    return imomd_host::great_circle_m(lat_a, lon_a, lat_b, lon_b);
}
