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
}
#endif
