// host/src/geometry.h -- great-circle helpers of the host facade (glibc math).
//
// Same formulas, constants and evaluation order as the reference's header templates
// (include/osm_converter/compute_haversine.hpp:41-58, compute_bearing.hpp:45-61): the
// planner constructor's selection probabilities must come out bit-identical, so this
// library is compiled without -march / FMA contraction like the reference (build.sh:2).
#ifndef IMOMD_HOST_GEOMETRY_H
#define IMOMD_HOST_GEOMETRY_H

#include <cmath>

namespace imomd_host {

inline double great_circle_m(double lat_a, double lon_a, double lat_b, double lon_b)
{
    static constexpr double kRadPerDeg = M_PI / 180;
    static const int kEarthRadiusM = 6371e3;   // integer metres, as upstream
    const double phi_a = lat_a * kRadPerDeg;
    const double phi_b = lat_b * kRadPerDeg;
    const double dphi = (lat_b - lat_a) * kRadPerDeg;
    const double dlam = (lon_b - lon_a) * kRadPerDeg;
    const double hav = std::pow(std::sin(dphi / 2), 2) +
                       std::cos(phi_a) * std::cos(phi_b) * std::pow(std::sin(dlam / 2), 2);
    const double angle = 2 * std::atan2(std::sqrt(hav), std::sqrt(1 - hav));
    return kEarthRadiusM * angle;
}

// initial bearing a -> b in [0, 2 pi)
inline double initial_bearing(double lat_a, double lon_a, double lat_b, double lon_b)
{
    static constexpr double kRadPerDeg = M_PI / 180;
    const double phi_a = lat_a * kRadPerDeg;
    const double phi_b = lat_b * kRadPerDeg;
    const double dlam = (lon_b - lon_a) * kRadPerDeg;
    const double east = std::sin(dlam) * std::cos(phi_b);
    const double north = std::cos(phi_a) * std::sin(phi_b) -
                         std::sin(phi_a) * std::cos(phi_b) * std::cos(dlam);
    double theta = std::atan2(east, north);
    theta += theta < 0 ? 2 * M_PI : 0;
    return theta;
}

}  // namespace imomd_host
#endif
