// host/compat/data/location_t.h -- boundary type of the planner API.
//
// Layout- and name-compatible with the reference's `location_t`
// (include/data/location_t.h:38-56 upstream): {size_t id; double latitude, longitude}.
// Only used when the host facade is built WITHOUT the reference's include tree on the
// include path; next to the reference tree the reference's own header is picked up
// (same include guard), so the two never meet in one translation unit.
#ifndef LOCATION_STRUCT_H
#define LOCATION_STRUCT_H

#include <cstddef>

typedef struct location
{
    size_t id;
    double latitude;
    double longitude;

    location() = default;
    location(size_t id_, double lat, double lon) : id(id_), latitude(lat), longitude(lon) {}
    location(double lat, double lon) : id(static_cast<size_t>(-1)), latitude(lat), longitude(lon) {}
    bool operator==(const location& other) const { return id == other.id; }
} location_t;

#endif /* LOCATION_STRUCT_H */
