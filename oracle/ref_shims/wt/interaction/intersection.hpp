// ORACLE -- TEST INFRASTRUCTURE ONLY.  Stand-in for include/wt/interaction/intersection.hpp: the declarations (and only those) that
// include/wt/math/shapes/elliptic_cone.hpp names, so that the reference's own cone header compiles for oracle/ref_cone.cpp.  No member that
// the pinned functions call is defined here.
#pragma once
#include <wt/math/common.hpp>
namespace wt {
struct intersection_footprint_t { pqvec2_t va, vb; const pqvec2_t& a() const { return va; } const pqvec2_t& b() const { return vb; } };
struct intersection_geo_t { dir3_t n; pqvec3_t to_world(const pqvec2_t&) const; };
struct intersection_surface_t { intersection_footprint_t footprint; intersection_geo_t geo; };
#ifdef WT_SHIM_WIDE_LANES
}
#include <wt/ads/common.hpp>
#include <wt/math/shapes/ray.hpp>
namespace wt {
// intersection.hpp:167-184, the members src/interaction/intersection.cpp:187-211 defines (oracle/ref_traverse.cpp)
struct intersection_edge_t { const ads::edge_t* edge; pqvec3_t wp; pqvec3_t offseted_ray_origin(const ray_t& ray) const noexcept; };
#endif
}
