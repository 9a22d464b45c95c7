// ORACLE -- TEST INFRASTRUCTURE ONLY.
// The reference's own src/math/elliptic_cone.cpp (cone_through_ellipse, cone_through_ellipsoid: the re-fit of a beam's envelope through its footprint
// at every interaction), compiled from where it lies into oracle/_ref/libref_cone.so next to oracle/ref_cone.cpp, which exports the comparisons.
// Its #include <wt/math/intersect/cone.hpp> resolves to ref_shims' forwarder to the build-time cut of that header (oracle/_ref/cone_scalar_part.hpp).
#define WT_SHIM_DISTINCT_PQ
#define WT_SHIM_WIDE_LANES
#include <wt/util/assert.hpp>
#include "/root/reference/include/wt/math/util.hpp"
#include "/root/reference/src/math/elliptic_cone.cpp"
