// ORACLE -- TEST INFRASTRUCTURE ONLY.
// The reference's own src/math/gaussian2d.cpp once more, compiled from where it lies in the shim mode of oracle/ref_traverse.cpp (vectors of lengths a type
// of their own) and linked into oracle/_ref/libref_traverse.so: plt_bdpt's find_closest_triangle integrates the beam's Gaussian over clipped triangles with it.
#define WT_SHIM_DISTINCT_PQ
#define WT_SHIM_WIDE_LANES
#include <cstdint>
#include <wt/util/assert.hpp>
#include "/root/reference/include/wt/math/util.hpp"
#include "/root/reference/src/math/gaussian2d.cpp"
