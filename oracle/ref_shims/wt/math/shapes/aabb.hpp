// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: src/math/gaussian2d.cpp includes the AABB header and uses nothing from it;
// math/intersect/misc.hpp (WT_SHIM_DISTINCT_PQ build, oracle/ref_misc.cpp) has an AABB-triangle test next to the edge tests under pin -- it needs the
// type to parse, and is not part of the pin.
#pragma once
#include <wt/math/common.hpp>
#ifdef WT_SHIM_DISTINCT_PQ
namespace wt {
struct aabb_t {
    pqvec3_t min, max;
    pqvec3_t centre() const noexcept { return f_t(.5) * (min + max); }
    pqvec3_t extent() const noexcept { return f_t(.5) * (max - min); }
};
}
#endif
