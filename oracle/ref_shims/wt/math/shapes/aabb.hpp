// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: src/math/gaussian2d.cpp includes the AABB header and uses nothing from it;
// math/intersect/misc.hpp (WT_SHIM_DISTINCT_PQ build, oracle/ref_misc.cpp) has an AABB-triangle test next to the edge tests under pin -- it needs the
// type to parse, and is not part of the pin.
#pragma once
#include <algorithm>
#include <wt/math/common.hpp>
#ifdef WT_SHIM_DISTINCT_PQ
namespace wt {
struct aabb_t {
    pqvec3_t min, max;
    pqvec3_t centre() const noexcept { return f_t(.5) * (min + max); }
    pqvec3_t extent() const noexcept { return f_t(.5) * (max - min); }
    // aabb.hpp:240-242: the union of the points' boxes, componentwise min / max
    static aabb_t from_points(const pqvec3_t& a, const pqvec3_t& b, const pqvec3_t& c) noexcept {
        return { { std::min(std::min(a.x, b.x), c.x), std::min(std::min(a.y, b.y), c.y), std::min(std::min(a.z, b.z), c.z) }, { std::max(std::max(a.x, b.x), c.x), std::max(std::max(a.y, b.y), c.y), std::max(std::max(a.z, b.z), c.z) } };
    }
};
}
#endif
