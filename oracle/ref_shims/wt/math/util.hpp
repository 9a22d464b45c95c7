// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: the three 2-D predicates of include/wt/math/util.hpp (which pulls in the eft and
// wide-vector headers) that src/math/gaussian2d.cpp calls, restated from util.hpp:27-30 and :64-78 (diff_prod = the compensated product of
// math/eft/eft.hpp as ot_math.h restates it).
#pragma once
#ifdef WT_SHIM_DISTINCT_PQ
// (the builds in which vectors of lengths are a type of their own compile the reference's own header)
#include "/root/reference/include/wt/math/util.hpp"
#else
#include <wt/math/common.hpp>
namespace wt::util {
[[nodiscard]] inline bool is_point_in_circle(const vec2_t& p, const f_t r, const vec2_t& o = { 0, 0 }) noexcept { return m::length2(p - o) <= m::sqr(r); }
[[nodiscard]] inline bool is_point_in_triangle(const vec2_t& p, const vec2_t& a, const vec2_t& b, const vec2_t& c) noexcept {
    constexpr auto sgn = [](auto p1, auto p2, auto p3) { return m::eft::diff_prod(p1.x - p3.x, p2.y - p3.y, p2.x - p3.x, p1.y - p3.y); };
    const auto s1 = sgn(p, a, b), s2 = sgn(p, b, c), s3 = sgn(p, c, a);
    const auto neg = s1 < 0 || s2 < 0 || s3 < 0, pos = s1 > 0 || s2 > 0 || s3 > 0;
    return !(neg && pos);
}
}
#endif
