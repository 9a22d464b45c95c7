// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: intersect_edge_circle of include/wt/math/intersect/misc.hpp:77-134 (the real header
// is written over mp-units quantities), restated on plain floats: only the `points` count is used by src/math/gaussian2d.cpp; and
// intersect_edge_plane (:163-180), which math/intersect/clip.hpp calls.
#pragma once
#ifdef WT_SHIM_DISTINCT_PQ
// (the builds in which vectors of lengths are a type of their own compile the reference's own header)
#include "/root/reference/include/wt/math/intersect/misc.hpp"
#else
#include <optional>
#include <utility>
#include <wt/math/common.hpp>
namespace wt::intersect {
struct intersect_edge_circle_ret_t { int points = 0; f_t t1 = 0, t2 = 0; vec2_t u1{}, u2{}; };
inline intersect_edge_circle_ret_t intersect_edge_circle(const vec2_t& point0, const vec2_t& point1, const f_t r) noexcept {
    const vec2_t recp_scale = f_t(1) / vec2_t{ r, r };
    const auto p0 = point0 * recp_scale, p1 = point1 * recp_scale;
    const auto d = p1 - p0;
    const auto a = m::dot(d, d), b = 2 * m::dot(p0, d), c = m::dot(p0, p0) - 1;
    const auto det2 = b * b - 4 * a * c;
    if (det2 <= 0 || a == 0) return {};
    const auto recp_a = 1 / a, det = m::sqrt(det2);
    auto t1 = f_t(.5) * (-b - m::sign(b) * det) * recp_a;
    auto t2 = t1 == 0 ? -b * recp_a : c * recp_a / t1;
    if (t1 > t2) std::swap(t1, t2);
    const bool u1valid = t1 >= 0 && 1 >= t1, u2valid = t2 >= 0 && 1 >= t2;
    intersect_edge_circle_ret_t ret; ret.t1 = t1; ret.t2 = t2;
    ret.points = (u1valid ? 1 : 0) + (u2valid ? 1 : 0);
    return ret;
}
inline std::optional<pqvec3_t> intersect_edge_plane(const pqvec3_t& p0, const pqvec3_t& p1, const pqvec3_t& pp, const dir3_t& n) noexcept {
    const auto d0 = m::dot(pp - p0, n), d1 = m::dot(pp - p1, n);
    const auto E = p1 - p0;
    const auto E_dot_N = m::dot(E, n);
    if (m::sign(d0) == m::sign(d1) || E_dot_N == 0) return std::nullopt;
    const auto d = d0 / E_dot_N;
    if (d >= 0 && 1 >= d) return p0 + d * E;
    return std::nullopt;
}
}
#endif
