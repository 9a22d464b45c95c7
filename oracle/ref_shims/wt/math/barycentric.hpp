// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: barycentric_t (include/wt/math/barycentric.hpp, which pulls in the mesh
// headers) as sampler.hpp's uniform_triangle returns it: the two free coordinates.
#pragma once
#include <optional>
#include <wt/math/common.hpp>
namespace wt {
struct barycentric_t { vec2_t uv{}; constexpr barycentric_t() = default; constexpr explicit barycentric_t(vec2_t v) : uv(v) {} };
// Stand-in for include/wt/math/barycentric.hpp:109-129, needed only so that src/math/gaussian2d.cpp links: it serves the Dirac branch of
// gaussian2d_t::integrate_triangle (sigma == 0), which no wavefront on the path reaches and which the pinning test does not exercise.
inline std::optional<barycentric_t> barycentric_if_point_inside(const vec2_t& a, const vec2_t& b, const vec2_t& c, const vec2_t& p) noexcept {
    const f_t A = (a.x * (b.y - c.y) - a.y * (b.x - c.x)) + (b.x * c.y - c.x * b.y);
    const f_t s = m::sign(A);
    const f_t u = s * (m::eft::diff_prod(b.x, c.y, c.x, b.y) + (b.y - c.y) * p.x + (c.x - b.x) * p.y);
    const f_t v = s * (m::eft::diff_prod(c.x, a.y, a.x, c.y) + (c.y - a.y) * p.x + (a.x - c.x) * p.y);
    if (u >= 0 && v >= 0 && u + v <= m::abs(A)) return barycentric_t(1 / m::abs(A) * vec2_t{ u, v });
    return std::nullopt;
}
}
