// ORACLE -- TEST INFRASTRUCTURE ONLY.  oracle/_ref/libref_gaussian2d.so: the REFERENCE'S OWN closed-form integral of an anisotropic 2-D Gaussian
// over a triangle (gaussian2d_t::integrate_triangle, src/math/gaussian2d.cpp:96-192, with its I_gauss* helpers :24-94 and the class of
// include/wt/math/distribution/gaussian2d.hpp), both compiled unmodified from where they lie.  The BDPT connection weights every triangle of a
// Fraunhofer aperture with it.  What the shims under ref_shims/ restate rather than include: the vector types (glm), the
// three 2-D predicates of math/util.hpp and intersect_edge_circle of math/intersect/misc.hpp (whose real headers are written over mp-units
// quantities and the wide-vector types), and barycentric_if_point_inside (Dirac branch only; not exercised).
// Pins ot_bdpt.h's gaussian2d_t::integrate_triangle: tests/test_oracle_kats.py::test_gaussian_triangle_integral_equals_the_reference_code.
#include <wt/math/common.hpp>
#include <wt/math/distribution/gaussian2d.hpp>

extern "C" {
// sigma: the two standard deviations (frame x = (1,0), mean 0: every wavefront on the path is built that way, beam_generic.hpp:130-139);
// tri: n x 6 floats (a.x a.y b.x b.y c.x c.y); out: n integrals
void ref_gaussian_integrate_triangles(float sx, float sy, unsigned n, const float* tri, float* out) {
    const wt::gaussian2d_t g(wt::vec2_t{ sx, sy });
    for (unsigned i = 0; i < n; ++i) {
        const float* t = tri + 6 * i;
        out[i] = g.integrate_triangle(wt::vec2_t{ t[0], t[1] }, wt::vec2_t{ t[2], t[3] }, wt::vec2_t{ t[4], t[5] });
    }
}
// pdf and the canonical-space map (gaussian2d.hpp:96-101, :190-197) that amplitude_magnitude() of the wavefront evaluates; pts: n x 2; out: n x 3
void ref_gaussian_pdf(float sx, float sy, unsigned n, const float* pts, float* out) {
    const wt::gaussian2d_t g(wt::vec2_t{ sx, sy });
    for (unsigned i = 0; i < n; ++i) {
        const wt::vec2_t p{ pts[2 * i], pts[2 * i + 1] };
        const auto c = g.to_canonical(p);
        out[3 * i] = g.pdf(p); out[3 * i + 1] = c.x; out[3 * i + 2] = c.y;
    }
}
}
