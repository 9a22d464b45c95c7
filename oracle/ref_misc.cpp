// ORACLE -- TEST INFRASTRUCTURE ONLY.
// The reference's own include/wt/math/util.hpp (is_point_in_triangle in 2-D and 3-D, is_point_in_circle / ellipse) and
// include/wt/math/intersect/misc.hpp (edge-ellipsoid, edge / line - ellipse, edge-plane, edge-edge: the primitive tests under
// the UTD edge clipping, the Gaussian-triangle integral, clip_triangle_z and the cone tests), compiled unmodified from where it lies with the shim
// mode in which vectors of lengths are a type of their own (WT_SHIM_DISTINCT_PQ) -> oracle/_ref/libref_misc.so.  tests/test_oracle_kats.py compares
// it bit for bit with ot_math.h.  (The other pins keep using ref_shims/wt/math/intersect/misc.hpp, a restatement on plain floats: this TU names the
// real header by its path.)
#define WT_SHIM_DISTINCT_PQ
#include <wt/util/assert.hpp>
#include "/root/reference/include/wt/math/intersect/misc.hpp"
#include "/root/reference/include/wt/math/util.hpp"
using namespace wt;
extern "C" {
// per item in: p0[3] p1[3] centre[3] x[3] y[3] axes[3]; out: t1 t2
void ref_edge_ellipsoid(unsigned n, const float* in, float* out) {
    for (unsigned i = 0; i < n; ++i) {
        const float* a = in + 18 * i;
        const auto r = intersect::intersect_edge_ellipsoid(pqvec3_t{ a[0], a[1], a[2] }, pqvec3_t{ a[3], a[4], a[5] }, pqvec3_t{ a[6], a[7], a[8] }, dir3_t{ a[9], a[10], a[11] }, dir3_t{ a[12], a[13], a[14] }, pqvec3_t{ a[15], a[16], a[17] });
        out[2 * i] = r.t1; out[2 * i + 1] = r.t2;
    }
}
// per item in: p0[2] p1[2] rx ry; out: points t1 t2 u1[2] u2[2] for the edge test, then the same 7 for the line test
void ref_edge_ellipse(unsigned n, const float* in, float* out) {
    for (unsigned i = 0; i < n; ++i) {
        const float* a = in + 6 * i; float* o = out + 14 * i;
        const auto e = intersect::intersect_edge_ellipse(pqvec2_t{ a[0], a[1] }, pqvec2_t{ a[2], a[3] }, a[4], a[5]);
        const auto l = intersect::intersect_line_ellipse(pqvec2_t{ a[0], a[1] }, pqvec2_t{ a[2], a[3] }, a[4], a[5]);
        o[0] = (float)e.points; o[1] = e.t1; o[2] = e.t2; o[3] = e.u1.x; o[4] = e.u1.y; o[5] = e.u2.x; o[6] = e.u2.y;
        o[7] = (float)l.points; o[8] = l.t1; o[9] = l.t2; o[10] = l.u1.x; o[11] = l.u1.y; o[12] = l.u2.x; o[13] = l.u2.y;
    }
}
// per item in: p0[3] p1[3] pp[3] n[3]; out: found x y z
void ref_edge_plane(unsigned n, const float* in, float* out) {
    for (unsigned i = 0; i < n; ++i) {
        const float* a = in + 12 * i; float* o = out + 4 * i;
        const auto r = intersect::intersect_edge_plane(pqvec3_t{ a[0], a[1], a[2] }, pqvec3_t{ a[3], a[4], a[5] }, pqvec3_t{ a[6], a[7], a[8] }, dir3_t{ a[9], a[10], a[11] });
        o[0] = r ? 1.f : 0.f; o[1] = r ? r->x : 0.f; o[2] = r ? r->y : 0.f; o[3] = r ? r->z : 0.f;
    }
}
// per item in: p[3] a[3] b[3] c[3]; out: 1 / 0  (util.hpp:88-107, the test the cone-plane stage of intersect_cone_tri ends with)
void ref_point_in_triangle3(unsigned n, const float* in, float* out) {
    for (unsigned i = 0; i < n; ++i) { const float* a = in + 12 * i; out[i] = util::is_point_in_triangle(vec3_t{ a[0], a[1], a[2] }, vec3_t{ a[3], a[4], a[5] }, vec3_t{ a[6], a[7], a[8] }, vec3_t{ a[9], a[10], a[11] }) ? 1.f : 0.f; }
}
// per item in: p[2] a[2] b[2] c[2]; out: 1 / 0  (util.hpp:69-82)
void ref_point_in_triangle2(unsigned n, const float* in, float* out) {
    for (unsigned i = 0; i < n; ++i) { const float* a = in + 8 * i; out[i] = util::is_point_in_triangle(vec2_t{ a[0], a[1] }, vec2_t{ a[2], a[3] }, vec2_t{ a[4], a[5] }, vec2_t{ a[6], a[7] }) ? 1.f : 0.f; }
}
}
