// ORACLE -- TEST INFRASTRUCTURE ONLY.  Part of oracle/_ref/libref_gaussian2d.so: the REFERENCE'S OWN clip of a triangle against a depth slab
// (clip_triangle_z and clip_triangle_ret_t::triangle, include/wt/math/intersect/clip.hpp:20-82), compiled unmodified.  The BDPT connection clips
// every aperture triangle to the beam's z-range with it before integrating the wavefront over the pieces.  The shims
// restate the interval type and intersect_edge_plane (misc.hpp:163-180), whose real headers are written over mp-units quantities.
// Pins ot_bdpt.h's clip_triangle_z: tests/test_oracle_kats.py::test_triangle_clip_equals_the_reference_code.
#include <cassert>
#include <optional>
#include <wt/math/common.hpp>
#include <wt/math/intersect/clip.hpp>

extern "C" {
// tri: n x 9 floats; zr: n x 2; ntris: n; polygon: n x 15 (the tris + 2 vertices of the clipped polygon, the rest zero); pieces: n x 27 (triangle(0..2), unused zero)
void ref_clip_triangles(unsigned n, const float* tri, const float* zr, int* ntris, float* polygon, float* pieces) {
    for (unsigned i = 0; i < n; ++i) {
        const float* t = tri + 9 * i;
        const auto r = wt::intersect::clip_triangle_z({ t[0], t[1], t[2] }, { t[3], t[4], t[5] }, { t[6], t[7], t[8] }, wt::pqrange_t<>{ zr[2 * i], zr[2 * i + 1] });
        ntris[i] = r.tris;
        for (int k = 0; k < 5; ++k) { const bool used = r.tris > 0 && k < r.tris + 2; polygon[15 * i + 3 * k] = used ? r.vs[k].x : 0; polygon[15 * i + 3 * k + 1] = used ? r.vs[k].y : 0; polygon[15 * i + 3 * k + 2] = used ? r.vs[k].z : 0; }
        for (int j = 0; j < 3; ++j) {
            if (j >= r.tris) { for (int k = 0; k < 9; ++k) pieces[27 * i + 9 * j + k] = 0; continue; }
            const auto p = r.triangle(j);
            for (int k = 0; k < 3; ++k) { pieces[27 * i + 9 * j + 3 * k] = p[k].x; pieces[27 * i + 9 * j + 3 * k + 1] = p[k].y; pieces[27 * i + 9 * j + 3 * k + 2] = p[k].z; }
        }
    }
}
}
