// ORACLE -- TEST INFRASTRUCTURE ONLY.
// The reference's own include/wt/math/frame.hpp (frame_t: build_orthogonal_frame, build_shading_frame, to_local / to_world for plain and length
// vectors, handness), compiled unmodified from where it lies, with a shim in which vectors of lengths are a type of their own (frame.hpp
// overloads on them) -> oracle/_ref/libref_frame.so.  tests/test_oracle_kats.py compares it bit for bit with ot_math.h's frame_t.
#define WT_SHIM_DISTINCT_PQ
#include <wt/util/assert.hpp>
#include <wt/math/frame.hpp>
#include <wt/math/rotation.hpp>
using namespace wt;
static void put(const frame_t& f, float* o) { o[0] = f.t.x; o[1] = f.t.y; o[2] = f.t.z; o[3] = f.b.x; o[4] = f.b.y; o[5] = f.b.z; o[6] = f.n.x; o[7] = f.n.y; o[8] = f.n.z; }
extern "C" {
void ref_frame_orthogonal(unsigned n, const float* nrm, float* out) { for (unsigned i = 0; i < n; ++i) put(frame_t::build_orthogonal_frame(dir3_t{ nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2] }), out + 9 * i); }
void ref_frame_shading(unsigned n, const float* nrm, const float* dpdu, float* out) {
    for (unsigned i = 0; i < n; ++i) put(frame_t::build_shading_frame(dir3_t{ nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2] }, pqvec3_t{ dpdu[3 * i], dpdu[3 * i + 1], dpdu[3 * i + 2] }), out + 9 * i);
}
// util::rotation_matrix(dir2_t from, dir2_t to) (math/rotation.hpp:66-77: the rotation Stokes / Mueller frames are re-expressed with): the 2x2 matrix, column-major
void ref_rotation2(unsigned n, const float* from, const float* to, float* out) {
    for (unsigned i = 0; i < n; ++i) { const mat2_t R = util::rotation_matrix(dir2_t{ from[2 * i], from[2 * i + 1] }, dir2_t{ to[2 * i], to[2 * i + 1] }); out[4 * i] = R[0].x; out[4 * i + 1] = R[0].y; out[4 * i + 2] = R[1].x; out[4 * i + 3] = R[1].y; }
}
// per item: to_local(vec3) 3, to_world(vec3) 3, to_local(pqvec3) 3, to_world(pqvec3) 3, to_local(vec2) 2, to_world(vec2) 3, to_local(dir3) 3, handness 1 = 21 floats
void ref_frame_xform(unsigned n, const float* fr, const float* v, float* out) {
    for (unsigned i = 0; i < n; ++i) {
        const float* f = fr + 9 * i; const float* p = v + 3 * i; float* o = out + 21 * i;
        const frame_t F{ dir3_t{ f[0], f[1], f[2] }, dir3_t{ f[3], f[4], f[5] }, dir3_t{ f[6], f[7], f[8] } };
        const vec3_t a = F.to_local(vec3_t{ p[0], p[1], p[2] }); o[0] = a.x; o[1] = a.y; o[2] = a.z;
        const vec3_t b = F.to_world(vec3_t{ p[0], p[1], p[2] }); o[3] = b.x; o[4] = b.y; o[5] = b.z;
        const pqvec3_t c = F.to_local(pqvec3_t{ p[0], p[1], p[2] }); o[6] = c.x; o[7] = c.y; o[8] = c.z;
        const pqvec3_t d = F.to_world(pqvec3_t{ p[0], p[1], p[2] }); o[9] = d.x; o[10] = d.y; o[11] = d.z;
        const vec2_t e = F.to_local(vec2_t{ p[0], p[1] }); o[12] = e.x; o[13] = e.y;
        const vec3_t g = F.to_world(vec2_t{ p[0], p[1] }); o[14] = g.x; o[15] = g.y; o[16] = g.z;
        const dir3_t h = F.to_local(dir3_t{ p[0], p[1], p[2] }); o[17] = h.x; o[18] = h.y; o[19] = h.z;
        o[20] = F.handness();
    }
}
}
