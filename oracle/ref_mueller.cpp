// ORACLE -- TEST INFRASTRUCTURE ONLY.
// The reference's own include/wt/interaction/polarimetric/mueller.hpp and stokes.hpp (Mueller operators: product, action on a Stokes vector, the
// rotation operator with its R*R trick and explicit transposes, the Fresnel operator and its reflection / transmission forms over the reference's own
// fresnel.hpp, change of incident / exitant frame with handness flips, compose(); Stokes vectors: reorient with handness detection, the frame-aware
// operator() forms), compiled unmodified from where they lie -> oracle/_ref/libref_mueller.so.  tests/test_oracle_kats.py compares it bit for bit with
// ot_polar.h.  What stands in beneath it (ref_shims/wt/math/common.hpp, WT_SHIM_MAT4): glm's vec4 / column-major mat4 with glm's product order, and
// the quantity aliases (plain floats).
#define WT_SHIM_DISTINCT_PQ
#define WT_SHIM_MAT4
#include <format>
#include <wt/util/assert.hpp>
#include "/root/reference/include/wt/interaction/polarimetric/mueller.hpp"
using namespace wt;
namespace {
inline mueller_operator_t load_m(const float* a) { mat4_t M; for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) M[c][r] = a[4 * c + r]; return mueller_operator_t{ M }; }
inline void put_m(const mueller_operator_t& M, float* o) { for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) o[4 * c + r] = M.matrix()[c][r]; }
inline frame_t load_f(const float* f) { return frame_t{ dir3_t{ f[0], f[1], f[2] }, dir3_t{ f[3], f[4], f[5] }, dir3_t{ f[6], f[7], f[8] } }; }
using S_t = stokes_parameters_t<f_t>;
inline void put_s(const S_t& s, float* o) { for (int i = 0; i < 4; ++i) o[i] = s.S[i]; }
}
extern "C" {
// per item in: A[16] B[16] (m[col][row]) S[4] F1[9] F2[9] (sharing a normal) F3[9] F4[9] (sharing a normal) t1[2] t2[2] fs[2] fp[2] eta[2] w[3] = 89
// out: A*B, A*S, rotation(t1,t2), fresnel(fs,fp), fresnel_reflection(eta,w), fresnel_transmission(eta,w), A.change_incident_frame(F1,F2),
//      A.change_exitant_frame(F1,F2), compose(A,B,F1,F2), S.reorient(F1,F2), A(S,F1,F2), A(S,F1,F2,F3,F4) = 8*16 + 4*4 = 144
void ref_mueller(unsigned n, const float* in, float* out) {
    for (unsigned i = 0; i < n; ++i) {
        const float* a = in + 89 * i; float* o = out + 144 * i;
        const auto A = load_m(a), B = load_m(a + 16);
        const S_t S{ .S = { a[32], a[33], a[34], a[35] } };
        const frame_t F1 = load_f(a + 36), F2 = load_f(a + 45), F3 = load_f(a + 54), F4 = load_f(a + 63);
        const dir2_t t1{ a[72], a[73] }, t2{ a[74], a[75] };
        const c_t fs{ a[76], a[77] }, fp{ a[78], a[79] }, eta{ a[80], a[81] };
        const dir3_t w{ a[82], a[83], a[84] };
        put_m(A * B, o); put_s(A * S, o + 16);
        put_m(mueller_operator_t::rotation(t1, t2), o + 20);
        put_m(mueller_operator_t::fresnel(fs, fp), o + 36);
        put_m(mueller_operator_t::fresnel_reflection(eta, w), o + 52);
        put_m(mueller_operator_t::fresnel_transmission(eta, w), o + 68);
        put_m(A.change_incident_frame(F1, F2), o + 84);
        put_m(A.change_exitant_frame(F1, F2), o + 100);
        put_m(compose(A, B, F1, F2), o + 116);
        put_s(S.reorient(F1, F2), o + 132);
        put_s(A(S, F1, F2), o + 136);
        put_s(A(S, F1, F2, F3, F4), o + 140);
    }
}
}
