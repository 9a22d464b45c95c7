// ORACLE -- TEST INFRASTRUCTURE ONLY.  oracle/_ref/libref_fresnel.so: the REFERENCE'S OWN Fresnel equations, compiled unmodified from
// /root/reference/include/wt/interaction/fresnel.hpp (reflect / refract / fresnel / fresnel_reflection) over the shim of oracle/ref_shims/.
// Pins ot_polar.h's restatement (and through the GPU parity tests the device's): tests/test_oracle_kats.py::test_fresnel_equals_the_reference_code.
#include <wt/interaction/fresnel.hpp>

extern "C" {
// out: rs, rp, ts, tp (re, im each), Ts, Tp, Z, t.xyz, eta_12 (re, im) = 16 floats -- fresnel() of fresnel.hpp:74-117, n = +z
void ref_fresnel(float eta_re, float eta_im, const float w[3], float out[16]) {
    const auto f = wt::fresnel(wt::c_t{ eta_re, eta_im }, wt::dir3_t{ w[0], w[1], w[2] });
    out[0] = f.rs.real(); out[1] = f.rs.imag(); out[2] = f.rp.real(); out[3] = f.rp.imag();
    out[4] = f.ts.real(); out[5] = f.ts.imag(); out[6] = f.tp.real(); out[7] = f.tp.imag();
    out[8] = f.Ts; out[9] = f.Tp; out[10] = f.Z; out[11] = f.t.x; out[12] = f.t.y; out[13] = f.t.z; out[14] = f.eta_12.real(); out[15] = f.eta_12.imag();
}
// out: rs, rp (re, im each) -- fresnel_reflection() of fresnel.hpp:128-144 (conductors), n = +z
void ref_fresnel_reflection(float eta_re, float eta_im, const float w[3], float out[4]) {
    const auto f = wt::fresnel_reflection(wt::c_t{ eta_re, eta_im }, wt::dir3_t{ w[0], w[1], w[2] });
    out[0] = f.rs.real(); out[1] = f.rs.imag(); out[2] = f.rp.real(); out[3] = f.rp.imag();
}
void ref_reflect(const float w[3], float out[3]) { const auto r = wt::reflect(wt::dir3_t{ w[0], w[1], w[2] }); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
}
