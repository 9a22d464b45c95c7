// ORACLE -- TEST INFRASTRUCTURE ONLY.  oracle/_ref/libref_utd.so: the REFERENCE'S OWN uniform theory of diffraction for a wedge
// (include/wt/interaction/fsd/utd.hpp: UTDa, the transition function UTDF, wedge_edge_t::diffraction_point both ways and wedge_edge_t::UTD with
// its soft / hard coefficients and the four transverse frames; include/wt/interaction/fsd/common.hpp for the types), compiled unmodified.  It is
// what plt_path's free-space diffraction evaluates per edge and per sample.  The shims supply plain-float stand-ins for the
// mp-units quantities (with the 1/mm x m -> 1000 unit ratio of u::to_num spelled out), the vecmath cross / dot / normalize and libcerf's cerfc,
// which the TEST installs (ref_set_cerfc) from an implementation independent of the oracle's own series.
// Pins ot_integrator.h's UTDa / UTDF / wedge_edge_t: tests/test_oracle_kats.py::test_utd_equals_the_reference_code.
#include <optional>
#include <vector>
#include <wt/util/assert.hpp>
#include <wt/math/common.hpp>
#include <wt/interaction/fsd/utd.hpp>

extern "C" {
ref_cerfc_fn ref_cerfc_hook = nullptr;
void ref_set_cerfc(ref_cerfc_fn f) { ref_cerfc_hook = f; }

void ref_utdf(unsigned n, const float* x, float* out) {
    for (unsigned i = 0; i < n; ++i) { const auto f = wt::utd::UTDF(x[i]); out[2 * i] = f.real(); out[2 * i + 1] = f.imag(); }
}
// wedge: n x 14 floats (v[3] l nff[3] tff[3] nbf[3] alpha); q: n x 8 (k [1/mm], wi[3], wo[3], ro [m]); out: n x 16 (Ds re im, Dh re im, si, hi, so, ho)
void ref_utd(unsigned n, const float* wedge, const float* q, float* out) {
    for (unsigned i = 0; i < n; ++i) {
        const float* w = wedge + 14 * i; const float* a = q + 8 * i; float* o = out + 16 * i;
        wt::utd::wedge_edge_t e{};
        e.v = { w[0], w[1], w[2] }; e.l = w[3];
        e.nff = wt::dir3_t{ w[4], w[5], w[6] }; e.tff = wt::dir3_t{ w[7], w[8], w[9] }; e.nbf = wt::dir3_t{ w[10], w[11], w[12] }; e.alpha = w[13];
        const auto r = e.UTD(wt::wavenumber_t{ a[0] }, wt::dir3_t{ a[1], a[2], a[3] }, wt::dir3_t{ a[4], a[5], a[6] }, a[7]);
        o[0] = r.Ds.real(); o[1] = r.Ds.imag(); o[2] = r.Dh.real(); o[3] = r.Dh.imag();
        const wt::dir3_t* f[4] = { &r.si, &r.hi, &r.so, &r.ho };
        for (int k = 0; k < 4; ++k) { o[4 + 3 * k] = f[k]->x; o[5 + 3 * k] = f[k]->y; o[6 + 3 * k] = f[k]->z; }
    }
}
// Fermat points: src, dst (points) -> found[2i], p; src, wo (direction) -> found[2i+1], p.  pts: n x 9 (src, dst, wo); out: n x 6
void ref_utd_diffraction_points(unsigned n, const float* wedge, const float* pts, int* found, float* out) {
    for (unsigned i = 0; i < n; ++i) {
        const float* w = wedge + 14 * i; const float* a = pts + 9 * i; float* o = out + 6 * i;
        wt::utd::wedge_edge_t e{};
        e.v = { w[0], w[1], w[2] }; e.l = w[3];
        e.nff = wt::dir3_t{ w[4], w[5], w[6] }; e.tff = wt::dir3_t{ w[7], w[8], w[9] }; e.nbf = wt::dir3_t{ w[10], w[11], w[12] }; e.alpha = w[13];
        const auto p = e.diffraction_point(wt::pqvec3_t{ a[0], a[1], a[2] }, wt::pqvec3_t{ a[3], a[4], a[5] });
        const auto d = e.diffraction_point(wt::pqvec3_t{ a[0], a[1], a[2] }, wt::dir3_t{ a[6], a[7], a[8] });
        found[2 * i] = p ? 1 : 0; found[2 * i + 1] = d ? 1 : 0;
        for (int k = 0; k < 6; ++k) o[k] = 0;
        if (p) { o[0] = p->x; o[1] = p->y; o[2] = p->z; }
        if (d) { o[3] = d->x; o[4] = d->y; o[5] = d->z; }
    }
}
}
