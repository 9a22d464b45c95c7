// ORACLE -- TEST INFRASTRUCTURE ONLY.  oracle/_ref/libref_fsd.so: the REFERENCE'S OWN Fraunhofer free-space-diffraction formulas, compiled
// unmodified from /root/reference/include/wt/interaction/fsd/fraunhofer/fsd.hpp (alpha1, alpha2, chi_e, chi_0, Psi, Psi2, sampling_density, ASF,
// P0, Pj) over the shim of oracle/ref_shims/.  Pins ot_bdpt.h's restatement: tests/test_oracle_kats.py::test_fraunhofer_formulas_equal_the_reference_code.
#include <wt/interaction/fsd/fraunhofer/fsd.hpp>

using namespace wt::fraunhofer::fsd;

extern "C" {
// edges: n x (e.x, e.y, v.x, v.y, a_b.re, a_b.im, iab_2.re, iab_2.im); out: ASF_unclamped, ASF, sampling_density, chi_e, chi_0, Pj(edge 0), P0(aperture), alpha1(xi), alpha2(xi)
void ref_fsd_eval(unsigned n, const float* edges, float P0v, float psi02, float xix, float xiy, float out[9]) {
    fsd_aperture_t ap; ap.P0 = P0v; ap.P0_pdf = 0; ap.psi02 = psi02; ap.recp_I = 1;
    for (unsigned i = 0; i < n; ++i) {
        const float* e = edges + 8 * i;
        edge_t ed; ed.e = { e[0], e[1] }; ed.v = { e[2], e[3] }; ed.a_b = { e[4], e[5] }; ed.iab_2 = { e[6], e[7] };
        ap.edges.push_back(ed);
    }
    const wt::vec2_t xi{ xix, xiy };
    out[0] = ASF_unclamped(ap, xi); out[1] = ASF(ap, xi); out[2] = sampling_density(ap, xi); out[3] = chi_e(xi); out[4] = chi_0(xi);
    out[5] = n ? Pj(ap.edges[0]) : 0.f; out[6] = P0(ap); out[7] = alpha1(xi); out[8] = alpha2(xi);
}
}
