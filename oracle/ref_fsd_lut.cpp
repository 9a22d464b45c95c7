// ORACLE -- TEST INFRASTRUCTURE ONLY.  oracle/_ref/libref_fsd_lut.so: the REFERENCE'S OWN importance-sampling table look-up for the Fraunhofer
// lobes (fsd_lut_t::sample and its linear / bilinear lerp, include/wt/interaction/fsd/fraunhofer/fsd_lut.hpp:35-69), compiled unmodified.  The
// function is a private static of a class whose constructor loads Git-LFS files, so the access specifier is lifted for this translation unit
// and the tables are supplied by the caller at the reference's fixed sizes (2048 / 3072 x 3072).
// Pins ot_bdpt.h's lut_t: tests/test_oracle_kats.py::test_fraunhofer_lut_sampling_equals_the_reference_code.
#include <memory>
#include <array>
#include <cstring>
// everything fsd_lut.hpp includes is included BEFORE the access specifier is lifted, so that only the class under test sees it
#include <wt/math/common.hpp>
#include <wt/wt_context.hpp>
#include <wt/util/array.hpp>
#define private public
#include <wt/interaction/fsd/fraunhofer/fsd_lut.hpp>
#undef private

using lut_t = wt::fraunhofer::fsd_sampler::fsd_lut_t;

#include <wt/math/erf_lut.hpp>

extern "C" {
// the 1024-entry erf table of include/wt/math/erf_lut.hpp:20-55 (used by gaussian2d_t::integrate_triangle and the film's reconstruction filter)
float ref_erf_lut(float x) { return wt::m::erf_lut(x); }
unsigned ref_fsd_lut_n(void) { return (unsigned)lut_t::Nsamples; }
unsigned ref_fsd_lut_m(void) { return (unsigned)lut_t::Msamples; }
// theta: 2048 floats; icdf: 3072 x 3072 floats (row = theta bin); rand: n x 3; out: n x 2
void ref_fsd_lut_sample(const float* theta, const float* icdf, unsigned n, const float* rand, float* out) {
    auto th = std::make_unique<wt::array_t<wt::f_t, lut_t::Nsamples>>();
    auto cd = std::make_unique<wt::array_t<wt::f_t, lut_t::Msamples, lut_t::Msamples>>();
    std::memcpy(th->data(), theta, sizeof(float) * lut_t::Nsamples);
    std::memcpy(cd->data(), icdf, sizeof(float) * lut_t::Msamples * lut_t::Msamples);
    for (unsigned i = 0; i < n; ++i) {
        const auto z = lut_t::sample(wt::vec3_t{ rand[3 * i], rand[3 * i + 1], rand[3 * i + 2] }, *th, *cd);
        out[2 * i] = z.x; out[2 * i + 1] = z.y;
    }
}
}
