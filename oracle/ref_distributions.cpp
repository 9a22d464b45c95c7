// ORACLE -- TEST INFRASTRUCTURE ONLY.  oracle/_ref/libref_distributions.so: the REFERENCE'S OWN tabulated 1-D distributions, compiled unmodified:
// binned_piecewise_linear_distribution_t (include/wt/math/distribution/binned_piecewise_linear_distribution.hpp: the constructor's f32 running
// trapezoid sum, normalisation and binned inverse, value / pdf, icdf / sample) -- the emitter x sensor product spectrum every sample draws its
// wavenumber from -- and discrete_distribution_t<f_t> (discrete_distribution.hpp:27-136: accumulate, normalise, icdf), which picks the emitter
// (SURVEY.md 8 row a19), and gaussian1d_t::integrate (gaussian1d.hpp:100-106), the film's reconstruction-filter weights (a20).  The class members are private: the access specifier is lifted for this translation unit, after every dependency
// has been included.  Pins (1) the tables wave_tracer_b200/scene.py bakes and (2) ot_scene.h's binned_icdf / binned_value / discrete_icdf:
// tests/test_oracle_kats.py::test_spectrum_distributions_equal_the_reference_code.
#include <memory>
#include <vector>
#include <optional>
#include <algorithm>
#include <numeric>
#include <stdexcept>
#include <cstdint>
#include <cstring>
#include <wt/math/common.hpp>
#include <wt/sampler/sampler.hpp>
#include <wt/sampler/measure.hpp>
#include <wt/math/range.hpp>
#define private public
#include <wt/math/distribution/binned_piecewise_linear_distribution.hpp>
#include <wt/math/distribution/discrete_distribution.hpp>
#undef private
#include <wt/math/distribution/gaussian1d.hpp>

extern "C" {
// ys: n values on a uniform grid over [xmin, xmax].  dcdf: n; binned: 4n; scalars: dx, recp_dx, sum, norm
void ref_binned_build(unsigned n, const float* ys, float xmin, float xmax, float* dcdf, unsigned* binned, float* scalars) {
    const wt::binned_piecewise_linear_distribution_t d(std::vector<wt::f_t>(ys, ys + n), wt::range_t<>{ xmin, xmax });
    std::memcpy(dcdf, d.dcdf.data(), sizeof(float) * n);
    for (std::size_t i = 0; i < d.binned_icdf.size(); ++i) binned[i] = d.binned_icdf[i];
    scalars[0] = d.dx; scalars[1] = d.recp_dx; scalars[2] = d.sum; scalars[3] = d.norm;
}
// v: m numbers in [0,1] -> icdf: m x 2 (x, y);  x: m abscissae -> value, pdf: m each
void ref_binned_eval(unsigned n, const float* ys, float xmin, float xmax, unsigned m, const float* v, float* icdf, const float* x, float* value, float* pdf) {
    const wt::binned_piecewise_linear_distribution_t d(std::vector<wt::f_t>(ys, ys + n), wt::range_t<>{ xmin, xmax });
    for (unsigned i = 0; i < m; ++i) {
        const auto r = d.icdf(v[i]); icdf[2 * i] = r.x; icdf[2 * i + 1] = r.y;
        value[i] = d.value(x[i]); pdf[i] = d.pdf(x[i]);
    }
}
// densities: n -> dcdf: n + 1;  v: m numbers -> idx: m
void ref_discrete(unsigned n, const float* densities, float* dcdf, unsigned m, const float* v, int* idx) {
    const wt::discrete_distribution_t<wt::f_t> d(std::vector<wt::f_t>(densities, densities + n));
    std::memcpy(dcdf, d.dcdf.data(), sizeof(float) * (n + 1));
    for (unsigned i = 0; i < m; ++i) idx[i] = (int)d.icdf(v[i]);
}
// gaussian1d_t::integrate over [mn, mx] (gaussian1d.hpp:100-106): the film's reconstruction-filter mass over a pixel (film.hpp:308-340)
void ref_gaussian1d_integrate(float sigma, unsigned n, const float* mn, const float* mx, float* out) {
    const wt::gaussian1d_t g(sigma);
    for (unsigned i = 0; i < n; ++i) out[i] = g.integrate(wt::range_t<wt::f_t>{ mn[i], mx[i] });
}
}
