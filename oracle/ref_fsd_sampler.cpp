// ORACLE -- TEST INFRASTRUCTURE ONLY.  oracle/_ref/libref_fsd_sampler.so: the REFERENCE'S OWN Fraunhofer direction sampler -- the translation unit
// /root/reference/src/interaction/fsd/fraunhofer/fsd_sampler.cpp (sampleP0 / sample1 / sampleN / sample_rejection) compiled unmodified together
// with the reference's own sampler.hpp (discrete(), normal2d(), the warps), fsd.hpp and fsd_lut.hpp, over the shims of oracle/ref_shims/.
// This file supplies what the reference takes from elsewhere: the table constructor of fsd_lut_t (the reference's loads Git-LFS files; here the
// caller's tables are copied in) and a sampler that replays a scripted sequence of numbers, so that the ORDER in which the reference consumes
// its random numbers is observable.  Pins ot_bdpt.h's sampleN / sample_rejection and ot_scene.h's warps:
// tests/test_oracle_kats.py::test_fraunhofer_rejection_sampler_equals_the_reference_code, ::test_sampler_warps_equal_the_reference_code.
#include <memory>
#include <array>
#include <cstring>
#include <vector>
#include <wt/math/common.hpp>
#include <wt/wt_context.hpp>
#include <wt/util/array.hpp>
#include <wt/sampler/sampler.hpp>
#include <wt/interaction/fsd/fraunhofer/fsd.hpp>
#define private public
#include <wt/interaction/fsd/fraunhofer/fsd_lut.hpp>
#undef private
#include <wt/interaction/fsd/fraunhofer/fsd_sampler.hpp>

using namespace wt;
using lut_t = fraunhofer::fsd_sampler::fsd_lut_t;

static const float *g_th1, *g_th2, *g_c1, *g_c2;
// fsd_lut.cpp:28-74 reads data/fsd/iCDFa{1,2}{,theta}.fp64; here the tables come from the caller
lut_t::fsd_lut_t(const wt_context_t&) : data(std::make_unique<data_t>()) {
    std::memcpy(data->iCDFtheta1.data(), g_th1, sizeof(float) * Nsamples); std::memcpy(data->iCDFtheta2.data(), g_th2, sizeof(float) * Nsamples);
    std::memcpy(data->iCDF1.data(), g_c1, sizeof(float) * Msamples * Msamples); std::memcpy(data->iCDF2.data(), g_c2, sizeof(float) * Msamples * Msamples);
}

struct scripted_sampler_t final : sampler::sampler_t {
    const float* script; unsigned n, i = 0;
    scripted_sampler_t(const float* s, unsigned n_) : sampler_t("scripted"), script(s), n(n_) {}
    f_t r() noexcept override { const f_t v = i < n ? script[i] : f_t(.5); ++i; return v; }
    vec2_t r2() noexcept override { const f_t a = r(); const f_t b = r(); return { a, b }; }            // uniform.hpp:36-50: components drawn in order
    vec3_t r3() noexcept override { const f_t a = r(); const f_t b = r(); const f_t c = r(); return { a, b, c }; }
    vec4_t r4() noexcept override { const f_t a = r(); const f_t b = r(); const f_t c = r(); const f_t d = r(); return { a, b, c, d }; }
    scene::element::info_t description() const override { return { "", "scripted" }; }
};

extern "C" {
// tables at the reference's sizes (2048 / 3072 x 3072); edges: n x 8 as in ref_fsd.cpp; out: xi.x, xi.y, pdf, weight, numbers consumed
void ref_fsd_sampler_sample(const float* th1, const float* th2, const float* c1, const float* c2, unsigned n_edges, const float* edges, const float* edge_pdfs,
                            float P0v, float P0_pdf, float psi02, float recp_I, const float* script, unsigned n_script, unsigned n_samples, float* out) {
    g_th1 = th1; g_th2 = th2; g_c1 = c1; g_c2 = c2;
    const wt_context_t ctx;
    const fraunhofer::fsd_sampler::fsd_sampler_t smp("fsd", ctx);
    fraunhofer::fsd::fsd_aperture_t ap; ap.P0 = P0v; ap.P0_pdf = P0_pdf; ap.psi02 = psi02; ap.recp_I = recp_I;
    for (unsigned i = 0; i < n_edges; ++i) {
        const float* e = edges + 8 * i;
        fraunhofer::fsd::edge_t ed; ed.e = { e[0], e[1] }; ed.v = { e[2], e[3] }; ed.a_b = { e[4], e[5] }; ed.iab_2 = { e[6], e[7] };
        ap.edges.push_back(ed); ap.edge_pdfs.push_back(edge_pdfs[i]);
    }
    scripted_sampler_t rng(script, n_script);
    for (unsigned s = 0; s < n_samples; ++s) {
        const auto r = smp.sample(rng, ap);
        out[5 * s] = r.xi.x; out[5 * s + 1] = r.xi.y; out[5 * s + 2] = r.pdf; out[5 * s + 3] = r.weight; out[5 * s + 4] = (float)rng.i;
    }
}
// the warps of sampler.hpp on explicit (u1, u2): out = cosine_hemisphere xyz, concentric_disk xy, uniform_sphere xyz, uniform_cone(sa) xyz, normal2d xy, uniform_hemisphere xyz
void ref_sampler_warps(float u1, float u2, float solid_angle, float out[16]) {
    using S = sampler::sampler_t;
    const vec2_t u{ u1, u2 };
    const auto a = S::cosine_hemisphere(u); const auto b = S::concentric_disk(u); const auto c = S::uniform_sphere(u); const auto d = S::uniform_cone(solid_angle, u);
    const auto e = S::normal2d(u); const auto f = S::uniform_hemisphere(u);
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = b.x; out[4] = b.y; out[5] = c.x; out[6] = c.y; out[7] = c.z; out[8] = d.x; out[9] = d.y; out[10] = d.z;
    out[11] = e.x; out[12] = e.y; out[13] = f.x; out[14] = f.y; out[15] = f.z;
}
}
