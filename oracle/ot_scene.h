// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// ot_scene.h: everything the walk calls sideways -- spectra, surface records, BSDFs, emitters, sensors, film --
// restated from the reference over the flattened tables of include/wtgpu.h.
#pragma once
#include "ot_beam.h"
#include "ot_ads.h"
#include "ot_rng.h"
#include <vector>
#include <mutex>

namespace ot {

// sampling densities with a discrete/continuous flag: include/wt/sampler/density.hpp:24-90
struct pd_t {
    f_t v = 0; bool is_discrete = true;
    static pd_t discrete(f_t m) { return { m, true }; }
    static pd_t density(f_t d) { return { d, false }; }
    f_t density_or_zero() const { return is_discrete ? 0.f : v; }
    bool is_zero() const { return v == 0; }
};

struct scene_t {
    const wtgpu_scene_desc* d;
    ads_t ads;
    explicit scene_t(const wtgpu_scene_desc* desc) : d(desc), ads(desc) {}

    // ---- spectra (spectrum_t::value / spectrum_real_t::f, baked)
    c_t spectrum_value(int32_t id, f_t k) const {
        if (id < 0) return { 1, 0 };
        const wtgpu_spectrum& s = d->spectra[id];
        if (s.type == WTGPU_SPECTRUM_CONSTANT) return { s.re, s.im };
        const f_t x = (k - s.k0) * s.inv_dk;
        if (!(x >= 0) || s.n == 0) return { 0, 0 };
        const uint32_t i0 = (uint32_t)x;
        if (i0 + 1 >= s.n) { if (x > (f_t)(s.n - 1)) return { 0, 0 }; const float* p = d->spectrum_data + 2 * (s.offset + s.n - 1); return { p[0], p[1] }; }
        const f_t f = x - (f_t)i0;
        const float* p0 = d->spectrum_data + 2 * (s.offset + i0);
        return { mix(p0[0], p0[2], f), mix(p0[1], p0[3], f) };
    }
    f_t spectrum_f(int32_t id, f_t k) const { return spectrum_value(id, k).real(); }

    // ---- surface records
    // intersection_surface_t(shape, geo_n, mesh_tri_idx, bary, centre) (src/interaction/intersection.cpp:33-70)
    surface_t make_surface(uint32_t tuid, v2 bary, v3 centre) const {
        const wtgpu_tri_shading& sh = d->tri_shading[tuid];
        const f_t w2 = 1 - bary.x - bary.y;
        auto lerp3 = [&](const float* a, const float* b, const float* c) {
            return v3{ bary.x * a[0] + bary.y * b[0] + w2 * c[0], bary.x * a[1] + bary.y * b[1] + w2 * c[1], bary.x * a[2] + bary.y * b[2] + w2 * c[2] };
        };
        surface_t s;
        s.wp = centre; s.bary = bary; s.tuid = tuid; s.has_shape = true;
        s.uv = sh.has_uv ? v2{ bary.x * sh.uv0[0] + bary.y * sh.uv1[0] + w2 * sh.uv2[0], bary.x * sh.uv0[1] + bary.y * sh.uv1[1] + w2 * sh.uv2[1] } : v2{ 0, 0 };
        const v3 n = normalize(lerp3(sh.n0, sh.n1, sh.n2));
        const v3 dpdu{ sh.dpdu[0], sh.dpdu[1], sh.dpdu[2] };
        s.geo = frame_t::build_shading_frame(ads.tri_n(tuid), dpdu);
        s.shading = frame_t::build_shading_frame(n, dpdu);       // bsdf_t::shading_frame default (bsdf.hpp:49-54)
        return s;
    }
    // intersection_surface_t(shape, mesh_tri_idx, bary): position from barycentrics (intersection.cpp:60-68)
    surface_t make_surface_at_bary(uint32_t tuid, v2 bary) const {
        const f_t w2 = 1 - bary.x - bary.y;
        const v3 a = ads.tri_a(tuid), b = ads.tri_b(tuid), c = ads.tri_c(tuid);
        const v3 p{ bary.x * a.x + bary.y * b.x + w2 * c.x, bary.x * a.y + bary.y * b.y + w2 * c.y, bary.x * a.z + bary.y * b.z + w2 * c.z };
        return make_surface(tuid, bary, p);
    }
    // dummy surface (intersection.hpp:97-103)
    static surface_t make_dummy_surface(v3 n, v3 p) {
        surface_t s; s.wp = p; s.geo = frame_t::build_orthogonal_frame(n); s.shading = s.geo; s.has_shape = false; return s;
    }

    // compute_intersection_triangle_fp_errors + offseted_ray_origin (intersection.cpp:149-185)
    static v3 tri_fp_errors(v3 a, v3 b, v3 c, v3 ro) {
        const f_t c0 = 3e-6f, c1 = 5e-6f, c2 = 3e-6f;
        const v3 v0 = absv(a), e1 = absv(b - a), e2 = absv(c - a);
        const v3 extents = e1 + e2 + absv(e1 - e2);
        const f_t extent = max_element(extents);
        const v3 obj_err = (c0 + c2) * v0 + v3{ c1 * extent, c1 * extent, c1 * extent };
        const v3 wrld_err = (c1 + c2) * absv(ro);
        return obj_err + wrld_err;
    }
    v3 offseted_ray_origin(const surface_t& s, const ray_t& ray) const {
        if (!s.has_shape) return ray.o;
        const v3 err = tri_fp_errors(ads.tri_a(s.tuid), ads.tri_b(s.tuid), ads.tri_c(s.tuid), ray.o);
        const f_t offset_dist = dot(err, absv(s.ng()));
        const v3 offset = offset_dist * s.ng();
        return ray.o + (dot(ray.d, offset) >= 0 ? offset : -offset);
    }
    // intersection_edge_t::offseted_ray_origin (intersection.cpp:187-211)
    v3 offseted_ray_origin_edge(uint32_t edge_idx, const ray_t& ray) const {
        const wtgpu_edge& e = d->edges[edge_idx];
        const v3 t1{ e.t1[0], e.t1[1], e.t1[2] }, t2{ e.t2[0], e.t2[1], e.t2[2] };
        const bool has2 = e.tri2 != WTGPU_INVALID_IDX;
        v3 dir{ 0, 0, 1 };
        if (!has2) dir = -t1;
        else { const v3 v = t1 + t2; dir = length2(v) > 1e-14f ? -normalize(v) : -t2; }
        const v3 err1 = tri_fp_errors(ads.tri_a(e.tri1), ads.tri_b(e.tri1), ads.tri_c(e.tri1), ray.o);
        f_t dd = dot(err1, absv(t1));
        if (has2) {
            const v3 err2 = tri_fp_errors(ads.tri_a(e.tri2), ads.tri_b(e.tri2), ads.tri_c(e.tri2), ray.o);
            dd = std::max(dd, dot(err2, absv(t2)));
        }
        return ray.o + dd * dir;
    }
};

// ================================================================================================
// BSDFs
// ================================================================================================
struct bsdf_sample_t { v3 wo; pd_t dpd; c_t eta{ 1, 0 }; mueller_t M; };
struct bsdf_query_t { const surface_t* surface; f_t k; bool forward; uint32_t lobes = 0xffffffffu; };

namespace fractal_details {     // include/wt/interaction/surface_profile/fractal.hpp:20-48
static constexpr f_t max_GGX_alpha = .75f;
static constexpr f_t maxT = 70.f * 70.f;            // mm^2
static constexpr f_t meank = two_pi / 550e-6f;      // wavelen_to_wavenum(550 nm) in 1/mm
inline f_t roughness_to_T(f_t alpha) {
    const f_t alpha2 = sqr(clampf(alpha, 0, max_GGX_alpha));
    return std::min(maxT, (1 - alpha2) / (4 * sqr(meank) * alpha2));
}
inline f_t roughness_to_alpha(f_t alpha) { return sqr(alpha / 9.f); }
}

struct fractal_params_t { f_t T, sigma2_norm, alpha; };

struct bsdf_eval_t {
    const scene_t& sc;
    explicit bsdf_eval_t(const scene_t& s) : sc(s) {}

    const wtgpu_bsdf& node(int32_t id) const { return sc.d->bsdfs[id]; }

    // resolve wrappers that only select a child by wavenumber (composite.hpp:64-70)
    int32_t composite_child(const wtgpu_bsdf& b, f_t k) const {
        for (uint32_t i = 0; i < b.n_bins; ++i) {
            const wtgpu_bsdf_bin& bin = sc.d->bsdf_bins[b.bin_first + i];
            if (bin.kmin <= k && k < bin.kmax) return bin.child;     // left-inclusive range
        }
        return -1;
    }

    // ---- fractal profile (fractal.hpp:67-232, src/interaction/surface_profile/fractal.cpp:27-69)
    fractal_params_t fractal_params(const wtgpu_bsdf& b, f_t k) const {
        const f_t gamma = b.gamma;
        auto sigma2_normalized = [&](f_t T) {
            const f_t x = 1 + k * k * T;
            const f_t p = gamma == 3 ? x : lm::pow(x, (gamma - 1) / 2);
            return 1 / (1 - 1.f / p);
        };
        if (b.profile_type == WTGPU_PROFILE_FRACTAL_ROUGHNESS) {
            const f_t roughness = sc.spectrum_f(b.prof_spec[0], k);
            const f_t T = fractal_details::roughness_to_T(roughness);
            return { T, sigma2_normalized(T), fractal_details::roughness_to_alpha(roughness) };
        }
        const f_t T = sc.spectrum_f(b.prof_spec[0], k);
        const f_t sigmah2 = sqr(sc.spectrum_f(b.prof_spec[1], k));
        return { T, sigma2_normalized(T), sigmah2 };
    }
    f_t fractal_psd(const wtgpu_bsdf& b, const fractal_params_t& p, v2 z, f_t k) const {
        const f_t gamma = b.gamma;
        const f_t x = 1 + p.T * dot(z, z);
        const f_t pw = gamma == 3 ? (x * x) : lm::pow(x, (gamma + 1) / 2);
        const f_t f = 1 / pw;
        return p.sigma2_norm * (inv_two_pi * k * k * (gamma - 1) * p.T * f);
    }
    // ---- gaussian profile (include/wt/interaction/surface_profile/gaussian.hpp:28-255)
    static bool is_gaussian(const wtgpu_bsdf& b) { return b.profile_type == WTGPU_PROFILE_GAUSSIAN || b.profile_type == WTGPU_PROFILE_GAUSSIAN_SIGMA; }
    struct gaussian_params_t { f_t sigma2, sigma2_norm, alpha; };
    gaussian_params_t gaussian_params(const wtgpu_bsdf& b, f_t k) const {     // gaussian.hpp:95-119
        f_t sigma2, alpha;
        if (b.profile_type == WTGPU_PROFILE_GAUSSIAN) {
            const f_t roughness = sc.spectrum_f(b.prof_spec[0], k);
            sigma2 = 1 / fractal_details::roughness_to_T(roughness);
            alpha = fractal_details::roughness_to_alpha(roughness);
        } else {
            sigma2 = sqr(sc.spectrum_f(b.prof_spec[0], k));
            alpha = sigma2;
        }
        return { sigma2, 1 / (1 - lm::exp(-(k * k / 2 / sigma2))), alpha };
    }
    f_t gaussian_psd(const gaussian_params_t& p, v2 z, f_t k) const {          // gaussian.hpp:121-130
        const f_t z2 = dot(z, z);
        const f_t e = lm::exp(-(z2 / 2 / p.sigma2));
        return e <= std::numeric_limits<f_t>::epsilon() ? 0.f : p.sigma2_norm * (inv_two_pi / p.sigma2 * k * k * e);
    }
    static f_t boxmueller_max_phi(f_t r, f_t l) {                               // gaussian.hpp:43-49, 70-76
        const f_t eps = std::numeric_limits<f_t>::epsilon();
        return (r < eps || l < eps) ? pi : std::max(1e-2f, lm::acos(clampf((sqr(r) + sqr(l) - 1) / (2 * r * l), -1, 1)));
    }
    // truncated Box-Mueller transform (gaussian.hpp:28-56); returns the point and its pdf
    static std::pair<v2, f_t> sample_boxmueller_truncated(v2 sample, v2 mean, f_t sigma2) {
        const f_t eps = std::numeric_limits<f_t>::epsilon();
        const f_t l = std::sqrt(std::min(1.f, dot(mean, mean)));
        const f_t coso = std::sqrt(std::max(0.f, 1 - dot(mean, mean)));
        const f_t phi_i = (mean.x != 0 || mean.y != 0) ? lm::atan2(mean.y, mean.x) : 0.f;
        const f_t s = lm::exp(-.5f * sqr(1 + l) / sigma2);
        const f_t x = (1 - s) * std::max(eps, sample.x) + s;
        const f_t r = std::sqrt(-2 * sigma2 * lm::log(x));
        const f_t max_phi = boxmueller_max_phi(r, l);
        const f_t phi = phi_i + pi + max_phi * (2 * sample.y - 1);
        const v2 p = r * v2{ lm::cos(phi), lm::sin(phi) };
        const f_t pdf = .5f * x / (max_phi * sigma2) * coso;
        return { p + mean, pdf };
    }
    static f_t boxmueller_truncated_pdf(v2 wo, v2 mean, f_t sigma2) {          // gaussian.hpp:58-79
        const f_t l = std::sqrt(std::min(1.f, dot(mean, mean)));
        const f_t coso = std::sqrt(std::max(0.f, 1 - dot(mean, mean)));
        wo = wo - mean;
        const f_t r2 = dot(wo, wo);
        const f_t x = lm::exp(-.5f * r2 / sigma2);
        const f_t r = std::sqrt(r2);
        const f_t max_phi = boxmueller_max_phi(r, l);
        return .5f * x / (max_phi * sigma2) * coso;
    }
    bool profile_is_delta_only(const wtgpu_bsdf& b, f_t k) const {
        if (b.profile_type == WTGPU_PROFILE_DIRAC) return true;
        return sc.spectrum_f(b.prof_spec[0], k) == 0;      // mean_value(k)==0 (fractal.hpp:161-169)
    }
    f_t profile_alpha(const wtgpu_bsdf& b, v3 wi, v3 wo, f_t k) const {
        if (b.profile_type == WTGPU_PROFILE_DIRAC) return 1;
        const f_t palpha = is_gaussian(b) ? gaussian_params(b, k).alpha : fractal_params(b, k).alpha;     // gaussian.hpp:162-169 == fractal.hpp
        const f_t a = sqr((std::fabs(wi.z) + std::fabs(wo.z)) * k) * palpha;
        return lm::exp(-a);
    }
    f_t profile_psd(const wtgpu_bsdf& b, v3 wi, v3 wo, f_t k) const {
        if (b.profile_type == WTGPU_PROFILE_DIRAC) return 0;
        const v2 z = k * (v2{ wi.x, wi.y } + v2{ wo.x, wo.y });
        if (is_gaussian(b)) return gaussian_psd(gaussian_params(b, k), z, k);       // gaussian.hpp:197-205
        const auto p = fractal_params(b, k);
        return fractal_psd(b, p, z, k);
    }
    f_t profile_pdf(const wtgpu_bsdf& b, v3 wi, v3 wo, f_t k) const {
        if (b.profile_type == WTGPU_PROFILE_DIRAC) return 0;
        if (is_gaussian(b)) {                                                       // gaussian.hpp:240-250
            const f_t s2 = gaussian_params(b, k).sigma2 / (k * k);
            return boxmueller_truncated_pdf(v2{ wo.x, wo.y }, v2{ -wi.x, -wi.y }, s2);
        }
        const auto p = fractal_params(b, k);
        const v2 zeta_k = v2{ wi.x, wi.y } + v2{ wo.x, wo.y };
        const f_t f_k = length(zeta_k);
        const f_t s = std::sqrt(std::max(0.f, 1 - sqr(wi.z)));
        const f_t phi_max = (f_k == 0 || s == 0) ? pi : lm::acos(clampf((sqr(f_k) + sqr(s) - 1) / (2 * f_k * s), -1, 1));
        const v2 zeta = zeta_k * k;
        const f_t psd = fractal_psd(b, p, zeta, k);
        const f_t w = inv_pi * phi_max;
        return w > 1e-2f ? 1 / w * std::fabs(wo.z) * psd : 0.f;
    }
    struct profile_sample_t { v3 wo; f_t pdf, psd, weight; };
    profile_sample_t profile_sample(const wtgpu_bsdf& b, v3 wi, f_t k, sampler_t& sampler) const {
        if (b.profile_type == WTGPU_PROFILE_DIRAC) return { { 0, 0, 1 }, 0, 0, 0 };
        if (is_gaussian(b)) {                                                       // gaussian.hpp:210-235
            const auto p = gaussian_params(b, k);
            const f_t s2 = p.sigma2 / (k * k);
            const v2 mean = v2{ -wi.x, -wi.y };
            const auto smp = sample_boxmueller_truncated(sampler.r2(), mean, s2);
            const v2 wo2 = smp.first;
            const f_t psd = gaussian_psd(p, k * (wo2 - mean), k);
            const f_t z = std::sqrt(std::max(0.f, 1 - dot(wo2, wo2)));
            return { { wo2.x, wo2.y, wi.z >= 0 ? z : -z }, smp.second, psd, psd / smp.second };
        }
        const f_t gamma = b.gamma;
        const auto p = fractal_params(b, k);
        const f_t s = std::sqrt(std::max(0.f, 1 - sqr(wi.z)));
        const f_t phi_i = s > 0 ? lm::atan2(wi.y, wi.x) : 0.f;
        const f_t sqrtT = std::sqrt(p.T);
        const v2 u2 = sampler.r2();
        const f_t k2T = sqr(k) * p.T;
        const f_t M = 1 - lm::pow(1 + k2T * sqr(1 + s), -(gamma - 1) / 2);
        const f_t f = std::sqrt(lm::pow(1 - M * u2.x, -2 / (gamma - 1)) - 1) / sqrtT;
        const f_t f_k = f / k;
        const f_t phi_max = (f == 0 || s == 0) ? pi : lm::acos(clampf((sqr(f_k) + sqr(s) - 1) / (2 * f_k * s), -1, 1));
        const f_t phi_f = phi_i + (2 * u2.y - 1) * phi_max;
        const v2 vf = f * v2{ lm::cos(phi_f), lm::sin(phi_f) };
        const v2 zeta = vf;
        const v2 zeta_k = zeta / k;
        const v2 wo = zeta_k - v2{ wi.x, wi.y };
        const f_t z = std::sqrt(std::max(0.f, 1 - dot(wo, wo)));
        const f_t psd = fractal_psd(b, p, zeta, k);
        const f_t w = inv_pi * phi_max;
        const f_t pdf = w > 1e-2f ? z * psd / w : 0.f;
        return { { wo.x, wo.y, wi.z >= 0 ? z : -z }, pdf, psd, w };
    }

    // ---- interface (bsdf.hpp:49-128)
    bool is_delta_only(int32_t id, f_t k) const {
        const wtgpu_bsdf& b = node(id);
        switch (b.type) {
        case WTGPU_BSDF_DIFFUSE: return false;
        case WTGPU_BSDF_DIELECTRIC: return true;
        case WTGPU_BSDF_SURFACE_SPM: return profile_is_delta_only(b, k);
        case WTGPU_BSDF_TWO_SIDED: case WTGPU_BSDF_SCALE: case WTGPU_BSDF_MASK: return is_delta_only(b.child, k);
        case WTGPU_BSDF_COMPOSITE: { const int32_t c = composite_child(b, k); return c < 0 ? true : is_delta_only(c, k); }
        }
        return true;
    }

    static v3 flip(v3 w, f_t z) { return z >= 0 ? w : v3{ w.x, w.y, -w.z }; }      // two_sided.cpp:24-26
    static v3 flip_wo(v3 wo, f_t eta) {                                             // surface_spm.cpp:27-34
        const f_t scale = wo.z > 0 ? eta : 1 / eta;
        const v2 xy = v2{ wo.x, wo.y } * scale;
        const f_t l2 = dot(xy, xy);
        return l2 > 1 ? v3{ 1, 0, 0 } : v3{ xy.x, xy.y, f_t(wo.z > 0 ? -1 : 1) * std::sqrt(std::max(0.f, 1 - l2)) };
    }
    static bool IOR_has_transmission(c_t IOR) { return sqr(std::fabs(IOR.imag())) / std::norm(IOR) <= 1e-2f; }

    c_t spm_IOR(const wtgpu_bsdf& b, f_t k) const { return sc.spectrum_value(b.spec[0], k) / sc.spectrum_value(b.spec[1], k); }
    f_t refl_scale(const wtgpu_bsdf& b, f_t k) const { return b.spec[2] >= 0 ? sc.spectrum_f(b.spec[2], k) : 1.f; }
    f_t trans_scale(const wtgpu_bsdf& b, f_t k) const { return b.spec[3] >= 0 ? sc.spectrum_f(b.spec[3], k) : 1.f; }

    mueller_t f(int32_t id, v3 wi, v3 wo, const bsdf_query_t& q) const {
        const wtgpu_bsdf& b = node(id);
        switch (b.type) {
        case WTGPU_BSDF_DIFFUSE: {      // diffuse.cpp:23-36
            const f_t refl = clampf(sc.spectrum_f(b.spec[0], q.k), 0, 1);
            return ((q.lobes & 1u) && wi.z > 0 && wo.z > 0 ? wo.z * inv_pi * refl : 0.f) * mueller_t::perfect_depolarizer();
        }
        case WTGPU_BSDF_DIELECTRIC: return {};      // dielectric.hpp:99-104
        case WTGPU_BSDF_SURFACE_SPM: {  // surface_spm.cpp:40-77
            const bool is_scatter = (q.lobes & 2u) && !profile_is_delta_only(b, q.k);
            const bool is_reflection = wi.z * wo.z >= 0;
            const c_t eta_12 = spm_IOR(b, q.k);
            const bool has_transmission = IOR_has_transmission(eta_12);
            if (wi.z == 0 || wo.z == 0 || !is_scatter || (!is_reflection && !has_transmission)) return {};
            const v3 abs_wo = is_reflection ? wo : flip_wo(wo, eta_12.real());
            const f_t alpha = profile_alpha(b, wi, abs_wo, q.k);
            f_t J = 1;
            if (!is_reflection && !q.forward) J = sqr(wi.z < 0 ? 1 / eta_12.real() : eta_12.real());
            const f_t scale = is_reflection ? refl_scale(b, q.k) : trans_scale(b, q.k);
            const v3 h = wi + abs_wo;
            const v3 m = normalize(wi.z < 0 ? -h : h);
            const mueller_t F = mueller_fresnel(c_t{ eta_12.real(), 0 }, is_reflection, wi, m);
            const f_t psd = profile_psd(b, wi, abs_wo, q.k);
            return ((1 - alpha) * J * std::fabs(wo.z) * psd * scale) * F;
        }
        case WTGPU_BSDF_TWO_SIDED: return f(b.child, flip(wi, wi.z), flip(wo, wi.z), q);
        case WTGPU_BSDF_SCALE: return sc.spectrum_f(b.spec[0], q.k) * f(b.child, wi, wo, q);      // scale.hpp
        case WTGPU_BSDF_COMPOSITE: { const int32_t c = composite_child(b, q.k); return c < 0 ? mueller_t{} : f(c, wi, wo, q); }
        }
        return {};
    }

    f_t pdf(int32_t id, v3 wi, v3 wo, const bsdf_query_t& q) const {
        const wtgpu_bsdf& b = node(id);
        switch (b.type) {
        case WTGPU_BSDF_DIFFUSE:        // diffuse.cpp:63-71
            return (q.lobes & 1u) && wi.z > 0 && wo.z > 0 ? cosine_hemisphere_pdf(wo.z) : 0.f;
        case WTGPU_BSDF_DIELECTRIC: return 0;
        case WTGPU_BSDF_SURFACE_SPM: {  // surface_spm.cpp:172-200
            const bool is_scatter = (q.lobes & 2u);
            const bool is_reflection = wi.z * wo.z >= 0;
            const c_t eta_12 = spm_IOR(b, q.k);
            const bool has_transmission = IOR_has_transmission(eta_12);
            if (wi.z == 0 || wo.z == 0 || !is_scatter || (!is_reflection && !has_transmission)) return 0;
            const v3 abs_wo = is_reflection ? wo : flip_wo(wo, eta_12.real());
            const f_t alpha = profile_alpha(b, wi, wi, q.k);
            const f_t pdf_specular = (q.lobes & 1u) ? alpha : 0.f;
            const auto fr = fresnel(c_t{ eta_12.real(), 0 }, wi);
            const f_t pdf_transmission = (fr.Ts + fr.Tp) / 2;
            return (1 - pdf_specular) * profile_pdf(b, wi, abs_wo, q.k) * (is_reflection ? 1 - pdf_transmission : pdf_transmission);
        }
        case WTGPU_BSDF_TWO_SIDED: return pdf(b.child, flip(wi, wi.z), flip(wo, wi.z), q);
        case WTGPU_BSDF_SCALE: return pdf(b.child, wi, wo, q);
        case WTGPU_BSDF_COMPOSITE: { const int32_t c = composite_child(b, q.k); return c < 0 ? 0.f : pdf(c, wi, wo, q); }
        }
        return 0;
    }

    std::optional<bsdf_sample_t> sample(int32_t id, v3 wi, const bsdf_query_t& q, sampler_t& sampler) const {
        const wtgpu_bsdf& b = node(id);
        switch (b.type) {
        case WTGPU_BSDF_DIFFUSE: {      // diffuse.cpp:38-61
            if (wi.z <= 0) return std::nullopt;
            const f_t refl = clampf(sc.spectrum_f(b.spec[0], q.k), 0, 1);
            const v3 wo = cosine_hemisphere(sampler.r2());
            return bsdf_sample_t{ wo, pd_t::density(cosine_hemisphere_pdf(wo.z)), { 1, 0 }, refl * mueller_t::perfect_depolarizer() };
        }
        case WTGPU_BSDF_DIELECTRIC: {   // dielectric.cpp:26-72
            const c_t er = sc.spectrum_value(b.spec[0], q.k) / sc.spectrum_value(b.spec[1], q.k);
            const f_t eta_12 = er.real();
            const auto fr = fresnel(c_t{ eta_12, 0 }, wi);
            const f_t T = (fr.Ts + fr.Tp) / 2;
            const bool is_reflection = sampler.r() >= T;
            const v3 wo = is_reflection ? reflect(wi) : fr.t;
            const f_t pdf = is_reflection ? 1 - T : T;
            const f_t scale = is_reflection ? refl_scale(b, q.k) : trans_scale(b, q.k);
            if (scale == 0) return std::nullopt;
            mueller_t M;
            if (is_reflection) M = scale * mueller_t::fresnel(fr.rs, fr.rp);
            else {
                M = (fr.Z * scale) * mueller_t::fresnel(fr.ts, fr.tp);
                if (!q.forward) M = M * (fr.eta_12 * fr.eta_12).real();
            }
            return bsdf_sample_t{ wo, pd_t::discrete(1), fr.eta_12, M / pdf };
        }
        case WTGPU_BSDF_SURFACE_SPM: {  // surface_spm.cpp:79-170
            const f_t alpha = profile_alpha(b, wi, wi, q.k);
            const bool has_specular = (q.lobes & 1u) && alpha > 0;
            const bool has_scatter = (q.lobes & 2u) && alpha < 1;
            const c_t eta_12 = spm_IOR(b, q.k);
            const bool has_transmission = IOR_has_transmission(eta_12);
            if (wi.z == 0 || (!has_specular && !has_scatter)) return std::nullopt;
            f_t pdf = 1;
            bool is_specular = has_specular;
            if (has_specular && has_scatter) {
                const f_t pdf_specular = alpha;
                is_specular = pdf_specular == 1 || sampler.r() < pdf_specular;
                pdf = is_specular ? pdf_specular : 1 - pdf_specular;
            }
            f_t J = 1;
            const auto fr = fresnel(eta_12, wi);
            const f_t pdf_transmission = (fr.Ts + fr.Tp) / 2;
            bool is_reflection = true;
            if (has_transmission) {
                is_reflection = sampler.r() >= pdf_transmission;
                pdf *= is_reflection ? 1 - pdf_transmission : pdf_transmission;
            }
            if (!is_reflection && !q.forward) J = sqr(fr.eta_12.real());
            const f_t scale = is_reflection ? refl_scale(b, q.k) : trans_scale(b, q.k);
            if (scale == 0 || (!is_reflection && !has_transmission)) return std::nullopt;
            if (is_specular) {
                const v3 wo = is_reflection ? reflect(wi) : fr.t;
                const mueller_t F = mueller_fresnel(eta_12, is_reflection, wi);
                const mueller_t M = (alpha * J * scale) * F;
                return bsdf_sample_t{ wo, pd_t::discrete(pdf), is_reflection ? c_t{ 1, 0 } : fr.eta_12, M / pdf };
            }
            const auto smp = profile_sample(b, wi, q.k, sampler);
            const v3 h = wi + smp.wo;
            const v3 m = normalize(wi.z < 0 ? -h : h);
            const mueller_t F = mueller_fresnel(eta_12, is_reflection, wi, m);
            const v3 wo = is_reflection ? smp.wo : flip_wo(smp.wo, eta_12.real());
            pdf *= smp.pdf;
            const mueller_t M = ((1 - alpha) * J * std::fabs(wo.z) * smp.psd * scale) * F;
            return bsdf_sample_t{ wo, pd_t::density(pdf), is_reflection ? c_t{ 1, 0 } : fr.eta_12, M / pdf };
        }
        case WTGPU_BSDF_TWO_SIDED: {    // two_sided.cpp:36-44
            auto s = sample(b.child, flip(wi, wi.z), q, sampler);
            if (s) s->wo = flip(s->wo, wi.z);
            return s;
        }
        case WTGPU_BSDF_SCALE: {
            auto s = sample(b.child, wi, q, sampler);
            if (s) s->M = sc.spectrum_f(b.spec[0], q.k) * s->M;
            return s;
        }
        case WTGPU_BSDF_COMPOSITE: { const int32_t c = composite_child(b, q.k); if (c < 0) return std::nullopt; return sample(c, wi, q, sampler); }
        }
        return std::nullopt;
    }
};

// ================================================================================================
// Emitters
// ================================================================================================
struct emitter_sample_t { beam_t beam; pd_t ppd, dpd; std::optional<surface_t> surface; };
struct emitter_direct_sample_t { int32_t emitter = -1; f_t emitter_pdf = 0; pd_t dpd; beam_t beam; std::optional<surface_t> surface; };

inline v3 mat3_mul(const float* M, v3 v) { return { M[0] * v.x + M[1] * v.y + M[2] * v.z, M[3] * v.x + M[4] * v.y + M[5] * v.z, M[6] * v.x + M[7] * v.y + M[8] * v.z }; }

// discrete_distribution_t::icdf (discrete_distribution.hpp:130-136): lower_bound, -1, clamp, skip empty bins.  dcdf has n + 1 entries.
inline int64_t discrete_icdf(const float* dcdf, uint32_t n, f_t v) {
    const float* it = std::lower_bound(dcdf, dcdf + n + 1, v);
    int64_t idx = std::min<int64_t>(std::max<int64_t>((it - dcdf) - 1, 0), (int64_t)n - 1);
    for (; idx < (int64_t)n - 1 && dcdf[idx + 1] - dcdf[idx] == 0; ++idx) {}
    return idx;
}
// binned_piecewise_linear_distribution_t::icdf (binned_piecewise_linear_distribution.hpp:251-280) -> (x, y); sample() = (x, y*norm) (:285-292).
// The reference walks from a binned guess to the bracketing knot; a binary search finds the same bracket.  ys, dcdf: n entries each.
inline v2 binned_icdf(const float* ys, const float* dcdf, uint32_t n, f_t k0, f_t dk, f_t v) {
    const float* it = std::upper_bound(dcdf, dcdf + n, v);
    uint32_t idx = (uint32_t)std::min<int64_t>(std::max<int64_t>((it - dcdf) - 1, 0), (int64_t)n - 2);
    while (idx + 1 < n - 1 && v > dcdf[idx + 1]) ++idx;
    const f_t f = (v - dcdf[idx]) / (dcdf[idx + 1] - dcdf[idx]);
    const f_t a = ys[idx], b = ys[idx + 1];
    if (a == b) return { ((f_t)idx + f) * dk, a };     // sic: no xrange.min offset (line 270)
    const f_t mm = mix(sqr(a), sqr(b), f);
    const f_t dd = std::sqrt(mm);
    const f_t t = clampf((a - dd) / (a - b), 0, 1);
    return { mix(k0 + (f_t)idx * dk, k0 + (f_t)(idx + 1) * dk, t), mix(a, b, t) };
}
// binned value(x) (binned_piecewise_linear_distribution.hpp:196-205)
inline f_t binned_value(const float* ys, uint32_t n, f_t k0, f_t dk, f_t k) {
    const f_t bin = (k - k0) * (1.f / dk);
    if (bin < 0 || bin > (f_t)(n - 1)) return 0;
    const size_t ii = (size_t)bin;
    const f_t fr = bin - std::floor(bin);
    return mix(ys[ii], ys[std::min<size_t>(n - 1, ii + 1)], fr);
}

struct emitters_t {
    const scene_t& sc;
    explicit emitters_t(const scene_t& s) : sc(s) {}
    const wtgpu_emitter& em(int32_t i) const { return sc.d->emitters[i]; }

    bool is_delta_position(int32_t i) const { const auto t = em(i).type; return t == WTGPU_EMITTER_POINT || t == WTGPU_EMITTER_SPOT; }
    bool is_delta_direction(int32_t i) const { return em(i).type == WTGPU_EMITTER_DIRECTIONAL; }
    bool is_area(int32_t i) const { return em(i).type == WTGPU_EMITTER_AREA; }
    bool is_infinite(int32_t i) const { return em(i).type == WTGPU_EMITTER_DIRECTIONAL; }

    // sourcing geometries: point.hpp:74-88, spot.hpp:115-130, area.hpp:151-165, directional.hpp:118-128
    sourcing_geometry_t sourcing_geometry(const wtgpu_emitter& e, f_t k) const {
        if (e.type == WTGPU_EMITTER_DIRECTIONAL) {
            const auto se = sourcing_geometry_t::source_mub_from_tan_alpha(e.tan_alpha, k).phase_space_extent().enlarge(e.pse_scale);
            return sourcing_geometry_t::source(se);
        }
        const f_t extent = (e.type != WTGPU_EMITTER_AREA && e.extent > 0) ? e.extent : 10.f * wavenum_to_wavelen(k);
        auto se = sourcing_geometry_t::source_mub_from_length(extent, k).phase_space_extent().enlarge(e.pse_scale);
        if (e.type == WTGPU_EMITTER_SPOT) se.tan_alpha = std::min(se.tan_alpha, lm::tan(e.falloff));
        return sourcing_geometry_t::source(se);
    }
    f_t spot_falloff(const wtgpu_emitter& e, v3 local_dir) const {       // spot.hpp:76-81
        const f_t cos_theta = local_dir.z;
        if (cos_theta <= lm::cos(e.cutoff)) return 0;
        if (cos_theta >= lm::cos(e.falloff)) return 1;
        return (e.cutoff - lm::acos(cos_theta)) * (1.f / (e.cutoff - e.falloff));
    }
    f_t area_radiance(const wtgpu_emitter& e, f_t k) const { return e.scale * sc.spectrum_f(e.spectrum, k); }   // area.hpp:104-117
    // area_t::Le (area.hpp:170-180)
    beam_t area_Le(const wtgpu_emitter& e, const ray_t& r, f_t k, const surface_t& surface) const {
        return beam_t::make_forward(r, area_radiance(e, k) * std::max(0.f, dot(r.d, surface.ng())), k, sourcing_geometry(e, k));
    }

    // shape_t::sample_position (src/scene/shape.cpp:70-89)
    struct position_sample_t { v3 p; f_t ppd; surface_t surface; };
    position_sample_t sample_shape_position(int32_t shape, sampler_t& sampler) const {
        const wtgpu_shape& sh = sc.d->shapes[shape];
        const v3 r = sampler.r3();
        const float* cdf = sc.d->shape_tri_cdf + sh.cdf_first;
        // discrete_distribution_t::icdf (discrete_distribution.hpp:102-108): lower_bound, -1, clamp, skip empty bins
        const float* it = std::lower_bound(cdf, cdf + sh.n_tris + 1, r.z);
        int64_t idx = std::min<int64_t>(std::max<int64_t>((it - cdf) - 1, 0), (int64_t)sh.n_tris - 1);
        for (; idx < (int64_t)sh.n_tris - 1 && cdf[idx + 1] - cdf[idx] == 0; ++idx) {}
        const v2 bary = uniform_triangle(v2{ r.x, r.y });
        const uint32_t tuid = sc.d->shape_tri_tuid[sh.tri_first + (uint32_t)idx];
        const surface_t s = sc.make_surface_at_bary(tuid, bary);
        return { s.wp, 1.f / sh.surface_area, s };
    }

    emitter_sample_t sample(int32_t i, sampler_t& sampler, f_t k) const {
        const wtgpu_emitter& e = em(i);
        const v3 pos{ e.pos[0], e.pos[1], e.pos[2] };
        switch (e.type) {
        case WTGPU_EMITTER_POINT: {     // point.cpp:28-41
            const v3 d = uniform_sphere(sampler.r2());
            beam_t b = beam_t::make_forward(ray_t{ pos, d }, sc.spectrum_f(e.spectrum, k), k, sourcing_geometry(e, k));
            b.mul(four_pi);
            return { b, pd_t::discrete(1), pd_t::density(inv_four_pi), std::nullopt };
        }
        case WTGPU_EMITTER_SPOT: {      // spot.cpp:29-46
            const f_t cutoff_sa = two_pi * (1 - lm::cos(e.cutoff));
            const v3 local_wo = uniform_cone(cutoff_sa, sampler.r2());
            const v3 wo = normalize(mat3_mul(e.rot, local_wo));
            const f_t w = spot_falloff(e, local_wo);
            const f_t dpd = uniform_cone_pdf(cutoff_sa);
            beam_t b = beam_t::make_forward(ray_t{ pos, wo }, sc.spectrum_f(e.spectrum, k), k, sourcing_geometry(e, k));
            b.mul(w); b.div(dpd);
            return { b, pd_t::discrete(1), pd_t::density(dpd), std::nullopt };
        }
        case WTGPU_EMITTER_DIRECTIONAL: {   // directional.cpp:28-47
            const v3 dir{ e.dir[0], e.dir[1], e.dir[2] };   // dir_to_emitter
            const frame_t fr = frame_t::build_orthogonal_frame(dir);
            const v2 p = concentric_disk(sampler.r2()) * e.world_radius;
            const v3 wc{ e.world_centre[0], e.world_centre[1], e.world_centre[2] };
            const v3 wp = wc + fr.to_world(p);
            const f_t surface_area = pi * sqr(e.world_radius);
            beam_t b = beam_t::make_forward(ray_t{ wp + e.far_dist * dir, -dir }, sc.spectrum_f(e.spectrum, k), k, sourcing_geometry(e, k));
            b.mul(surface_area);
            return { b, pd_t::density(1.f / surface_area), pd_t::discrete(1), std::nullopt };
        }
        case WTGPU_EMITTER_AREA: {      // area.cpp:52-80
            const auto ps = sample_shape_position(e.shape, sampler);
            v3 d = cosine_hemisphere(sampler.r2());
            const f_t dn = d.z;
            d = ps.surface.geo.to_world(d);
            const f_t dpd = cosine_hemisphere_pdf(dn);
            const f_t ppd = ps.ppd;
            f_t recp_pdf = 1.f / (dpd * ppd);
            if (dpd * ppd == 0) recp_pdf = 0;
            beam_t b = area_Le(e, ray_t{ ps.p, d }, k, ps.surface);
            b.mul(recp_pdf);
            return { b, pd_t::density(ppd), pd_t::density(dpd), ps.surface };
        }
        }
        return {};
    }

    f_t pdf_position_density(int32_t i) const {      // density_or_zero of pdf_position
        const wtgpu_emitter& e = em(i);
        if (e.type == WTGPU_EMITTER_AREA) return 1.f / sc.d->shapes[e.shape].surface_area;
        return 0;
    }
    // emitter_t::pdf_direction (density_or_zero)
    f_t pdf_direction_density(int32_t i, v3 dir, const surface_t* surface) const {
        const wtgpu_emitter& e = em(i);
        switch (e.type) {
        case WTGPU_EMITTER_POINT: return inv_four_pi;
        case WTGPU_EMITTER_SPOT: return uniform_cone_pdf(two_pi * (1 - lm::cos(e.cutoff)));
        case WTGPU_EMITTER_AREA: return cosine_hemisphere_pdf(std::max(0.f, dot(dir, surface->ng())));
        default: return 0;
        }
    }
    // area_t::pdf_direct (area.cpp:130-141)
    f_t area_pdf_direct(const wtgpu_emitter& e, v3 wp, const ray_t& r, const surface_t& surface) const {
        const f_t ppd = 1.f / sc.d->shapes[e.shape].surface_area;
        const f_t l2 = length2(wp - r.o);
        const f_t dn = std::max(0.f, dot(r.d, surface.ng()));
        const f_t recp_dn = dn > 0 ? 1 / dn : 0.f;
        return ppd * l2 * recp_dn;
    }

    emitter_direct_sample_t sample_direct(int32_t i, sampler_t& sampler, v3 wp, f_t k) const {
        const wtgpu_emitter& e = em(i);
        const v3 pos{ e.pos[0], e.pos[1], e.pos[2] };
        emitter_direct_sample_t ret; ret.emitter = i;
        switch (e.type) {
        case WTGPU_EMITTER_POINT: {     // point.cpp:43-60
            const v3 dl = wp - pos;
            const f_t recp_dist2 = 1 / length2(dl);
            const v3 d = dl * std::sqrt(recp_dist2);
            ret.beam = beam_t::make_forward(ray_t{ pos, d }, sc.spectrum_f(e.spectrum, k), k, sourcing_geometry(e, k));
            ret.beam.mul(recp_dist2);
            ret.dpd = pd_t::discrete(1);
            return ret;
        }
        case WTGPU_EMITTER_SPOT: {      // spot.cpp:48-67
            const v3 dl = wp - pos;
            const f_t recp_dist2 = 1 / length2(dl);
            const v3 d = dl * std::sqrt(recp_dist2);
            const v3 local_wo = normalize(mat3_mul(e.inv_rot, d));
            const f_t w = spot_falloff(e, local_wo);
            ret.beam = beam_t::make_forward(ray_t{ pos, d }, sc.spectrum_f(e.spectrum, k), k, sourcing_geometry(e, k));
            ret.beam.mul(w); ret.beam.mul(recp_dist2);
            ret.dpd = pd_t::discrete(1);
            return ret;
        }
        case WTGPU_EMITTER_DIRECTIONAL: {   // directional.cpp:49-72
            const v3 dir{ e.dir[0], e.dir[1], e.dir[2] };
            const frame_t fr = frame_t::build_orthogonal_frame(dir);
            const v3 wc{ e.world_centre[0], e.world_centre[1], e.world_centre[2] };
            const v3 pl = fr.to_local(wp - wc);
            const v2 p{ pl.x, pl.y };
            const f_t r2 = length2(p);
            const f_t scale = r2 <= sqr(e.world_radius) ? 1.f : 0.f;
            const v3 targetwp = wc + fr.to_world(p);
            ret.beam = beam_t::make_forward(ray_t{ targetwp + e.far_dist * dir, -dir }, sc.spectrum_f(e.spectrum, k), k, sourcing_geometry(e, k));
            ret.beam.mul(scale);
            ret.dpd = pd_t::discrete(1);
            return ret;
        }
        case WTGPU_EMITTER_AREA: {      // area.cpp:82-105
            const auto ps = sample_shape_position(e.shape, sampler);
            const v3 d = normalize(wp - ps.p);
            const f_t dpd = area_pdf_direct(e, wp, ray_t{ ps.p, d }, ps.surface);
            const f_t recp_dpd = dpd > 0 ? 1 / dpd : 0.f;
            ret.beam = area_Le(e, ray_t{ ps.p, d }, k, ps.surface);
            ret.beam.mul(recp_dpd);
            ret.dpd = pd_t::density(dpd);
            ret.surface = ps.surface;
            return ret;
        }
        }
        return ret;
    }

    // area_t::Li (area.cpp:35-50)
    stokes_t Li(int32_t i, const beam_t& Sbeam, const surface_t* surface) const {
        const wtgpu_emitter& e = em(i);
        if (e.type != WTGPU_EMITTER_AREA || !surface) return {};
        const f_t dn = dot(-Sbeam.dir(), surface->ng());
        if (dn <= 0) return {};
        beam_t Ibeam = area_Le(e, ray_t{ surface->wp, -Sbeam.dir() }, Sbeam.k, *surface);
        Ibeam.div(dn);
        return integrate_beams(Sbeam, Ibeam);
    }

    // ---- scene-level sampling (scene.hpp:96-200, scene_sensor.cpp:19-59)
    int32_t sample_emitter(sampler_t& sampler) const { return (int32_t)discrete_icdf(sc.d->emitter_cdf, sc.d->n_emitters, sampler.r()); }
    f_t pdf_emitter(int32_t i) const { return sc.d->emitter_cdf[i + 1] - sc.d->emitter_cdf[i]; }

    struct wavenumber_sample_t { f_t k; pd_t wpd; };
    // distribution1d_t::sample of the emitter x sensor product spectrum
    wavenumber_sample_t sample_wavenumber(int32_t i, sampler_t& sampler) const {
        const wtgpu_kdist& kd = sc.d->emitter_kdist[i];
        const float* data = sc.d->kdist_data + kd.first;
        const f_t v = sampler.r();
        const uint32_t n = kd.n;
        if (kd.type == WTGPU_KDIST_DISCRETE) {
            // discrete_distribution_t<vec2_t>::icdf/sample (discrete_distribution.hpp:258-272)
            const float* ks = data; const float* ys = data + n; const float* dcdf = data + 2 * n;
            const int64_t idx = discrete_icdf(dcdf, n, v);
            return { ks[idx], pd_t::discrete(ys[idx] * kd.norm) };
        }
        const v2 xy = binned_icdf(data, data + n, n, kd.k0, kd.dk, v);
        return { xy.x, pd_t::density(xy.y * kd.norm) };
    }
    // emitter_sampling_data_t::pdf_wavenumber (scene_sensor.hpp:63-70)
    f_t pdf_wavenumber(int32_t i, f_t k) const {
        const wtgpu_kdist& kd = sc.d->emitter_kdist[i];
        const float* data = sc.d->kdist_data + kd.first;
        const uint32_t n = kd.n;
        if (kd.type == WTGPU_KDIST_DISCRETE) {      // discrete_distribution.hpp:241-248: mass of an exactly matching line
            const float* ks = data; const float* dcdf = data + 2 * n;
            const float* it = std::lower_bound(ks, ks + n, k);
            if (it == ks + n || *it != k) return 0;
            const size_t idx = it - ks;
            return dcdf[idx + 1] - dcdf[idx];
        }
        return binned_value(data, n, kd.k0, kd.dk, k) * kd.norm;       // pdf = value(x)*norm (binned_piecewise_linear_distribution.hpp:241-244)
    }
    f_t sum_spectral_pdf_for_all_emitters(f_t k) const {                // scene_sensor.hpp:115-123
        f_t s = 0;
        for (uint32_t i = 0; i < sc.d->n_emitters; ++i) s += pdf_emitter((int32_t)i) * pdf_wavenumber((int32_t)i, k);
        return s;
    }
    // scene_t::sample_emitter_direct (scene.hpp:128-141)
    emitter_direct_sample_t sample_emitter_direct(sampler_t& sampler, v3 wp, f_t k) const {
        const int32_t e = sample_emitter(sampler);
        const f_t pd = pdf_emitter(e);
        auto s = sample_direct(e, sampler, wp, k);
        s.emitter_pdf = pd;
        s.beam.div(pd);
        return s;
    }
};

// ================================================================================================
// Sensors + film
// ================================================================================================
struct element_sample_t { uint32_t ex = 0, ey = 0; f_t ox = 0, oy = 0; };
struct sensor_sample_t { beam_t beam; pd_t ppd, dpd; element_sample_t element; std::optional<surface_t> surface; };
struct sensor_direct_sample_t { beam_t beam; pd_t dpd; element_sample_t element; std::optional<surface_t> surface; };
struct sensor_direct_connection_t { beam_t beam; element_sample_t element; std::optional<surface_t> surface; };

inline void mat4_mul(const float* M, const float v[4], float out[4]) {
    for (int r = 0; r < 4; ++r) out[r] = M[4 * r] * v[0] + M[4 * r + 1] * v[1] + M[4 * r + 2] * v[2] + M[4 * r + 3] * v[3];
}

struct sensor_eval_t {
    const scene_t& sc;
    const wtgpu_sensor& s;
    explicit sensor_eval_t(const scene_t& scn) : sc(scn), s(scn.d->sensor) {}

    bool is_virtual() const { return s.type == WTGPU_SENSOR_VIRTUAL_PLANE; }
    bool is_delta_position() const { return s.type == WTGPU_SENSOR_PERSPECTIVE; }
    bool is_delta_direction() const { return false; }

    // ---- perspective (sensor/perspective.hpp)
    static constexpr f_t image_plane_z = 0.01f;     // 1 cm
    static constexpr f_t beam_source_spatial_stddev = .25f;
    v3 persp_point_on_sensor(v2 film_pos) const {    // perspective.hpp:66-70
        const float v[4] = { film_pos.x, film_pos.y, 1, 1 }; float p[4];
        mat4_mul(s.s2c, v, p);
        return v3{ p[0], p[1], p[2] } / p[3];
    }
    v2 persp_point_on_film(v3 dir) const {           // perspective.hpp:73-77
        const v3 p = dir / std::fabs(dir.z);
        const float v[4] = { p.x, p.y, 1, 1 }; float q[4];
        mat4_mul(s.c2s, v, q);
        return v2{ q[0], q[1] } / q[3];
    }
    v3 persp_pos() const { return { s.pos[0], s.pos[1], s.pos[2] }; }
    v2 persp_sensor_extent() const {                  // perspective.hpp:139-145
        const v3 f0 = persp_point_on_sensor({ 0, 0 });
        const v3 fW = persp_point_on_sensor({ (f_t)s.width, 0 });
        const v3 fH = persp_point_on_sensor({ 0, (f_t)s.height });
        return { length(fW - f0), length(fH - f0) };
    }
    f_t persp_recp_sa_density(v3 d) const {           // perspective.hpp:183-185
        const v2 e = persp_sensor_extent();
        return (e.x * e.y) / sqr(image_plane_z) * (d.z * d.z * d.z);
    }
    sourcing_geometry_t persp_sourcing(f_t k) const { // perspective.hpp:190-206
        const v2 e = persp_sensor_extent();
        const f_t elem_x = e.x / (f_t)s.width;
        const f_t ise = elem_x * beam_source_spatial_stddev * beam_cross_section_envelope;
        const auto se = sourcing_geometry_t::source(ise, s.sourcing_tan_alpha, k).phase_space_extent().enlarge(s.pse_scale);
        return sourcing_geometry_t::source(se);
    }
    beam_t persp_Se(const ray_t& r, f_t recp_sa_density, f_t k) const { // perspective.hpp:88-99, 219-223
        const f_t J = 1.f / recp_sa_density;
        return beam_t::make_backward(r, J, k, persp_sourcing(k));
    }

    // ---- virtual plane (sensor/virtual_plane_sensor.hpp, src/sensor/virtual_plane_sensor.cpp)
    frame_t vp_frame() const { return { { s.frame_t[0], s.frame_t[1], s.frame_t[2] }, { s.frame_b[0], s.frame_b[1], s.frame_b[2] }, { s.frame_n[0], s.frame_n[1], s.frame_n[2] } }; }
    v3 vp_origin() const { return { s.origin[0], s.origin[1], s.origin[2] }; }
    v2 vp_extent() const { return { s.extent[0], s.extent[1] }; }
    v2 vp_elem_extent() const { return { s.extent[0] / (f_t)s.width, s.extent[1] / (f_t)s.height }; }
    f_t vp_area() const { return s.extent[0] * s.extent[1]; }
    sourcing_geometry_t vp_sourcing(f_t k) const {   // virtual_plane_sensor.hpp:137-153
        const v2 ee = vp_elem_extent();
        const f_t ise = (ee.x + ee.y) / 2 * beam_source_spatial_stddev * beam_cross_section_envelope;
        if (s.requested_tan_alpha >= 0) return sourcing_geometry_t::source(ise, s.requested_tan_alpha, k);
        return sourcing_geometry_t::source_mub_from_length(ise, k);
    }
    beam_t vp_Se(const ray_t& r, f_t k) const {      // virtual_plane_sensor.hpp:165-183
        const f_t W = 1.f / pi * (1.f / vp_area());
        const f_t dn = std::max(0.f, dot(r.d, vp_frame().n));
        return beam_t::make_backward(r, W * dn, k, vp_sourcing(k));
    }
    element_sample_t vp_element_for_position(v3 wp) const {     // virtual_plane_sensor.hpp:114-126
        const v3 sp = wp - vp_origin();
        const frame_t f = vp_frame();
        const v2 ee = vp_elem_extent();
        const v2 efp{ dot(sp, f.t) * (1.f / ee.x), dot(sp, f.b) * (1.f / ee.y) };
        const uint32_t ex = (uint32_t)efp.x, ey = (uint32_t)efp.y;
        return { ex, ey, efp.x - (f_t)ex - .5f, efp.y - (f_t)ey - .5f };
    }

    // ---- interface
    sensor_sample_t sample(sampler_t& sampler, uint32_t ex, uint32_t ey, f_t k) const {
        if (s.type == WTGPU_SENSOR_PERSPECTIVE) {   // perspective.hpp:229-269
            const v3 centre = persp_point_on_sensor(v2{ (f_t)ex, (f_t)ey } + v2{ .5f, .5f });
            const v2 off = sampler.r2() - v2{ .5f, .5f };
            const v3 ddx = persp_point_on_sensor({ 1, 0 }) - persp_point_on_sensor({ 0, 0 });
            const v3 ddy = persp_point_on_sensor({ 0, 1 }) - persp_point_on_sensor({ 0, 0 });
            const v3 dir_local = normalize(centre + off.x * ddx + off.y * ddy);
            const v3 dir = normalize(mat3_mul(s.rot, dir_local));
            const f_t recp_dpd = persp_recp_sa_density(dir_local);
            beam_t b = persp_Se(ray_t{ persp_pos(), dir }, recp_dpd, k);
            b.mul(recp_dpd);
            return { b, pd_t::discrete(1), pd_t::density(1 / recp_dpd), { ex, ey, off.x, off.y }, std::nullopt };
        }
        // virtual_plane_sensor.cpp:101-132
        const v2 r2 = sampler.r2();
        const v2 off = r2 - v2{ .5f, .5f };
        const v2 ee = vp_elem_extent();
        // position_for_element (virtual_plane_sensor.hpp:101-107): "+.5" is a double literal there
        const v2 local = v2{ (f_t)((double)((f_t)ex + off.x) + .5), (f_t)((double)((f_t)ey + off.y) + .5) } * ee;
        const frame_t f = vp_frame();
        const v3 p = vp_origin() + local.x * f.t + local.y * f.b;
        const f_t recp_ppd = vp_area();
        const v3 wo = cosine_hemisphere(sampler.r2());
        const f_t dpd = cosine_hemisphere_pdf(wo.z);
        beam_t b = vp_Se(ray_t{ p, f.to_world(wo) }, k);
        b.mul(recp_ppd); b.mul(dpd > 0 ? 1 / dpd : 0.f);
        return { b, pd_t::density(1 / recp_ppd), pd_t::density(dpd), { ex, ey, off.x, off.y }, scene_t::make_dummy_surface(f.n, p) };
    }

    sensor_direct_sample_t sample_direct(sampler_t& sampler, v3 wp, f_t k) const {
        if (s.type == WTGPU_SENSOR_PERSPECTIVE) {   // perspective.hpp:274-314
            const v3 wdl = wp - persp_pos();
            const f_t recp_dist2 = 1 / length2(wdl);
            const v3 wd = wdl * std::sqrt(recp_dist2);
            const v3 dir_local = normalize(mat3_mul(s.inv_rot, wd));
            const v2 fp = persp_point_on_film(dir_local);
            const f_t recp_sa = persp_recp_sa_density(dir_local);
            const bool inside = dir_local.z > std::numeric_limits<f_t>::epsilon() && fp.x >= 0 && fp.y >= 0 && fp.x < (f_t)s.width && fp.y < (f_t)s.height;
            const uint32_t ex = inside ? (uint32_t)fp.x : 0, ey = inside ? (uint32_t)fp.y : 0;
            const v2 off{ (fp.x - std::floor(fp.x)) - .5f, (fp.y - std::floor(fp.y)) - .5f };
            beam_t b = persp_Se(ray_t{ persp_pos(), wd }, recp_sa, k);
            b.mul(recp_dist2); b.mul(inside ? 1.f : 0.f);
            return { b, pd_t::discrete(1), { ex, ey, off.x, off.y }, std::nullopt };
        }
        // virtual_plane_sensor.cpp:134-176
        const frame_t f = vp_frame();
        const v2 splocal = sampler.r2() * vp_extent();
        const v3 sp = vp_origin() + splocal.x * f.t + splocal.y * f.b;
        const v2 ee = vp_elem_extent();
        const v2 efp{ splocal.x / ee.x, splocal.y / ee.y };
        const uint32_t ex = (uint32_t)efp.x, ey = (uint32_t)efp.y;
        const v2 off{ efp.x - (f_t)ex - .5f, efp.y - (f_t)ey - .5f };
        const v3 wdl = wp - sp;
        const f_t dist2 = length2(wdl);
        const v3 wd = wdl / std::sqrt(dist2);
        const v3 wd_local = f.to_local(wd);
        const f_t recp_dn = wd_local.z > 0 ? 1 / wd_local.z : 0.f;
        const f_t dpd = (1.f / vp_area()) * dist2 * recp_dn;
        const f_t recp_dpd = dpd > 0 ? 1 / dpd : 0.f;
        beam_t b = vp_Se(ray_t{ sp, wd }, k);
        b.mul(recp_dpd); b.mul(recp_dn);
        return { b, pd_t::density(dpd), { ex, ey, off.x, off.y }, scene_t::make_dummy_surface(f.n, sp) };
    }

    // virtual_plane_sensor_t::Si (virtual_plane_sensor.cpp:65-99)
    std::optional<sensor_direct_connection_t> Si(const beam_t& beam, range_t range) const {
        if (!is_virtual()) return std::nullopt;
        const frame_t f = vp_frame();
        const v3 n = f.n;
        const f_t dn = dot(-beam.dir(), n);
        if (dn <= 0) return std::nullopt;
        const v3 o = vp_origin(); const v2 ext = vp_extent();
        const v3 a = o, b = o + ext.x * f.t, c = o + ext.y * f.b, dd = o + ext.x * f.t + ext.y * f.b;
        const ray_t& ray = beam.envelope.r;
        const auto i1 = intersect_ray_tri(ray, a, b, c, range);
        const auto i2 = intersect_ray_tri(ray, c, b, dd, range);
        if (!i1 && !i2) return std::nullopt;
        const v3 p = i1 ? ray.propagate(i1->dist) : ray.propagate(i2->dist);
        beam_t se = vp_Se(ray_t{ p, -beam.dir() }, beam.k);
        se.div(dn);
        return sensor_direct_connection_t{ se, vp_element_for_position(p), scene_t::make_dummy_surface(n, p) };
    }

    f_t pdf_position_density() const { return is_virtual() ? 1.f / vp_area() : 0.f; }
    f_t pdf_direction_density(v3 dir) const {
        if (is_virtual()) return cosine_hemisphere_pdf(std::max(vp_frame().to_local(dir).z, 0.f));   // virtual_plane_sensor.cpp:182-186
        const v3 d = normalize(mat3_mul(s.inv_rot, dir));                                             // perspective.hpp:326-333
        return d.z > std::numeric_limits<f_t>::epsilon() ? 1 / persp_recp_sa_density(d) : 0.f;
    }
};

// erf lookup table: include/wt/math/erf_lut.hpp:20-55
struct erf_lut_t {
    static constexpr int N = 1024;
    static constexpr f_t maxX = 3.5f;
    f_t lut[N];
    erf_lut_t() { for (int i = 0; i < N; ++i) lut[i] = std::erf((f_t)i / (N - 1) * maxX); }
    f_t operator()(f_t x) const {
        const f_t sgn = sign(x);
        x = std::fabs(x) * ((f_t)(N - 1) / maxX);
        const f_t fr = x - std::floor(x);
        const size_t idx0 = (size_t)x, idx1 = idx0 + 1;
        return idx1 >= (size_t)N || idx1 == 0 ? sgn : sgn * mix(lut[idx0], lut[idx1], fr);
    }
};
inline const erf_lut_t& erf_lut() { static erf_lut_t l; return l; }

// film_t / film_storage_t (sensor/film/film.hpp:214-340, film_storage.hpp:196-291): double accumulation
// gaussian1d_t::integrate over a range (math/distribution/gaussian1d.hpp:100-106), mu = 0: the reconstruction filter's mass over one pixel
inline f_t gaussian1d_integrate(f_t sigma, f_t mn, f_t mx) {
    if (sigma == 0) return (mn <= 0 && 0 <= mx) ? 1.f : 0.f;
    const f_t n = inv_sqrt_two * (1.f / sigma);
    return (erf_lut()(mx * n) - erf_lut()(mn * n)) / 2;
}
struct film_t {
    const scene_t& sc;
    uint32_t W, H, C;
    int r;
    f_t sigma;
    std::vector<double> block;      // [y][x][c][2]
    std::vector<double> light;      // [y][x][c]
    explicit film_t(const scene_t& s) : sc(s), W(s.d->sensor.width), H(s.d->sensor.height), C(s.d->sensor.channels),
        r((int)s.d->sensor.rf_radius), sigma(s.d->sensor.rfilter_stddev), block((size_t)W * H * C * 2, 0.0), light((size_t)W * H * C, 0.0) {}

    // gaussian1d_t::integrate (math/distribution/gaussian1d.hpp:100-106), mu = 0
    f_t rf_integrate(f_t mn, f_t mx) const { return gaussian1d_integrate(sigma, mn, mx); }
    // film.hpp:308-340; weights ordered x-major (for_range: last dimension fastest)
    f_t weights(f_t ox, f_t oy, f_t* w) const {
        const int Wd = 2 * r + 1;
        f_t tx[16], ty[16];
        for (int x = -r; x <= r; ++x) { tx[x + r] = rf_integrate(x + ox - .5f, x + ox + .5f); ty[x + r] = rf_integrate(x + oy - .5f, x + oy + .5f); }
        f_t tw = 0; int i = 0;
        for (int x = 0; x < Wd; ++x) for (int y = 0; y < Wd; ++y) { const f_t v = std::max(0.f, 1.f * tx[x] * ty[y]); tw += v; w[i++] = v; }
        return tw > 0 ? 1.f / tw : 0.f;
    }
    f_t response(uint32_t c, f_t k) const { return sc.spectrum_f(sc.d->sensor.response[c], k); }

    // film_t::splat (film.hpp:254-288) + film_storage_t::write_block (film_storage.hpp:196-222)
    void splat(const element_sample_t& e, const stokes_t& L, f_t k) {
        f_t w[64]; const f_t rtw = weights(e.ox, e.oy, w);
        for (uint32_t c = 0; c < C; ++c) {
            f_t val = L.intensity() * response(c, k);
            val = (val >= 0 && std::isfinite(val)) ? val : 0.f;
            int i = 0;
            for (int dx = -r; dx <= r; ++dx) for (int dy = -r; dy <= r; ++dy) {
                const f_t ww = w[i++] * rtw;
                const int64_t px = (int64_t)e.ex + dx, py = (int64_t)e.ey + dy;
                if (px < 0 || py < 0 || px >= W || py >= H) continue;
                double* p = &block[(((size_t)py * W + px) * C + c) * 2];
                p[0] += (double)(ww * val); p[1] += (double)ww;
            }
        }
    }
    // film_t::splat_direct (film.hpp:214-252) + write_light_splat (film_storage.hpp:224-245)
    void splat_direct(const element_sample_t& e, const stokes_t& L, f_t k) {
        f_t w[64]; const f_t rtw = weights(e.ox, e.oy, w);
        for (uint32_t c = 0; c < C; ++c) {
            const f_t val = L.intensity() * response(c, k);
            if (val <= 0 || !std::isfinite(val)) continue;
            int i = 0;
            for (int dx = -r; dx <= r; ++dx) for (int dy = -r; dy <= r; ++dy) {
                const f_t ww = w[i++] * rtw;
                const int64_t px = (int64_t)e.ex + dx, py = (int64_t)e.ey + dy;
                if (px < 0 || py < 0 || px >= W || py >= H) continue;
                light[((size_t)py * W + px) * C + c] += (double)(ww * val);
            }
        }
    }
    void merge(const film_t& o) { for (size_t i = 0; i < block.size(); ++i) block[i] += o.block[i]; for (size_t i = 0; i < light.size(); ++i) light[i] += o.light[i]; }
};

} // namespace ot
