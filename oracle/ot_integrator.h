// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// ot_integrator.h: traverse() state machine, UTD free-space diffraction and the plt_path integrator, restated from
//   include/wt/integrator/traversal.hpp, include/wt/interaction/fsd/{utd.hpp,free_space_diffraction.hpp},
//   src/interaction/fsd/free_space_diffraction.cpp, include/wt/integrator/plt_path/plt_path_detail.hpp.
#pragma once
#include "ot_scene.h"

namespace ot {

// vertex_geo_variant_t (traversal.hpp:251) + intersection_edge_t
struct geo_t {
    enum kind_e { NONE, POINT, SURFACE, EDGE } kind = NONE;
    v3 p{};
    surface_t s{};
    uint32_t edge = WTGPU_INVALID_IDX;
    static geo_t point(v3 p) { geo_t g; g.kind = POINT; g.p = p; return g; }
    static geo_t surface(const surface_t& s) { geo_t g; g.kind = SURFACE; g.s = s; g.p = s.wp; return g; }
    static geo_t on_edge(uint32_t e, v3 p) { geo_t g; g.kind = EDGE; g.edge = e; g.p = p; return g; }
    v3 position() const { return p; }
};
inline v3 offseted_ray_origin(const scene_t& sc, const geo_t& g, const ray_t& ray) {
    if (g.kind == geo_t::SURFACE) return sc.offseted_ray_origin(g.s, ray);
    if (g.kind == geo_t::EDGE) return sc.offseted_ray_origin_edge(g.edge, ray);
    return ray.o;
}

static constexpr f_t ballistic_scale = 1.001f;      // traversal.hpp:26

// traversal.hpp:28-37
inline f_t calculate_min_ballistic_distance(const elliptic_cone_t& envelope, const ray_t& ray) {
    if (ray.o != envelope.o()) {
        const v3 rl = envelope.frame().to_local(ray.o - envelope.o()) * v3{ 1, envelope.e, 1 };
        const f_t d = (length(v2{ rl.x, rl.y }) - envelope.x0) / envelope.tan_alpha - rl.z;
        return max3(0.f, -rl.z, d);
    }
    return 0;
}
// traversal.hpp:39-57
inline f_t max_ballistic_distance(f_t lambda, uint32_t segment, f_t min_ballistic_distance) {
    const f_t min_dist = min_ballistic_distance * 1.05f;
    const uint64_t B = std::min<uint64_t>(1ull << 16, segment >= 31 ? (1ull << 16) : (8ull << (2 * segment + 1)));
    return segment >= 16 ? inf : min_dist + lambda * (f_t)B;
}

struct traversal_result_t {
    v3 origin{};
    bool ballistic = false;
    bool empty = true;
    ray_hit_t ray;          // ballistic record
    cone_record_t cone;     // diffusive record
    f_t intersection_region_depth = 0;
    f_t distance() const { return ballistic ? ray.dist : cone.dist; }
};

// traversal.hpp:94-172
inline traversal_result_t traverse(const scene_t& sc, const elliptic_cone_t& envelope, f_t lambda, f_t distance,
                                   bool force_ray_tracing, bool detect_edges, ads_counters_t* ctr) {
    const ray_t& ray = envelope.r;
    traversal_result_t res;
    if (force_ray_tracing || envelope.is_ray()) {
        res.origin = ray.o; res.ballistic = true;
        res.ray = intersect_ray(sc.ads, ray, { 0, distance }, ctr);
        res.empty = res.ray.empty();
        return res;
    }
    const f_t min_ballistic_distance = calculate_min_ballistic_distance(envelope, ray);
    const f_t z_search_range = major_axis_to_z_scale;
    f_t dist = 0;
    for (uint32_t seg = 0;; ++seg) {
        const f_t ballistic_dist = max_ballistic_distance(lambda, seg, min_ballistic_distance);
        const auto bl = intersect_ray(sc.ads, ray, { dist, std::min(distance, dist + ballistic_dist * ballistic_scale) }, ctr);
        if (!bl.empty()) { res.origin = ray.o; res.ballistic = true; res.ray = bl; res.empty = false; return res; }
        dist += ballistic_dist;
        if (ballistic_dist == inf || dist >= distance) { res.origin = ray.o; res.ballistic = true; res.empty = true; return res; }
        const f_t min_df_prog = envelope.axes(dist).x / 2.f;
        auto df = intersect_cone(sc.ads, envelope, { dist, distance }, z_search_range, detect_edges, ctr);
#ifdef ORACLE_CONE_HOOK
        ORACLE_CONE_HOOK(envelope, dist, df);
#endif
        if (df.empty() || df.dist - dist >= min_df_prog) {
            res.origin = envelope.o(); res.ballistic = false;
            res.empty = df.empty();
            res.intersection_region_depth = df.empty() ? 0.f : z_search_range * envelope.axes(df.dist).x;
            res.cone = std::move(df);
            return res;
        }
    }
}
// traversal.hpp:276-286
inline traversal_result_t traverse(const scene_t& sc, const elliptic_cone_t& cone, const geo_t& intrs, f_t lambda,
                                   bool force_rt, bool detect_edges, ads_counters_t* ctr) {
    elliptic_cone_t envelope = cone;
    envelope.set_o(offseted_ray_origin(sc, intrs, cone.r));
    return traverse(sc, envelope, lambda, inf, force_rt, detect_edges, ctr);
}
// traversal.hpp:319-333
inline bool shadow(const scene_t& sc, const geo_t& start, const geo_t& end, ads_counters_t* ctr) {
    const v3 start_wp = start.position(), end_wp = end.position();
    const ray_t ray{ start_wp, normalize(end_wp - start_wp) };
    const v3 o = offseted_ray_origin(sc, start, ray);
    const v3 t = offseted_ray_origin(sc, end, ray_t{ end_wp, -ray.d });
    const f_t dist = length(t - o);
    const v3 d = (t - o) / dist;
    return shadow_ray(sc.ads, ray_t{ o, d }, { 0, dist }, ctr);
}

// ================================================================================================
// UTD free-space diffraction
// ================================================================================================
static constexpr f_t utd_min_sin_beta = 1e-3f;
static constexpr f_t utd_IS_sigma_scale = 45;

// Complex erfc(e^{i pi/4} s), s>=0 real -- the only cerfc call of the path (utd.hpp:42).  libcerf (deps/libcerf @ 09b98c1)
// is not vendored; restated through the Fresnel integrals: erf(e^{i pi/4} s) = (2/sqrt(pi)) e^{i pi/4} (C~(s) - i S~(s)),
// C~(s)=int_0^s cos(u^2)du, S~(s)=int_0^s sin(u^2)du, evaluated by their power series in double precision.
inline std::complex<double> cerfc_rot45(double s) {
    const double s2 = s * s;
    double C = 0, S = 0;
    // C~ = sum (-1)^n s^(4n+1)/((2n)! (4n+1)),  S~ = sum (-1)^n s^(4n+3)/((2n+1)! (4n+3))
    double term = s;    // s^(2m+1)/m!  for m = 0
    for (int m = 0; m < 200; ++m) {
        const double contrib = term / (2.0 * m + 1.0);
        const int n = m / 2;
        if (m % 2 == 0) C += (n % 2 ? -contrib : contrib);
        else S += (n % 2 ? -contrib : contrib);
        term *= s2 / (m + 1.0);
        if (std::fabs(term) < 1e-30 && m > 4) break;
    }
    const std::complex<double> rot = std::exp(std::complex<double>(0, 0.78539816339744830962));
    const std::complex<double> erf = (2.0 / 1.7724538509055160273) * rot * std::complex<double>(C, -S);
    return 1.0 - erf;
}

// erfc(z) for a general complex z with |z| < 2.5: Maclaurin series of erf in double precision (terms peak below e^6, so ~1e-13 absolute).
// UTDF's argument is NOT exactly on the 45-degree ray: the reference forms exp(i pi/4) * sqrt(x) in complex<float> (utd.hpp:42) and libcerf
// takes that f32-rounded point; evaluating on the exact ray instead moves F by up to 6e-7 relative (found by the pin against the reference's
// own utd.hpp, tests/test_oracle_kats.py::test_utd_equals_the_reference_code).
inline std::complex<double> cerfc_series(std::complex<double> z) {
    const std::complex<double> z2 = z * z;
    std::complex<double> term = z, sum = z;     // term = (-1)^n z^(2n+1) / n!
    for (int n = 1; n < 200; ++n) {
        term *= -z2 / (double)n;
        sum += term / (2.0 * n + 1.0);
        if (std::abs(term) < 1e-30 && n > 4) break;
    }
    return 1.0 - (2.0 / 1.7724538509055160273) * sum;
}

// utd.hpp:36-57
inline c_t UTDF(f_t x) {
    const f_t absx = std::fabs(x);
    c_t result;
    if (absx < 6) {
        const f_t sqrt_x = std::sqrt(absx);
        const c_t zf = lm::expi(pi_4) * sqrt_x;
        const std::complex<double> ce = cerfc_series(std::complex<double>(zf.real(), zf.imag()));
        const c_t cerf{ (f_t)ce.real(), (f_t)ce.imag() };
        result = c_t{ 1, 1 } * sqrt_pi_2 * sqrt_x * lm::expi(absx) * cerf;
    } else {
        const f_t r = 1 / (2 * absx);
        const f_t r2 = r * r, r3 = r2 * r, r4 = r2 * r2;
        result = 1.f + c_t{ 0, 1 } * r - 3 * r2 - c_t{ 0, 15 } * r3 + 75 * r4;
    }
    return x < 0 ? std::conj(result) : result;
}
// utd.hpp:26-31
inline f_t UTDa(int sgn, f_t phi, f_t n) {
    const f_t N = std::round((f_t)(sgn * pi + phi) * inv_two_pi / n);
    return 2 * sqr(lm::cos(pi * n * N - phi / 2));
}
inline f_t fmod_pos(f_t a, f_t b) { return a - b * std::floor(a / b); }    // glm::mod

struct wedge_edge_t {       // interaction/fsd/common.hpp
    v3 v; f_t l;
    v3 nff, tff, nbf;
    f_t alpha;
    uint32_t ads_edge_idx;
    v3 e() const { return cross(nff, tff); }
    // utd.hpp:62-80
    std::optional<v3> diffraction_point(v3 src, v3 dst) const {
        const v3 ee = e();
        const f_t sl = length(v2{ dot(src - v, tff), dot(src - v, nff) });
        const f_t dl = length(v2{ dot(dst - v, tff), dot(dst - v, nff) });
        const f_t dist = dot(ee, src - v) + dot(dst - src, ee) * sl / (sl + dl);
        if (std::fabs(dist) > l / 2) return std::nullopt;
        const v3 p = v + ee * dist;
        if (p == src || p == dst) return std::nullopt;
        return p;
    }
    // utd.hpp:85-110
    std::optional<v3> diffraction_point_dir(v3 src, v3 wo) const {
        const v3 ee = e();
        const f_t cos_beta = dot(wo, ee);
        const f_t sin_beta = std::sqrt(std::max(0.f, 1 - sqr(cos_beta)));
        if (sin_beta < utd_min_sin_beta) return std::nullopt;
        const f_t sl = length(v2{ dot(src - v, tff), dot(src - v, nff) });
        const v3 prj_src = v + dot(src - v, ee) * ee;
        const v3 p = prj_src + sl * (cos_beta / sin_beta) * ee;
        if (length2(p - v) > sqr(l / 2)) return std::nullopt;
        if (p == src) return std::nullopt;
        return p;
    }
    struct UTD_ret_t { c_t Ds, Dh; };
    // utd.hpp:115-172
    UTD_ret_t UTD(f_t k, v3 wi, v3 wo, f_t ro) const {
        const v3 ee = e();
        const f_t n = 2 - alpha * inv_pi;
        const f_t sin_beta2 = std::max(0.f, 1 - sqr(dot(wi, ee)));
        const f_t sin_beta = std::sqrt(sin_beta2);
        const f_t phii = lm::atan2(dot(nff, wi), dot(tff, wi));
        const f_t phio = lm::atan2(dot(nff, wo), dot(tff, wo));
        const f_t Li = ro * sin_beta2;
        const f_t a1 = UTDa(+1, phii - phio, n), a2 = UTDa(-1, phii - phio, n);
        const f_t a3 = UTDa(+1, phii + phio, n), a4 = UTDa(-1, phii + phio, n);
        const f_t kL = k_times_len(k, Li);
        const c_t F1 = UTDF(kL * a1), F2 = UTDF(kL * a2), F3 = UTDF(kL * a3), F4 = UTDF(kL * a4);
        auto cot = [](f_t x) { return 1.f / lm::tan(x); };
        const c_t D1 = -cot((pi + (phii - phio)) / (2 * n)) * F1;
        const c_t D2 = -cot((pi - (phii - phio)) / (2 * n)) * F2;
        const c_t D3 = -cot((pi + (phii + phio)) / (2 * n)) * F3;
        const c_t D4 = -cot((pi - (phii + phio)) / (2 * n)) * F4;
        const f_t kro = k_times_len(k, ro);
        const c_t D = (1 / (2 * n * std::sqrt(kro) * sin_beta) * inv_sqrt_two_pi) * lm::expi(-pi_4);
        const f_t t1 = fmod_pos(phii + phio, pi_2);
        const f_t t2 = fmod_pos(phii - phio, pi_2);
        const bool z = std::fabs(t1) < 1e-5f || std::fabs(t2) < 1e-5f;
        const c_t Ds = z ? c_t{ 0, 0 } : D1 + D2 - (D3 + D4);
        const c_t Dh = z ? c_t{ 0, 0 } : D1 + D2 + (D3 + D4);
        return { -D * Ds, -D * Dh };
    }
};

struct fsd_t {      // free_space_diffraction_t
    std::vector<wedge_edge_t> edges;
    f_t k = 0;
    v3 interaction_wp{};
    bool empty() const { return edges.empty(); }

    // free_space_diffraction.cpp:23-82
    static fsd_t build(const scene_t& sc, v3 interaction_wp, const frame_t& region_frame, v3 region_size, v3 wi, f_t k, const std::vector<uint32_t>& edge_ids) {
        fsd_t f; f.k = k; f.interaction_wp = interaction_wp;
        for (uint32_t ed : edge_ids) {
            const wtgpu_edge& E = sc.d->edges[ed];
            const v3 n1{ E.n1[0], E.n1[1], E.n1[2] }, n2{ E.n2[0], E.n2[1], E.n2[2] };
            const v3 t1{ E.t1[0], E.t1[1], E.t1[2] }, t2{ E.t2[0], E.t2[1], E.t2[2] };
            const v3 ea{ E.a[0], E.a[1], E.a[2] }, eb{ E.b[0], E.b[1], E.b[2] };
            const bool f1_is_front = dot(wi, n1) > 0;
            const v3 nff = f1_is_front ? n1 : n2, tff = f1_is_front ? t1 : t2, nbf = f1_is_front ? n2 : n1;
            if (dot(wi, nff) <= 0) continue;
            v3 v1 = ea, v2p = eb;
            if (isfinite3(region_size)) {
                const auto intr = intersect_edge_ellipsoid(ea, eb, interaction_wp, region_frame.t, region_frame.b, region_size);
                const f_t tt1 = clampf(intr.t1, 0, 1), tt2 = clampf(intr.t2, 0, 1);
                v1 = { mix(ea.x, eb.x, tt1), mix(ea.y, eb.y, tt1), mix(ea.z, eb.z, tt1) };
                v2p = { mix(ea.x, eb.x, tt2), mix(ea.y, eb.y, tt2), mix(ea.z, eb.z, tt2) };
            }
            if (v1 == v2p) continue;
            const v3 v = (v1 + v2p) / 2.f;
            const f_t l = length(v2p - v1);
            f.edges.push_back({ v, l, nff, tff, nbf, E.alpha, ed });
        }
        return f;
    }

    struct diffracting_edge_t { wedge_edge_t::UTD_ret_t utd; uint32_t edge_idx; v3 p; f_t ri, ro; };
    // free_space_diffraction.cpp:196-234
    std::vector<diffracting_edge_t> f(v3 src, v3 dst) const {
        std::vector<diffracting_edge_t> ret;
        for (const auto& e : edges) {
            const auto p = e.diffraction_point(src, dst);
            if (!p) continue;
            const v3 ui = src - *p, uo = dst - *p;
            if ((dot(uo, e.nff) <= 0 && dot(uo, e.nbf) <= 0) || (dot(ui, e.nff) <= 0 && dot(ui, e.nbf) <= 0)) continue;
            const f_t ri = length(ui), ro = length(uo);
            const v3 wi = ui / ri, wo = uo / ro;
            const auto u = e.UTD(k, wi, wo, ro);
            if (u.Dh == c_t{ 0, 0 } && u.Ds == c_t{ 0, 0 }) continue;
            ret.push_back({ u, e.ads_edge_idx, *p, ri, ro });
        }
        return ret;
    }
    // free_space_diffraction.cpp:152-194
    f_t pdf(v3 src, v3 wo) const {
        if (edges.empty()) return 0;
        f_t ret = 0;
        for (const auto& edge : edges) {
            const auto p = edge.diffraction_point_dir(src, wo);
            if (!p) continue;
            const v3 ui = src - *p;
            if ((dot(wo, edge.nff) <= 0 && dot(wo, edge.nbf) <= 0) || (dot(ui, edge.nff) <= 0 && dot(ui, edge.nbf) <= 0)) continue;
            const f_t ri = length(src - *p);
            const v3 wi = (src - *p) / ri;
            const f_t phii = lm::atan2(dot(edge.nff, wi), dot(edge.tff, wi));
            const f_t phio = lm::atan2(dot(edge.nff, wo), dot(edge.tff, wo));
            const f_t sigma = std::sqrt(utd_IS_sigma_scale / k_times_len(k, ri));
            const f_t mean_phi1 = pi + phii, mean_phi2 = pi - phii;
            f_t x1 = std::fabs(fmod_pos(phio - mean_phi1, two_pi));
            f_t x2 = std::fabs(fmod_pos(phio - mean_phi2, two_pi));
            if (x1 > pi) x1 -= two_pi;
            if (x2 > pi) x2 -= two_pi;
            const f_t apd = inv_sqrt_two_pi / sigma * (lm::exp(-.5f * sqr(x1 / sigma)) + lm::exp(-.5f * sqr(x2 / sigma))) / 2;
            ret += apd;
        }
        return ret / (f_t)(edges.size() + 1);
    }
    struct sample_ret_t { v3 wo{ 0, 0, 1 }; f_t weight = 0; bool valid = false; };
    // free_space_diffraction.cpp:84-150
    sample_ret_t sample(v3 src, sampler_t& sampler) const {
        const int eidx = sampler.uniform_int_interval(0, (int)edges.size() + 1);
        if (eidx == (int)edges.size()) {
            const v3 wi = normalize(src - interaction_wp);
            return { -wi, (f_t)(edges.size() + 1), true };
        }
        const auto& edge = edges[eidx];
        const v3 p = edge.v + (sampler.r() - .5f) * edge.l * edge.e();
        const v3 ui = src - p;
        if (dot(ui, edge.nff) <= 0 && dot(ui, edge.nbf) <= 0) return {};
        const f_t ri = length(src - p);
        const v3 wi = (src - p) / ri;
        const f_t phii = lm::atan2(dot(edge.nff, wi), dot(edge.tff, wi));
        const f_t sigma = std::sqrt(utd_IS_sigma_scale / k_times_len(k, ri));
        const f_t smp = sigma * normal2d(sampler.r2()).x;
        const f_t mean_phi1 = pi + phii, mean_phi2 = pi - phii;
        const f_t phio = (sampler.r() < .5f ? mean_phi1 : mean_phi2) + smp;
        const v3 e = edge.e();
        const f_t cos_beta = dot(wi, e);
        const f_t sin_beta = std::sqrt(std::max(0.f, 1 - sqr(cos_beta)));
        const v3 wo = sin_beta * (lm::cos(phio) * edge.tff + lm::sin(phio) * edge.nff) - cos_beta * e;
        if (dot(wo, edge.nff) <= 0 && dot(wo, edge.nbf) <= 0) return {};
        if (sin_beta < utd_min_sin_beta) return {};
        const f_t dpd = pdf(src, wo);
        if (dpd == 0) return {};
        return { wo, 1.f / dpd, true };
    }
};

// ================================================================================================
// plt_path (plt_path_detail.hpp)
// ================================================================================================
struct path_stats_t { uint64_t segments = 0, surface = 0, fsd = 0, null = 0, splats = 0; ads_counters_t ads; };

struct plt_path_t {
    const scene_t& sc;
    bsdf_eval_t bsdfs;
    emitters_t emitters;
    sensor_eval_t sensor;
    film_t& film;
    path_stats_t* stats;
    uint32_t max_depth; bool RR, FSD;
    bool force_rt;

    plt_path_t(const scene_t& s, film_t& f, path_stats_t* st) : sc(s), bsdfs(s), emitters(s), sensor(s), film(f), stats(st),
        max_depth(s.d->integrator.max_depth), RR(s.d->integrator.russian_roulette != 0), FSD(s.d->integrator.fsd != 0),
        force_rt(s.d->sensor.ray_trace_only != 0) {}

    struct walk_t {         // path_walk_data_t (plt_path_detail.hpp:33-143)
        beam_t beam;
        geo_t prev_vert_geo;
        std::optional<beam_t> prev_vert_beam;
        bool sampled_fsd = false;
        pd_t from_previous_dpd = pd_t::discrete(0);
        std::optional<fsd_t> fsd_bsdf;
        f_t throughput = 1;
        sampler_t* sampler = nullptr;
    };

    ads_counters_t* ctr() const { return stats ? &stats->ads : nullptr; }

    static f_t MIS(f_t pd1, f_t pd2) { if (pd2 == 0) return 1; return pd1 * pd1 / (pd1 * pd1 + pd2 * pd2); }     // :303-308

    bool continue_walk(walk_t& d, uint32_t depth, bool allow_RR) const {      // :123-142
        if (depth >= max_depth) return false;
        if (d.beam.intensity() == 0) return false;
        if (!allow_RR || !RR) return true;
        const f_t r = d.throughput < 1 ? std::max(d.throughput, .5f) : 1.f;
        if (d.sampler->r() <= r) { const f_t scale = 1 / r; d.beam.mul(scale); d.throughput *= scale; return true; }
        return false;
    }

    // :311-346
    std::pair<c_t, c_t> do_fsd(const elliptic_cone_t& cone_from_src, const geo_t& src_geo, v3 dst, const fsd_t& fsd_bsdf, f_t k) const {
        const v3 src = cone_from_src.o();
        const geo_t dst_geo = geo_t::point(dst);
        c_t ts{}, th{};
        for (const auto& f : fsd_bsdf.f(src, dst)) {
            const geo_t eintr = geo_t::on_edge(f.edge_idx, f.p);
            if (shadow(sc, eintr, src_geo, ctr()) || shadow(sc, eintr, dst_geo, ctr())) continue;
            const f_t dd = f.ro + f.ri;
            const c_t phase = lm::expi(-k_times_len(k, dd));
            ts += phase * f.utd.Ds; th += phase * f.utd.Dh;
        }
        if (cone_from_src.contains(dst)) {
            if (!shadow(sc, src_geo, dst_geo, ctr())) {
                const f_t dd = length(dst - src);
                const c_t phase = lm::expi(-k_times_len(k, dd));
                ts += phase; th += phase;
            }
        }
        return { ts, th };
    }

    int32_t bsdf_of(const surface_t& s) const { return sc.d->shapes[sc.d->tri_meta[s.tuid].shape_idx].bsdf; }
    int32_t emitter_of(const surface_t& s) const { return sc.d->shapes[sc.d->tri_meta[s.tuid].shape_idx].emitter; }

    // :156-203
    bool sample_surface_interaction(walk_t& d, const surface_t& intersection) const {
        const f_t k = d.beam.k;
        const int32_t bsdf = bsdf_of(intersection);
        const bsdf_query_t q{ &intersection, k, d.beam.forward };
        const v3 ng = intersection.ng();
        const v3 wiworld = -d.beam.dir();
        const v3 wi = intersection.shading.to_local(wiworld);
        const f_t wig = dot(wiworld, ng), wis = wi.z;
        if (wig * wis <= 0) return false;
        const auto smp = bsdfs.sample(bsdf, wi, q, *d.sampler);
        if (!smp || smp->dpd.is_zero()) return false;
        const v3 wo = smp->wo;
        const v3 woworld = normalize(intersection.shading.to_world(wo));
        const f_t wog = dot(woworld, ng), wos = wo.z;
        if (stats) stats->surface++;
        if (wog * wos <= 0) return false;
        // path_walk_data_t::transform_surface_interaction (:64-82)
        d.from_previous_dpd = smp->dpd;
        d.prev_vert_geo = geo_t::surface(intersection);
        d.prev_vert_beam = d.beam;
        d.sampled_fsd = false;
        d.beam.transform_surface_interaction(intersection, woworld, smp->M, 1);
        d.throughput *= 1 * smp->M.mean_intensity();
        const f_t eta = smp->eta.real();
        if (eta != 1) d.throughput /= sqr(eta);
        return true;
    }

    // :253-276
    struct wf_intersection_t { uint32_t primary = WTGPU_INVALID_IDX; f_t dist = inf; v2 bary{ -1, -1 }; std::optional<surface_t> intersection; };
    wf_intersection_t find_closest_triangle(const std::vector<uint32_t>& tris, range_t zr, v3 origin, v3 beam_dir) const {
        wf_intersection_t id;
        for (uint32_t tuid : tris) {
            const v3 a = sc.ads.tri_a(tuid), b = sc.ads.tri_b(tuid), c = sc.ads.tri_c(tuid);
            const f_t fptol = cone_intersection_tolerance(origin, aabb_t::from_points(a, b, c));
            const auto intr = intersect_ray_tri(ray_t{ origin, beam_dir }, a, b, c, zr.grow(fptol));
            if (intr && intr->dist < id.dist) { id.primary = tuid; id.dist = intr->dist; id.bary = intr->bary; }
        }
        return id;
    }

    // :350-424
    stokes_t nee_backward(walk_t& d, const wf_intersection_t& wf) const {
        const beam_t& beam = d.beam;
        const f_t k = beam.k;
        if (!wf.intersection) return {};
        const surface_t& intersection = *wf.intersection;
        const int32_t bsdf = bsdf_of(intersection);
        if (bsdfs.is_delta_only(bsdf, k)) return {};
        const auto ds = emitters.sample_emitter_direct(*d.sampler, intersection.wp, k);
        const f_t sampled_emitter_pm = ds.emitter_pdf;
        if (ds.beam.intensity() == 0) return {};
        const v3 wiworld = -beam.dir(), woworld = -ds.beam.dir();
        const v3 ng = intersection.ng();
        const v3 wi = intersection.shading.to_local(wiworld), wo = intersection.shading.to_local(woworld);
        const f_t wig = dot(wiworld, ng), wog = dot(woworld, ng);
        if (wi.z * wig <= 0 || wo.z * wog <= 0) return {};
        const bsdf_query_t q{ &intersection, k, beam.forward };
        const mueller_t f = bsdfs.f(bsdf, wi, wo, q);
        if (f.mean_intensity() == 0) return {};
        const geo_t emitter_geo = ds.surface ? geo_t::surface(*ds.surface) : geo_t::point(ds.beam.origin());
        if (shadow(sc, geo_t::surface(intersection), emitter_geo, ctr())) return {};
        beam_t nee_beam = beam;
        nee_beam.transform_surface_interaction(intersection, woworld, f, 1);
        const stokes_t sL = integrate_beams(nee_beam, ds.beam);
        f_t mis = 1;
        if (!ds.dpd.is_discrete) {
            const f_t pd_brdf = bsdfs.pdf(bsdf, wi, wo, q);
            const f_t pd_direct = ds.dpd.v * sampled_emitter_pm;
            mis = MIS(pd_direct, pd_brdf);
        }
        return sL * mis;
    }
    // :427-465
    stokes_t emission(walk_t& d, const surface_t& intersection) const {
        const int32_t em = emitter_of(intersection);
        if (em < 0) return {};
        const beam_t& beam = d.beam;
        const stokes_t sL = emitters.Li(em, beam, &intersection);
        f_t mis = 1;
        if (!d.from_previous_dpd.is_discrete) {
            const f_t emitter_pm = emitters.pdf_emitter(em);
            const f_t emitter_ppd = emitters.pdf_position_density(em);
            const f_t dn = dot(-beam.dir(), intersection.ng());
            const f_t recp_dn = dn != 0 ? 1 / std::fabs(dn) : 0.f;
            const f_t l2 = length2(beam.origin() - intersection.wp);
            const f_t pd_nee = emitter_ppd * l2 * recp_dn;
            const f_t pd_brdf = d.from_previous_dpd.v;
            const f_t pd_direct = pd_nee * emitter_pm;
            mis = MIS(pd_brdf, pd_direct);
        }
        return sL * mis;
    }
    // :468-510
    void nee_forward(walk_t& d, v3 interaction_wp, f_t beam_dist, f_t recp_spectral_pd) const {
        const beam_t& beam = d.beam;
        const f_t k = beam.k;
        if (!d.fsd_bsdf) return;
        if (!sensor.is_virtual()) return;
        auto sd = sensor.sample_direct(*d.sampler, interaction_wp, k);
        if ((sd.dpd.is_discrete || !sd.dpd.is_zero()) && sd.beam.intensity() > 0) {
            const auto fsd = do_fsd(beam.envelope, d.prev_vert_geo, sd.beam.origin(), *d.fsd_bsdf, k);
            const f_t f = (std::norm(fsd.first) + std::norm(fsd.second)) / 2;
            if (f == 0) return;
            beam_t fsd_beam = beam;
            fsd_beam.transform_region_interaction(interaction_wp, beam_dist, -sd.beam.dir(), f);
            const stokes_t sL = integrate_beams(sd.beam, fsd_beam);
            film.splat_direct(sd.element, sL * recp_spectral_pd, k);
            if (stats) stats->splats++;
        }
    }
    // :513-540
    void sensing(walk_t& d, v3 origin_wp, f_t beam_propagation_distance, f_t recp_spectral_pd) const {
        const beam_t& beam = d.beam;
        if (!sensor.is_virtual()) return;
        const f_t max_distance = beam_propagation_distance - std::max(0.f, dot(beam.dir(), origin_wp - beam.origin()));
        const auto dc = sensor.Si(beam, { 0, max_distance });
        if (dc) {
            const stokes_t sL = integrate_beams(dc->beam, beam);
            film.splat_direct(dc->element, sL * recp_spectral_pd, beam.k);
            if (stats) stats->splats++;
        }
    }

    // :542-762 (tail recursion unrolled into a loop)
    stokes_t random_walk(walk_t& data, f_t recp_spectral_pd) const {
        stokes_t L{};
        uint32_t depth = 1;
        for (;;) {
            beam_t& beam = data.beam;
            if (stats) stats->segments++;
            const auto intersection = traverse(sc, beam.envelope, data.prev_vert_geo, wavenum_to_wavelen(beam.k), force_rt, FSD, ctr());
            if (intersection.empty) return L;
            const f_t dist_to_interaction = intersection.distance();
            const bool is_ballistic = intersection.ballistic || beam.is_ray();
            const frame_t beam_frame = beam.envelope.frame();
            const elliptic_cone_t envelope = beam.envelope;
            std::vector<uint32_t> ray_tris;
            if (intersection.ballistic) ray_tris.push_back(intersection.ray.tuid);
            const std::vector<uint32_t>& tris = intersection.ballistic ? ray_tris : intersection.cone.tris;
            std::vector<uint32_t> edges = intersection.ballistic ? std::vector<uint32_t>{} : intersection.cone.edges;
            const v3 origin_wp = intersection.origin;
            const v3 interaction_wp = origin_wp + dist_to_interaction * beam.dir();

            // evaluate fsd from previous interaction (:591-610)
            if (data.fsd_bsdf) {
                const elliptic_cone_t prev_cone = data.prev_vert_beam->envelope;
                const auto fsd = do_fsd(prev_cone, data.prev_vert_geo, interaction_wp, *data.fsd_bsdf, beam.k);
                data.fsd_bsdf.reset();
                const f_t f = (std::norm(fsd.first) + std::norm(fsd.second)) / 2;
                if (data.sampled_fsd) data.beam.mul(f);
                else {
                    data.prev_vert_beam->transform_region_interaction(origin_wp, length(origin_wp - prev_cone.o()), beam.dir(), f);
                    data.beam.add(*data.prev_vert_beam);
                }
            }

            // primary triangle (:616-652)
            wf_intersection_t wf;
            if (is_ballistic) { wf.primary = intersection.ray.tuid; wf.dist = intersection.ray.dist; wf.bary = intersection.ray.bary; }
            else wf = find_closest_triangle(tris, { dist_to_interaction, dist_to_interaction + intersection.intersection_region_depth }, origin_wp, beam.dir());
            const f_t interaction_region_end = wf.primary != WTGPU_INVALID_IDX ? wf.dist : dist_to_interaction;
            if (wf.primary != WTGPU_INVALID_IDX) {
                const v3 sampled_tri_wp = origin_wp + wf.dist * beam.dir();
                surface_t s = sc.make_surface(wf.primary, wf.bary, sampled_tri_wp);
                s.footprint = data.beam.surface_footprint_static(s, dist_to_interaction);
                wf.intersection = s;
            }

            // ballistic: find edges around the intersection (:656-660)
            if (is_ballistic && !beam.is_ray() && !force_rt) {
                const f_t zdist = envelope.axes(dist_to_interaction).x * major_axis_to_z_scale;
                const auto eintr = intersect_cone(sc.ads, envelope, { dist_to_interaction - zdist / 2, dist_to_interaction + zdist / 2 }, 1, true, ctr());
                edges = eintr.edges;
            }
            // construct fsd BSDF (:663-679)
            if (!edges.empty()) {
                const v3 footprint = beam.footprint(dist_to_interaction);
                fsd_t f = fsd_t::build(sc, interaction_wp, beam_frame, footprint, -data.beam.dir(), beam.k, edges);
                if (!f.empty()) data.fsd_bsdf = std::move(f);
                if (stats) stats->fsd++;
            }

            // NEE (:691-705)
            if (!beam.forward) { if (depth < max_depth) L = L + nee_backward(data, wf); }
            else if (depth < max_depth) nee_forward(data, interaction_wp, dist_to_interaction, recp_spectral_pd);

            // organic connections (:711-723)
            if (!beam.forward) { if (wf.intersection) L = L + emission(data, *wf.intersection); }
            else sensing(data, origin_wp, interaction_region_end, recp_spectral_pd);

            // interactions (:729-749)
            bool sampled_null = false;
            if (wf.intersection) {
                if (!sample_surface_interaction(data, *wf.intersection)) return L;
            } else if (data.fsd_bsdf) {
                // sample_fsd_interaction (:221-237)
                const v3 prev_wp = data.prev_vert_geo.position();
                const auto smp = data.fsd_bsdf->sample(prev_wp, *data.sampler);
                // transform_fsd_interaction (:101-116)
                data.from_previous_dpd = pd_t::discrete(0);
                data.prev_vert_geo = geo_t::point(interaction_wp);
                data.prev_vert_beam = data.beam;
                data.sampled_fsd = true;
                data.beam.transform_region_interaction(interaction_wp, dist_to_interaction, smp.wo, smp.weight);
                data.throughput *= smp.weight;
            } else {
                sampled_null = true;
                data.beam.transform_restart(interaction_wp, dist_to_interaction);   // sample_null_interaction (:207-217)
                if (stats) stats->null++;
            }

            if (!continue_walk(data, depth, !sampled_null)) return L;
            if (!sampled_null) depth++;
        }
    }

    // recp_spectral_pd (:778-780 / :815): 1/mass for a discrete wavenumber sample, else 1/sum of densities
    // :764-801
    void integrate_backward(uint32_t ex, uint32_t ey, sampler_t& sampler) const {
        if (max_depth == 0) return;
        const int32_t em = emitters.sample_emitter(sampler);
        const auto ws = emitters.sample_wavenumber(em, sampler);
        const f_t k = ws.k;
        const f_t recp_spectral_pd = ws.wpd.is_discrete ? 1.f / ws.wpd.v : 1.f / emitters.sum_spectral_pdf_for_all_emitters(k);
        const auto ss = sensor.sample(sampler, ex, ey, k);
        sampler.end_scene_draws();      // scene->sampler() is used up to here (plt_path_detail.hpp:772,783); the walk has its own uniform sampler (:59-60,146)
        walk_t data; data.beam = ss.beam; data.prev_vert_geo = geo_t::point(ss.beam.origin()); data.sampler = &sampler;
        const stokes_t L = random_walk(data, recp_spectral_pd);
        film.splat(ss.element, L * recp_spectral_pd, k);
        if (stats) stats->splats++;
    }
    // :804-828
    void integrate_forward(sampler_t& sampler) const {
        if (max_depth == 0) return;
        const int32_t em = emitters.sample_emitter(sampler);
        const auto ws = emitters.sample_wavenumber(em, sampler);
        const f_t k = ws.k;
        const auto es = emitters.sample(em, sampler, k);
        const f_t recp_spectral_pd = 1.f / emitters.sum_spectral_pdf_for_all_emitters(k);
        sampler.end_scene_draws();
        walk_t data; data.beam = es.beam; data.prev_vert_geo = geo_t::point(es.beam.origin()); data.sampler = &sampler;
        random_walk(data, recp_spectral_pd);
    }
    // plt_path_t::integrate (src/integrator/plt_path.cpp:40-51), one sample
    void integrate(uint32_t ex, uint32_t ey, sampler_t& sampler) const {
        if (sc.d->integrator.direction == WTGPU_DIRECTION_BACKWARD) integrate_backward(ex, ey, sampler);
        else integrate_forward(sampler);
    }
};

} // namespace ot
