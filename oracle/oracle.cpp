// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// oracle.cpp: C entry points of the CPU oracle (ctypes-loaded by tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py).  PARITY UNPINNED by the reference (no golden vectors, reference
// not compilable here -- SURVEY.md 8c); pinned by analytic KATs only.
#include <cstdlib>
#include <cstdio>
#include "ot_bdpt.h"
#include "oracle.h"
#include <thread>
#include <atomic>
#include <chrono>

using namespace ot;

static bool sc_has_lut(const wtgpu_scene_desc* d) { return d->fsd_lut_n > 1 && d->fsd_lut_m > 1 && d->fsd_icdf_theta1 && d->fsd_icdf1 && d->fsd_icdf_theta2 && d->fsd_icdf2; }

extern "C" {

int oracle_render(const wtgpu_scene_desc* desc, const wtgpu_render_opts* opts, double* film_block, double* film_light,
                  uint32_t n_threads, oracle_stats* st) {
    if (!desc || !opts) return -1;
    const bool bdpt = desc->integrator.type == WTGPU_INTEGRATOR_PLT_BDPT;
    if (bdpt && !sc_has_lut(desc) && desc->integrator.fsd && !desc->sensor.ray_trace_only) return -4;
    std::unique_ptr<sobol_ctx_t> sob;
    if (opts->sampler == WTGPU_SAMPLER_SOBOLLD) { if (!desc->sobol_table || opts->spp == 0) return -1; sob.reset(new sobol_ctx_t(desc->sobol_table)); }
    scene_t sc(desc);
    const uint32_t W = desc->sensor.width, H = desc->sensor.height, C = desc->sensor.channels;
    const uint32_t x0 = opts->tile_x0, y0 = opts->tile_y0, x1 = std::min(opts->tile_x1, W), y1 = std::min(opts->tile_y1, H);
    if (x1 <= x0 || y1 <= y0) return 0;
    const uint64_t tw = x1 - x0, npix = tw * (y1 - y0);
    if (n_threads == 0) n_threads = std::max(1u, std::thread::hardware_concurrency());
    n_threads = (uint32_t)std::min<uint64_t>(n_threads, npix);

    std::vector<film_t> films; films.reserve(n_threads);
    for (uint32_t t = 0; t < n_threads; ++t) films.emplace_back(sc);
    std::vector<path_stats_t> stats(n_threads);
    std::atomic<uint64_t> next{ 0 };
    const uint64_t chunk = 64;
    const auto t_start = std::chrono::steady_clock::now();
    std::vector<bdpt_stats_t> bstats(n_threads);
    auto worker = [&](uint32_t tid) {
        plt_path_t integ(sc, films[tid], &stats[tid]);
        plt_bdpt_t binteg(sc, films[tid], &bstats[tid]);
        for (;;) {
            const uint64_t b = next.fetch_add(chunk);
            if (b >= npix) break;
            for (uint64_t i = b; i < std::min(npix, b + chunk); ++i) {
                const uint32_t ex = x0 + (uint32_t)(i % tw), ey = y0 + (uint32_t)(i / tw);
                for (uint32_t s = opts->sample_begin; s < opts->sample_end; ++s) {
                    sampler_t smp; smp.seed = opts->seed; smp.pixel = ey * W + ex; smp.sample = s;
                    smp.begin_scene_draws(sob.get(), opts->spp);
                    if (bdpt) binteg.integrate(ex, ey, smp); else integ.integrate(ex, ey, smp);
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (uint32_t t = 1; t < n_threads; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto& t : th) t.join();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();

    for (uint32_t t = 1; t < n_threads; ++t) films[0].merge(films[t]);
    if (film_block) for (size_t i = 0; i < (size_t)W * H * C * 2; ++i) film_block[i] += films[0].block[i];
    if (film_light) for (size_t i = 0; i < (size_t)W * H * C; ++i) film_light[i] += films[0].light[i];
    if (st) {
        memset(st, 0, sizeof(*st));
        st->samples = npix * (opts->sample_end - opts->sample_begin);
        st->seconds = secs; st->threads = n_threads;
        for (auto& s : bstats) { st->segments += s.vertices; st->fsd += s.connections; st->splats += s.splats;
            st->nodes += s.ads.nodes; st->tris += s.ads.tris; st->ray_casts += s.ads.ray_casts; st->cone_casts += s.ads.cone_casts; st->shadow_casts += s.ads.shadow_casts; }
        for (auto& s : stats) {
            st->segments += s.segments; st->surface += s.surface; st->fsd += s.fsd; st->null_ += s.null; st->splats += s.splats;
            st->nodes += s.ads.nodes; st->tris += s.ads.tris; st->ray_casts += s.ads.ray_casts; st->cone_casts += s.ads.cone_casts; st->shadow_casts += s.ads.shadow_casts;
        }
    }
    return 0;
}

// sobolld: the literal generate_points() of batch `batch` under our seeding contract; out arrays hold n_points*47 values (point-major)
int oracle_sobol_batch(const wtgpu_sobol_entry* table, uint64_t seed, uint64_t batch, uint32_t n_points, uint32_t* out_numerators, float* out_values) {
    if (!table) return -1;
    sobol::gf3_t gf3(table);
    sobol::sobolls_sampler gen(sobol::N, gf3);
    uint64_t seeds[sobol::D]; sobol_ctx_t::seeds_for_batch(seed, batch, seeds);
    std::vector<float> v; std::vector<uint32_t> num;
    gen.generate_points(seeds, n_points, v, &num);
    for (size_t i = 0; i < v.size(); ++i) { if (out_values) out_values[i] = v[i]; if (out_numerators) out_numerators[i] = num[i]; }
    return (int)(v.size() / sobol::D);
}
// the same with explicit per-dimension seeds (what the reference draws from its RNG), and the seeds our contract derives for a batch:
// the hooks that let tests/test_sobol.py put the restatement next to the reference's own code (oracle/_ref/libref_sobol.so)
int oracle_sobol_points_with_seeds(const wtgpu_sobol_entry* table, const uint64_t* seeds, uint32_t n_points, uint32_t* out_numerators, float* out_values) {
    if (!table || !seeds) return -1;
    sobol::gf3_t gf3(table);
    sobol::sobolls_sampler gen(sobol::N, gf3);
    uint64_t sd[sobol::D]; for (size_t i = 0; i < sobol::D; ++i) sd[i] = seeds[i];
    std::vector<float> v; std::vector<uint32_t> num;
    gen.generate_points(sd, n_points, v, &num);
    for (size_t i = 0; i < v.size(); ++i) { if (out_values) out_values[i] = v[i]; if (out_numerators) out_numerators[i] = num[i]; }
    return (int)(v.size() / sobol::D);
}
void oracle_sobol_seeds(uint64_t seed, uint64_t batch, uint64_t* out) { sobol_ctx_t::seeds_for_batch(seed, batch, out); }
// generator matrices as gen_mat builds them: out[dim][row][col], 47 x 11 x 11 digits
int oracle_sobol_matrices(const wtgpu_sobol_entry* table, int32_t* out) {
    if (!table) return -1;
    sobol::gf3_t gf3(table);
    sobol::sobolls_sampler gen(sobol::N, gf3);
    for (size_t d = 0; d < sobol::D; ++d) for (size_t r = 0; r < sobol::N; ++r) for (size_t c = 0; c < sobol::N; ++c) out[(d * sobol::N + r) * sobol::N + c] = (int32_t)gen.matrix[d][r][c];
    return 0;
}

int oracle_intersect_rays(const wtgpu_scene_desc* desc, uint32_t n, const wtgpu_ray_query* q, wtgpu_ray_hit* out) {
    ads_t ads(desc);
    for (uint32_t i = 0; i < n; ++i) {
        const ray_t r{ { q[i].o[0], q[i].o[1], q[i].o[2] }, { q[i].d[0], q[i].d[1], q[i].d[2] } };
        const auto h = intersect_ray(ads, r, { q[i].tmin, q[i].tmax });
        out[i].tuid = h.tuid; out[i].dist = h.dist; out[i].bary[0] = h.bary.x; out[i].bary[1] = h.bary.y; out[i].front_face = h.front_face;
    }
    return 0;
}
int oracle_shadow_rays(const wtgpu_scene_desc* desc, uint32_t n, const wtgpu_ray_query* q, uint32_t* out) {
    ads_t ads(desc);
    for (uint32_t i = 0; i < n; ++i) {
        const ray_t r{ { q[i].o[0], q[i].o[1], q[i].o[2] }, { q[i].d[0], q[i].d[1], q[i].d[2] } };
        out[i] = shadow_ray(ads, r, { q[i].tmin, q[i].tmax }) ? 1u : 0u;
    }
    return 0;
}
// brute force over all triangles with the scalar Moeller-Trumbore: pins the BVH traversal itself
int oracle_intersect_rays_bruteforce(const wtgpu_scene_desc* desc, uint32_t n, const wtgpu_ray_query* q, wtgpu_ray_hit* out) {
    ads_t ads(desc);
    for (uint32_t i = 0; i < n; ++i) {
        const ray_t r{ { q[i].o[0], q[i].o[1], q[i].o[2] }, { q[i].d[0], q[i].d[1], q[i].d[2] } };
        ray_hit_t rec;
        for (uint32_t t = 0; t < desc->n_tris; ++t) {
            const auto w = intersect_ray_tri_w(r.o, r.d, ads.tri_a(t), ads.tri_b(t), ads.tri_c(t), { q[i].tmin, q[i].tmax });
            if (w.result != -inf && w.result < rec.dist) { rec.dist = w.result; rec.tuid = t; rec.bary = { w.baryx, w.baryy }; rec.front_face = dot(ads.tri_n(t), r.d) <= 0; }
        }
        out[i].tuid = rec.tuid; out[i].dist = rec.dist; out[i].bary[0] = rec.bary.x; out[i].bary[1] = rec.bary.y; out[i].front_face = rec.front_face;
    }
    return 0;
}
static elliptic_cone_t cone_of(const wtgpu_cone_query& q) {
    const ray_t r{ { q.o[0], q.o[1], q.o[2] }, { q.d[0], q.d[1], q.d[2] } };
    return elliptic_cone_t::make(r, { q.x[0], q.x[1], q.x[2] }, q.x0, q.tan_alpha, 1.f / q.e, q.e);
}
int oracle_intersect_cones(const wtgpu_scene_desc* desc, uint32_t n, const wtgpu_cone_query* q, wtgpu_cone_hit* out) {
    ads_t ads(desc);
    for (uint32_t i = 0; i < n; ++i) {
        const auto rec = intersect_cone(ads, cone_of(q[i]), { q[i].tmin, q[i].tmax }, q[i].z_scale, true);
        wtgpu_cone_hit& h = out[i];
        memset(&h, 0, sizeof(h));
        h.dist = rec.empty() ? inf : rec.dist; h.front_face = rec.front_face;
        h.n_tris = (uint32_t)rec.tris.size(); h.n_edges = (uint32_t)rec.edges.size();
        for (uint32_t j = 0; j < std::min<uint32_t>(h.n_tris, WTGPU_MAX_CONE_TRIS); ++j) h.tris[j] = rec.tris[j];
        for (uint32_t j = 0; j < std::min<uint32_t>(h.n_edges, WTGPU_MAX_CONE_EDGES); ++j) h.edges[j] = rec.edges[j];
    }
    return 0;
}
// brute-force cone query: distance of the closest triangle over ALL triangles (no BVH, no range shrinking)
int oracle_cone_closest_bruteforce(const wtgpu_scene_desc* desc, uint32_t n, const wtgpu_cone_query* q, float* dist) {
    ads_t ads(desc);
    for (uint32_t i = 0; i < n; ++i) {
        const auto cone = cone_of(q[i]);
        float best = inf;
        for (uint32_t t = 0; t < desc->n_tris; ++t) {
            const auto r = intersect_cone_tri(cone, ads.tri_a(t), ads.tri_b(t), ads.tri_c(t), ads.tri_n(t), { q[i].tmin, q[i].tmax });
            if (r && r->dist < best) best = r->dist;
        }
        dist[i] = best;
    }
    return 0;
}
int oracle_rng(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, float* out) {
    sampler_t s; s.seed = seed; s.pixel = pixel; s.sample = sample;
    for (uint32_t i = 0; i < n; ++i) out[i] = s.r();
    return 0;
}

// ---- analytic known-answer hooks
void oracle_svd(const float A[4], float out[6]) {
    const auto s = SVD(mat2{ A[0], A[1], A[2], A[3] });
    out[0] = s.Ucos; out[1] = s.Usin; out[2] = s.Vcos; out[3] = s.Vsin; out[4] = s.sigma1; out[5] = s.sigma2;
}
void oracle_svd_n(uint32_t n, const float* A, float* out) { for (uint32_t i = 0; i < n; ++i) oracle_svd(A + 4 * i, out + 6 * i); }
// wedge_edge_t::UTD and ::diffraction_point for n wedges -- same layouts as oracle/ref_utd.cpp (wedge: v[3] l nff[3] tff[3] nbf[3] alpha; q: k wi[3] wo[3] ro)
static wedge_edge_t wedge_from(const float* w) {
    wedge_edge_t e{}; e.v = { w[0], w[1], w[2] }; e.l = w[3]; e.nff = { w[4], w[5], w[6] }; e.tff = { w[7], w[8], w[9] }; e.nbf = { w[10], w[11], w[12] }; e.alpha = w[13];
    return e;
}
void oracle_utd(uint32_t n, const float* wedge, const float* q, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = q + 8 * i; float* o = out + 4 * i;
        const auto r = wedge_from(wedge + 14 * i).UTD(a[0], v3{ a[1], a[2], a[3] }, v3{ a[4], a[5], a[6] }, a[7]);
        o[0] = r.Ds.real(); o[1] = r.Ds.imag(); o[2] = r.Dh.real(); o[3] = r.Dh.imag();
    }
}
void oracle_utd_diffraction_points(uint32_t n, const float* wedge, const float* pts, int* found, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = pts + 9 * i; float* o = out + 6 * i;
        const auto e = wedge_from(wedge + 14 * i);
        const auto p = e.diffraction_point(v3{ a[0], a[1], a[2] }, v3{ a[3], a[4], a[5] });
        const auto d = e.diffraction_point_dir(v3{ a[0], a[1], a[2] }, v3{ a[6], a[7], a[8] });
        found[2 * i] = p ? 1 : 0; found[2 * i + 1] = d ? 1 : 0;
        for (int k = 0; k < 6; ++k) o[k] = 0;
        if (p) { o[0] = p->x; o[1] = p->y; o[2] = p->z; }
        if (d) { o[3] = d->x; o[4] = d->y; o[5] = d->z; }
    }
}
void oracle_utdf_n(uint32_t n, const float* x, float* out) { for (uint32_t i = 0; i < n; ++i) { const c_t f = UTDF(x[i]); out[2 * i] = f.real(); out[2 * i + 1] = f.imag(); } }
void oracle_utdf(float x, float out[2]) { const c_t f = UTDF(x); out[0] = f.real(); out[1] = f.imag(); }
void oracle_cerfc_rot45(double s, double out[2]) { const auto c = cerfc_rot45(s); out[0] = c.real(); out[1] = c.imag(); }
void oracle_fresnel(float eta_re, float eta_im, const float w[3], float out[12]) {
    const auto f = fresnel(c_t{ eta_re, eta_im }, v3{ w[0], w[1], w[2] });
    out[0] = f.rs.real(); out[1] = f.rs.imag(); out[2] = f.rp.real(); out[3] = f.rp.imag();
    out[4] = f.ts.real(); out[5] = f.ts.imag(); out[6] = f.tp.real(); out[7] = f.tp.imag();
    out[8] = f.Ts; out[9] = f.Tp; out[10] = f.Z; out[11] = f.t.z;
}
void oracle_fresnel_full(float eta_re, float eta_im, const float w[3], float out[16]) {
    const auto f = fresnel(c_t{ eta_re, eta_im }, v3{ w[0], w[1], w[2] });
    out[0] = f.rs.real(); out[1] = f.rs.imag(); out[2] = f.rp.real(); out[3] = f.rp.imag();
    out[4] = f.ts.real(); out[5] = f.ts.imag(); out[6] = f.tp.real(); out[7] = f.tp.imag();
    out[8] = f.Ts; out[9] = f.Tp; out[10] = f.Z; out[11] = f.t.x; out[12] = f.t.y; out[13] = f.t.z; out[14] = f.eta_12.real(); out[15] = f.eta_12.imag();
}
void oracle_fresnel_reflection(float eta_re, float eta_im, const float w[3], float out[4]) {
    const auto f = fresnel_reflection(c_t{ eta_re, eta_im }, v3{ w[0], w[1], w[2] });
    out[0] = f.rs.real(); out[1] = f.rs.imag(); out[2] = f.rp.real(); out[3] = f.rp.imag();
}
void oracle_reflect(const float w[3], float out[3]) { const v3 r = reflect(v3{ w[0], w[1], w[2] }); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
// minimum-uncertainty sourcing: returns sbp (beam_geometry.hpp:43-47) of source_mub_from(length, k)
float oracle_mub_sbp(float length, float k) {
    const auto g = sourcing_geometry_t::source_mub_from_length(length, k);
    const auto e = g.phase_space_extent();
    const float area_stddev = e.spatial_extent / sqr(beam_cross_section_envelope);
    const float wv = sqr(e.k * e.tan_alpha / beam_cross_section_envelope);
    return area_stddev * wv * 1e6f;
}
// bsdf sampling energy: mean of weighted_bsdf mean intensity over n samples at incidence wi (albedo estimator)
float oracle_bsdf_albedo(const wtgpu_scene_desc* desc, int32_t bsdf, const float wi[3], float k, uint32_t n, uint64_t seed) {
    scene_t sc(desc); bsdf_eval_t be(sc);
    surface_t s; s.geo = frame_t::canonical(); s.shading = s.geo;
    const bsdf_query_t q{ &s, k, true };
    double acc = 0;
    for (uint32_t i = 0; i < n; ++i) {
        sampler_t smp; smp.seed = seed; smp.pixel = 0; smp.sample = i;
        const auto r = be.sample(bsdf, v3{ wi[0], wi[1], wi[2] }, q, smp);
        if (r) acc += r->M.mean_intensity();
    }
    return (float)(acc / n);
}
// surface-profile unit checks (surface_profile.hpp interface: alpha / psd / pdf / sample)
void oracle_profile_eval(const wtgpu_scene_desc* desc, int32_t bsdf, const float wi[3], const float wo[3], float k, float out[3]) {
    scene_t sc(desc); bsdf_eval_t be(sc);
    const wtgpu_bsdf& b = be.node(bsdf);
    const v3 i{ wi[0], wi[1], wi[2] }, o{ wo[0], wo[1], wo[2] };
    out[0] = be.profile_alpha(b, i, o, k); out[1] = be.profile_psd(b, i, o, k); out[2] = be.profile_pdf(b, i, o, k);
}
// n samples at incidence wi: out[0] = mean psd/pdf, out[1] / out[2] = largest relative deviation of pdf(wi, wo) / psd(wi, wo) evaluated at the
// sampled direction from the values the sampler returned, out[3] = fraction of samples outside the unit disk (clamped to grazing)
void oracle_profile_check(const wtgpu_scene_desc* desc, int32_t bsdf, const float wi[3], float k, uint32_t n, uint64_t seed, float out[4]) {
    scene_t sc(desc); bsdf_eval_t be(sc);
    const wtgpu_bsdf& b = be.node(bsdf);
    const v3 i{ wi[0], wi[1], wi[2] };
    double acc = 0, dpdf = 0, dpsd = 0, outside = 0;
    for (uint32_t j = 0; j < n; ++j) {
        sampler_t smp; smp.seed = seed; smp.pixel = 0; smp.sample = j;
        const auto r = be.profile_sample(b, i, k, smp);
        if (r.pdf > 0) acc += r.psd / r.pdf;
        if (dot(v2{ r.wo.x, r.wo.y }, v2{ r.wo.x, r.wo.y }) > 1) { outside += 1; continue; }
        const f_t pdf = be.profile_pdf(b, i, r.wo, k), psd = be.profile_psd(b, i, r.wo, k);
        if (r.pdf > 0) dpdf = std::max(dpdf, (double)std::fabs(pdf - r.pdf) / r.pdf);
        if (r.psd > 0) dpsd = std::max(dpsd, (double)std::fabs(psd - r.psd) / r.psd);
    }
    out[0] = (float)(acc / n); out[1] = (float)dpdf; out[2] = (float)dpsd; out[3] = (float)(outside / n);
}
// ---- checks of the two result-preserving shortcuts the PRODUCT takes and the reference does not (restated here as predicates):
// (1) the separating-axis rejection in front of the cone-triangle test (csrc/dmath.cuh intersect_cone_tri) and
// (2) the range culling of ray-query children (csrc/dtrav.cuh RayCull).  Each fuzzer draws random configurations, biased towards the
// borderline ones, and counts the cases where the shortcut would drop something the literal reference test accepts (must be 0).
static bool product_cone_quick_reject(const elliptic_cone_t& cone, v3 a, v3 b, v3 c, range_t range) {
    if (cone.is_ray() || !(cone.tan_alpha >= 0)) return false;
    const frame_t frame = cone.frame();
    const v3 o = cone.o();
    const v3 vs[3] = { frame.to_local(a - o), frame.to_local(b - o), frame.to_local(c - o) };
    const f_t closest_z = min3(vs[0].z, vs[1].z, vs[2].z), farthest_z = max3(vs[0].z, vs[1].z, vs[2].z);
    if (farthest_z < range.min || closest_z > range.max) return true;
    const f_t r = std::fma(std::min(farthest_z, range.max), cone.tan_alpha, cone.x0);
    if (!(r > 0 && r < inf)) return false;
    const f_t u0 = vs[0].x, u1 = vs[1].x, u2 = vs[2].x, v0 = vs[0].y * cone.e, v1 = vs[1].y * cone.e, v2 = vs[2].y * cone.e;
    const f_t ext = std::max(std::max(max3(std::fabs(u0), std::fabs(u1), std::fabs(u2)), max3(std::fabs(v0), std::fabs(v1), std::fabs(v2))), std::max(std::fabs(closest_z), std::fabs(farthest_z)));
    const f_t bb = r + (1e-3f * r + 1e-5f * ext), bd = 1.41421356f * r + (2e-3f * r + 2e-5f * ext);
    const f_t p0 = u0 + v0, p1 = u1 + v1, p2 = u2 + v2, q0 = u0 - v0, q1 = u1 - v1, q2 = u2 - v2;
    return min3(u0, u1, u2) > bb || max3(u0, u1, u2) < -bb || min3(v0, v1, v2) > bb || max3(v0, v1, v2) < -bb ||
           min3(p0, p1, p2) > bd || max3(p0, p1, p2) < -bd || min3(q0, q1, q2) > bd || max3(q0, q1, q2) < -bd;
}
struct fuzz_rng_t { uint64_t s; f_t u() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (f_t)((s >> 40) & 0xffffff) / 16777216.f; } f_t sym() { return 2 * u() - 1; } };
// out[0] = cases, out[1] = rejected by the shortcut, out[2] = accepted by the literal test, out[3] = VIOLATIONS (rejected but accepted)
void oracle_fuzz_cone_quick_reject(uint32_t n, uint64_t seed, uint64_t out[4]) {
    fuzz_rng_t g{ seed * 2654435761ull + 12345 };
    out[0] = out[1] = out[2] = out[3] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const f_t scale = std::pow(10.f, -3 + 5 * g.u());                       // scene scales 1 mm .. 100 m
        const v3 d = normalize(v3{ g.sym(), g.sym(), g.sym() + 1e-3f });
        const v3 o = scale * v3{ g.sym(), g.sym(), g.sym() };
        const f_t ta = g.u() < .1f ? 0.f : std::pow(10.f, -4 + 4 * g.u());
        const f_t x0 = g.u() < .2f ? 0.f : scale * std::pow(10.f, -4 + 3 * g.u());
        if (ta == 0 && x0 == 0) continue;
        const f_t ecc = g.u() < .5f ? 0.f : .95f * g.u();
        const frame_t of = frame_t::build_orthogonal_frame(d);
        const f_t phi = 6.2831853f * g.u();
        const v3 x = std::cos(phi) * of.t + std::sin(phi) * of.b;
        const auto cone = elliptic_cone_t::make_ecc(ray_t{ o, d }, x, ta, ecc, x0);
        const f_t z = scale * (g.u() < .1f ? 100 * g.u() : 3 * g.u());
        range_t range{ 0, inf };
        const f_t ru = g.u();
        if (ru < .3f) range = { z * g.u(), z * (1 + g.u()) }; else if (ru < .5f) range = { z * .5f * g.u(), inf };
        // triangle near the cone boundary at depth z: centre at (1 +- spread) radii from the axis, size from tiny to huge
        const f_t rad = ta * z + x0;
        const f_t psi = 6.2831853f * g.u();
        const f_t off = rad * (g.u() < .6f ? 1 + .2f * g.sym() : 3 * g.u());
        const f_t size = std::max(rad, 1e-6f * scale) * std::pow(10.f, -2 + 4 * g.u());
        const frame_t cf = cone.frame();
        const v3 ctr = o + z * d + off * (std::cos(psi) * cf.t + std::sin(psi) * cf.b / std::max(cone.e, 1e-3f));
        const v3 a = ctr + size * v3{ g.sym(), g.sym(), g.sym() }, b = ctr + size * v3{ g.sym(), g.sym(), g.sym() }, c = ctr + size * v3{ g.sym(), g.sym(), g.sym() };
        const v3 nn = cross(b - a, c - a);
        if (length(nn) == 0) continue;
        // triangles that collapse at f32 resolution (edges below ~1000 ulp of their distance from the cone's origin) are left out: there the literal
        // test's answer is a rounding artifact (it reports hits for zero-area triangles well outside the cone)
        if (size < 1e-4f * length(ctr - o)) continue;
        const v3 tn = normalize(nn);
        ++out[0];
        const bool rej = product_cone_quick_reject(cone, a, b, c, range);
        const bool acc = intersect_cone_tri(cone, a, b, c, tn, range).has_value();
        out[1] += rej; out[2] += acc; out[3] += (rej && acc);
        if (rej && acc && getenv("ORACLE_FUZZ_VERBOSE")) {
            const auto r = intersect_cone_tri(cone, a, b, c, tn, range);
            const frame_t fr = cone.frame();
            const v3 la = fr.to_local(a - o), lb = fr.to_local(b - o), lc = fr.to_local(c - o);
            fprintf(stderr, "violation: ta %g x0 %g e %g range [%g,%g] dist %g | a (%g %g %g) b (%g %g %g) c (%g %g %g) rad@z %g size %g\n", cone.tan_alpha, cone.x0, cone.e, range.min, range.max, r->dist,
                    la.x, la.y, la.z, lb.x, lb.y, lb.z, lc.x, lc.y, lc.z, rad, size);
        }
    }
}
// ray culling: random triangles and rays, the child's box is the triangle's AABB (what a leaf's parent stores); a violation is a triangle
// the literal ray test accepts within the range while the slab interval of its box, widened by the slack, misses the range
void oracle_fuzz_ray_cull(uint32_t n, uint64_t seed, uint64_t out[4]) {
    fuzz_rng_t g{ seed * 2654435761ull + 999 };
    out[0] = out[1] = out[2] = out[3] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const f_t scale = std::pow(10.f, -3 + 5 * g.u());
        const v3 o = scale * v3{ g.sym(), g.sym(), g.sym() };
        v3 d = normalize(v3{ g.sym(), g.sym(), g.sym() });
        if (g.u() < .2f) d = normalize(v3{ d.x, d.y * 1e-4f, d.z });                 // near axis-aligned
        const f_t t = scale * std::pow(10.f, -4 + 5 * g.u());
        const f_t size = scale * std::pow(10.f, -3 + 4 * g.u());
        const v3 ctr = o + t * d;
        v3 a = ctr + size * v3{ g.sym(), g.sym(), g.sym() }, b = ctr + size * v3{ g.sym(), g.sym(), g.sym() }, c = ctr + size * v3{ g.sym(), g.sym(), g.sym() };
        if (g.u() < .3f) { a.y = b.y = c.y = ctr.y; }                                 // flat (axis-aligned) triangles: degenerate boxes
        // range with an end near the hit distance
        const f_t ru = g.u();
        range_t range = ru < .4f ? range_t{ 0, t * (1 + 1e-3f * g.sym()) } : ru < .8f ? range_t{ t * (1 + 1e-3f * g.sym()), t * 4 } : range_t{ t * .5f, t * (1 + 1e-5f * g.sym()) };
        const f_t mabs = std::max(std::max(max3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)), max3(std::fabs(b.x), std::fabs(b.y), std::fabs(b.z))),
                                  std::max(max3(std::fabs(c.x), std::fabs(c.y), std::fabs(c.z)), max3(std::fabs(o.x), std::fabs(o.y), std::fabs(o.z))));
        const f_t cull_abs = 1e-5f * mabs;                                             // wtgpu_scene_create: 1e-5 x largest |coordinate| of the scene (>= this)
        const v3 mn{ min3(a.x, b.x, c.x), min3(a.y, b.y, c.y), min3(a.z, b.z, c.z) }, mx{ max3(a.x, b.x, c.x), max3(a.y, b.y, c.y), max3(a.z, b.z, c.z) };
        const v3 inv{ 1 / d.x, 1 / d.y, 1 / d.z };
        auto vmax = [](f_t p, f_t q) { return p > q ? p : q; }; auto vmin = [](f_t p, f_t q) { return p < q ? p : q; };
        const bool nx = std::signbit(inv.x), ny = std::signbit(inv.y), nz = std::signbit(inv.z);
        const f_t t1x = ((nx ? mx.x : mn.x) - o.x) * inv.x, t2x = ((nx ? mn.x : mx.x) - o.x) * inv.x;
        const f_t t1y = ((ny ? mx.y : mn.y) - o.y) * inv.y, t2y = ((ny ? mn.y : mx.y) - o.y) * inv.y;
        const f_t t1z = ((nz ? mx.z : mn.z) - o.z) * inv.z, t2z = ((nz ? mn.z : mx.z) - o.z) * inv.z;
        const f_t rmin = vmax(vmax(vmax(t1x, t1y), t1z), 0.f), rmax = vmin(vmin(vmin(t2x, t2y), t2z), inf);
        const f_t cmx = range.max + (1e-4f * range.max + cull_abs), cmn = range.min - (1e-4f * std::fabs(range.min) + cull_abs);
        const bool pushed = rmin <= rmax;
        const bool kept = pushed && !(rmin > cmx) && !(rmax < cmn);
        const bool acc = intersect_ray_tri_w(o, d, a, b, c, range).result != -inf;
        ++out[0]; out[1] += (pushed && !kept); out[2] += acc; out[3] += (acc && pushed && !kept);
    }
}
// cone-through-ellipse / ellipsoid re-fit (for unit parity with the device functions)
void oracle_cone_through_ellipsoid(const float axes[3], const float frame[9], const float o[3], const float d[3], float tan_alpha, float out[8]) {
    const frame_t f{ { frame[0], frame[1], frame[2] }, { frame[3], frame[4], frame[5] }, { frame[6], frame[7], frame[8] } };
    const auto c = elliptic_cone_t::cone_through_ellipsoid({ axes[0], axes[1], axes[2] }, f, ray_t{ { o[0], o[1], o[2] }, { d[0], d[1], d[2] } }, tan_alpha);
    out[0] = c.tangent.x; out[1] = c.tangent.y; out[2] = c.tangent.z; out[3] = c.x0; out[4] = c.e; out[5] = c.one_over_e; out[6] = c.tan_alpha; out[7] = c.z_apex;
}

// integral of the unit-covariance-scaled Gaussian over a triangle (gaussian2d_t::integrate_triangle)
float oracle_gaussian_integrate_triangle(float sx, float sy, const float tri[6]) {
    return gaussian2d_t(v2{ sx, sy }).integrate_triangle({ tri[0], tri[1] }, { tri[2], tri[3] }, { tri[4], tri[5] });
}
// the same for n triangles (tri: n x 6) -- same layout as oracle/ref_gaussian2d.cpp's ref_gaussian_integrate_triangles
void oracle_gaussian_integrate_triangles(float sx, float sy, uint32_t n, const float* tri, float* out) {
    const gaussian2d_t g(v2{ sx, sy });
    for (uint32_t i = 0; i < n; ++i) { const float* t = tri + 6 * i; out[i] = g.integrate_triangle({ t[0], t[1] }, { t[2], t[3] }, { t[4], t[5] }); }
}

// ot_scene.h's look-ups on caller-supplied tables -- same layouts as oracle/ref_distributions.cpp
void oracle_binned_eval(uint32_t n, const float* ys, const float* dcdf, float k0, float dk, float norm, uint32_t m, const float* v, float* icdf, const float* x, float* value, float* pdf) {
    for (uint32_t i = 0; i < m; ++i) {
        const v2 r = binned_icdf(ys, dcdf, n, k0, dk, v[i]); icdf[2 * i] = r.x; icdf[2 * i + 1] = r.y;
        value[i] = binned_value(ys, n, k0, dk, x[i]); pdf[i] = value[i] * norm;
    }
}
void oracle_gaussian1d_integrate(float sigma, uint32_t n, const float* mn, const float* mx, float* out) { for (uint32_t i = 0; i < n; ++i) out[i] = gaussian1d_integrate(sigma, mn[i], mx[i]); }
void oracle_discrete_icdf(uint32_t n, const float* dcdf, uint32_t m, const float* v, int* idx) { for (uint32_t i = 0; i < m; ++i) idx[i] = (int)discrete_icdf(dcdf, n, v[i]); }
void oracle_gaussian_pdf(float sx, float sy, uint32_t n, const float* pts, float* out) {
    const gaussian2d_t g(v2{ sx, sy });
    for (uint32_t i = 0; i < n; ++i) {
        const v2 p{ pts[2 * i], pts[2 * i + 1] };
        const v2 c = g.to_canonical(p);
        out[3 * i] = g.pdf(p); out[3 * i + 1] = c.x; out[3 * i + 2] = c.y;
    }
}
// clip_triangle_z + clip_ret_t::triangle for n triangles -- same layout as oracle/ref_clip.cpp's ref_clip_triangles
void oracle_clip_triangles(uint32_t n, const float* tri, const float* zr, int* ntris, float* polygon, float* pieces) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* t = tri + 9 * i;
        const auto r = clip_triangle_z({ t[0], t[1], t[2] }, { t[3], t[4], t[5] }, { t[6], t[7], t[8] }, range_t{ zr[2 * i], zr[2 * i + 1] });
        ntris[i] = r.tris;
        for (int k = 0; k < 5; ++k) { const bool used = r.tris > 0 && k < r.tris + 2; polygon[15 * i + 3 * k] = used ? r.vs[k].x : 0; polygon[15 * i + 3 * k + 1] = used ? r.vs[k].y : 0; polygon[15 * i + 3 * k + 2] = used ? r.vs[k].z : 0; }
        for (int j = 0; j < 3; ++j) {
            if (j >= r.tris) { for (int k = 0; k < 9; ++k) pieces[27 * i + 9 * j + k] = 0; continue; }
            v3 p[3]; r.triangle(j, p);
            for (int k = 0; k < 3; ++k) { pieces[27 * i + 9 * j + 3 * k] = p[k].x; pieces[27 * i + 9 * j + 3 * k + 1] = p[k].y; pieces[27 * i + 9 * j + 3 * k + 2] = p[k].z; }
        }
    }
}

// |sum_e Psi_e(xi)|^2 for explicit aperture edges (e.x,e.y,v.x,v.y,a_b.re,a_b.im,iab_2.re,iab_2.im each): Fraunhofer ASF (fsd.hpp:127-140)
float oracle_fraunhofer_asf(uint32_t n, const float* edges, float xix, float xiy) {
    ffsd::aperture_t ap;
    for (uint32_t i = 0; i < n; ++i) { const float* e = edges + 8 * i; ap.edges.push_back({ { e[0], e[1] }, { e[2], e[3] }, { e[4], e[5] }, { e[6], e[7] } }); }
    return ap.ASF_unclamped({ xix, xiy });
}

// everything fsd.hpp defines, for one aperture and one xi -- same layout as oracle/ref_fsd.cpp's ref_fsd_eval
void oracle_fsd_eval(uint32_t n, const float* edges, float P0v, float psi02, float xix, float xiy, float out[9]) {
    ffsd::aperture_t ap; ap.P0 = P0v; ap.psi02 = psi02;
    for (uint32_t i = 0; i < n; ++i) { const float* e = edges + 8 * i; ap.edges.push_back({ { e[0], e[1] }, { e[2], e[3] }, { e[4], e[5] }, { e[6], e[7] } }); }
    const v2 xi{ xix, xiy };
    out[0] = ap.ASF_unclamped(xi); out[1] = ap.ASF(xi); out[2] = ap.sampling_density(xi); out[3] = ffsd::chi_e(xi); out[4] = ffsd::chi_0(xi);
    out[5] = n ? ffsd::Pj(ap.edges[0]) : 0.f; out[6] = two_pi * sqr(ffsd::P0_sigma) * ap.psi02; out[7] = ffsd::alpha1(xi.x, xi.y); out[8] = ffsd::alpha2(xi.x, xi.y);
}

// ffsd::lut_t::sample for caller-supplied tables (theta: n, icdf: m x m); rand: cnt x 3, out: cnt x 2 -- same layout as oracle/ref_fsd_lut.cpp
void oracle_fsd_lut_sample(uint32_t n, uint32_t m, const float* theta, const float* icdf, uint32_t cnt, const float* rand, float* out) {
    const ffsd::lut_t lut{ n, m, theta, theta, icdf, icdf };
    for (uint32_t i = 0; i < cnt; ++i) { const v2 z = lut.sample(v3{ rand[3 * i], rand[3 * i + 1], rand[3 * i + 2] }, theta, icdf); out[2 * i] = z.x; out[2 * i + 1] = z.y; }
}

// fraunhofer_fsd_t::sample_rejection on a caller-supplied aperture and tables with a scripted number sequence -- same layout as
// oracle/ref_fsd_sampler.cpp's ref_fsd_sampler_sample; out: xi.x, xi.y, pdf, weight, numbers consumed
void oracle_fsd_sampler_sample(uint32_t n, uint32_t m, const float* th1, const float* th2, const float* c1, const float* c2, uint32_t n_edges, const float* edges,
                               const float* edge_pdfs, float P0v, float P0_pdf, float psi02, float recp_I, const float* script, uint32_t n_script, uint32_t n_samples, float* out) {
    const ffsd::lut_t lut{ n, m, th1, th2, c1, c2 };
    fraunhofer_fsd_t fs(fraunhofer_fsd_t::for_test_t{}, &lut);
    fs.ap.P0 = P0v; fs.ap.P0_pdf = P0_pdf; fs.ap.psi02 = psi02; fs.ap.recp_I = recp_I;
    for (uint32_t i = 0; i < n_edges; ++i) { const float* e = edges + 8 * i; fs.ap.edges.push_back({ { e[0], e[1] }, { e[2], e[3] }, { e[4], e[5] }, { e[6], e[7] } }); fs.ap.edge_pdfs.push_back(edge_pdfs[i]); }
    sampler_t smp; smp.script = script; smp.script_n = n_script;
    for (uint32_t s = 0; s < n_samples; ++s) {
        const auto r = fs.sample_rejection(smp);
        out[5 * s] = r.xi.x; out[5 * s + 1] = r.xi.y; out[5 * s + 2] = r.pdf; out[5 * s + 3] = r.weight; out[5 * s + 4] = (float)smp.d;
    }
}
// the warps of sampler.hpp:139-286 on explicit (u1, u2) -- same layout as ref_sampler_warps (uniform_hemisphere has no caller on the path: not restated)
void oracle_sampler_warps(float u1, float u2, float solid_angle, float out[13]) {
    const v2 u{ u1, u2 };
    const v3 a = cosine_hemisphere(u); const v2 b = concentric_disk(u); const v3 c = uniform_sphere(u); const v3 d = uniform_cone(solid_angle, u); const v2 e = normal2d(u);
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = b.x; out[4] = b.y; out[5] = c.x; out[6] = c.y; out[7] = c.z; out[8] = d.x; out[9] = d.y; out[10] = d.z; out[11] = e.x; out[12] = e.y;
}

float oracle_erf_lut(float x) { return erf_lut()(x); }
// pmath.h on the host (fn as wtgpu_debug_pmath); in the glibc build (OT_PORTABLE_LIBM=0) the same entry evaluates the host libm through lm::
void oracle_pmath(int fn, uint32_t n, const float* x, const float* y, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float a = x[i], b = y ? y[i] : 0.f; float r = 0.f;
        switch (fn) {
        case 0: r = lm::sin(a); break; case 1: r = lm::cos(a); break; case 2: r = lm::tan(a); break; case 3: r = lm::exp(a); break;
        case 4: r = lm::log(a); break; case 5: r = lm::pow(a, b); break; case 6: r = lm::atan2(a, b); break; case 7: r = lm::acos(a); break;
        case 8: r = lm::hypot(a, b); break; case 9: r = UTDF(a).real(); break; case 10: r = UTDF(a).imag(); break;
        }
        out[i] = r;
    }
}

} // extern "C"

extern "C" void oracle_debug_cone_hist(int on, uint64_t out[96]) {
    if (out) { for (int i = 0; i < 94; ++i) out[i] = ot::g_cone_hist[i].load(); out[94] = ot::g_cone_late[0].load(); out[95] = ot::g_cone_late[1].load(); }
    if (on >= 0) { ot::g_cone_hist_on = on; if (on) { for (int i = 0; i < 96; ++i) ot::g_cone_hist[i] = 0; ot::g_cone_late[0] = ot::g_cone_late[1] = 0; } }
}

// ---- frame_t (ot_math.h; include/wt/math/frame.hpp) for the pin against the reference's own header (oracle/ref_frame.cpp)
static void frame_put(const ot::frame_t& f, float* o) { o[0] = f.t.x; o[1] = f.t.y; o[2] = f.t.z; o[3] = f.b.x; o[4] = f.b.y; o[5] = f.b.z; o[6] = f.n.x; o[7] = f.n.y; o[8] = f.n.z; }
extern "C" void oracle_frame_orthogonal(uint32_t n, const float* nrm, float* out) { for (uint32_t i = 0; i < n; ++i) frame_put(ot::frame_t::build_orthogonal_frame({ nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2] }), out + 9 * i); }
extern "C" void oracle_frame_shading(uint32_t n, const float* nrm, const float* dpdu, float* out) {
    for (uint32_t i = 0; i < n; ++i) frame_put(ot::frame_t::build_shading_frame({ nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2] }, { dpdu[3 * i], dpdu[3 * i + 1], dpdu[3 * i + 2] }), out + 9 * i);
}
extern "C" void oracle_frame_xform(uint32_t n, const float* fr, const float* v, float* out) {      // same 21 values per item as ref_frame_xform
    for (uint32_t i = 0; i < n; ++i) {
        const float* f = fr + 9 * i; const float* p = v + 3 * i; float* o = out + 21 * i;
        const ot::frame_t F{ { f[0], f[1], f[2] }, { f[3], f[4], f[5] }, { f[6], f[7], f[8] } };
        const ot::v3 a = F.to_local(ot::v3{ p[0], p[1], p[2] }), b = F.to_world(ot::v3{ p[0], p[1], p[2] });
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = b.x; o[4] = b.y; o[5] = b.z; o[6] = a.x; o[7] = a.y; o[8] = a.z; o[9] = b.x; o[10] = b.y; o[11] = b.z;
        const ot::v2 e = F.to_local2(ot::v2{ p[0], p[1] }); o[12] = e.x; o[13] = e.y;
        const ot::v3 g = F.to_world(ot::v2{ p[0], p[1] }); o[14] = g.x; o[15] = g.y; o[16] = g.z;
        o[17] = a.x; o[18] = a.y; o[19] = a.z;
        o[20] = F.handness();
    }
}
extern "C" void oracle_rotation2(uint32_t n, const float* from, const float* to, float* out) {      // ot_math.h rotation_matrix (math/rotation.hpp:66-77), column-major like ref_rotation2
    for (uint32_t i = 0; i < n; ++i) { const ot::mat2 R = ot::rotation_matrix(ot::v2{ from[2 * i], from[2 * i + 1] }, ot::v2{ to[2 * i], to[2 * i + 1] }); out[4 * i] = R.m[0][0]; out[4 * i + 1] = R.m[0][1]; out[4 * i + 2] = R.m[1][0]; out[4 * i + 3] = R.m[1][1]; }
}

// ---- the edge tests of ot_math.h (include/wt/math/intersect/misc.hpp) for the pin against the reference's own header (oracle/ref_misc.cpp); same layouts
extern "C" void oracle_edge_ellipsoid(uint32_t n, const float* in, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 18 * i;
        const auto r = ot::intersect_edge_ellipsoid({ a[0], a[1], a[2] }, { a[3], a[4], a[5] }, { a[6], a[7], a[8] }, { a[9], a[10], a[11] }, { a[12], a[13], a[14] }, { a[15], a[16], a[17] });
        out[2 * i] = r.t1; out[2 * i + 1] = r.t2;
    }
}
extern "C" void oracle_edge_ellipse(uint32_t n, const float* in, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 6 * i; float* o = out + 14 * i;
        const auto e = ot::intersect_edge_ellipse({ a[0], a[1] }, { a[2], a[3] }, a[4], a[5], false), l = ot::intersect_edge_ellipse({ a[0], a[1] }, { a[2], a[3] }, a[4], a[5], true);
        o[0] = (float)e.points; o[1] = e.t1; o[2] = e.t2; o[3] = e.u1.x; o[4] = e.u1.y; o[5] = e.u2.x; o[6] = e.u2.y;
        o[7] = (float)l.points; o[8] = l.t1; o[9] = l.t2; o[10] = l.u1.x; o[11] = l.u1.y; o[12] = l.u2.x; o[13] = l.u2.y;
    }
}
extern "C" void oracle_edge_plane(uint32_t n, const float* in, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 12 * i; float* o = out + 4 * i;
        const auto r = ot::intersect_edge_plane({ a[0], a[1], a[2] }, { a[3], a[4], a[5] }, { a[6], a[7], a[8] }, { a[9], a[10], a[11] });
        o[0] = r ? 1.f : 0.f; o[1] = r ? r->x : 0.f; o[2] = r ? r->y : 0.f; o[3] = r ? r->z : 0.f;
    }
}
// the cone stages of intersect_cone_tri, laid out like oracle/ref_cone.cpp
static ot::elliptic_cone_t kat_cone(const float* c) {
    return ot::elliptic_cone_t::make_ecc(ot::ray_t{ { c[0], c[1], c[2] }, { c[3], c[4], c[5] } }, { c[6], c[7], c[8] }, c[9], c[10], c[11]);
}
extern "C" void oracle_cone_edge(uint32_t n, int in_local, const float* in, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 20 * i; float* o = out + 10 * i;
        const auto r = ot::intersect_cone_edge(kat_cone(a), { a[12], a[13], a[14] }, { a[15], a[16], a[17] }, { a[18], a[19] }, in_local != 0);
        for (int k = 0; k < 10; ++k) o[k] = 0.f;
        if (!r) continue;
        o[0] = 1.f; o[1] = r->p0.x; o[2] = r->p0.y; o[3] = r->p0.z;
        if (r->pts == 2) { o[4] = r->p1.x; o[5] = r->p1.y; o[6] = r->p1.z; }
        o[7] = r->range.min; o[8] = r->range.max; o[9] = (float)r->pts;
    }
}
extern "C" void oracle_cone_plane(uint32_t n, int in_local, const float* in, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 18 * i; float* o = out + 9 * i;
        const auto r = ot::intersect_cone_plane(kat_cone(a), { a[12], a[13], a[14] }, a[15], { a[16], a[17] }, in_local != 0);
        for (int k = 0; k < 9; ++k) o[k] = 0.f;
        if (r.range.empty()) continue;
        o[0] = 1.f; o[1] = r.range.min; o[2] = r.range.max;
        o[3] = r.near_.x; o[4] = r.near_.y; o[5] = r.near_.z; o[6] = r.far_.x; o[7] = r.far_.y; o[8] = r.far_.z;
    }
}
extern "C" void oracle_cone_tri(uint32_t n, const float* in, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 26 * i; float* o = out + 6 * i;
        const auto cone = kat_cone(a);
        const ot::v3 A{ a[12], a[13], a[14] }, B{ a[15], a[16], a[17] }, Cc{ a[18], a[19], a[20] };
        const ot::range_t range{ a[24], a[25] };
        const auto h = ot::intersect_cone_tri(cone, A, B, Cc, { a[21], a[22], a[23] }, range);
        o[0] = h ? 1.f : 0.f; o[1] = h ? h->dist : 0.f; o[2] = h ? h->p.x : 0.f; o[3] = h ? h->p.y : 0.f; o[4] = h ? h->p.z : 0.f;
        o[5] = ot::test_cone_tri(cone, A, B, Cc, range) ? 1.f : 0.f;
    }
}
extern "C" void oracle_ray_tri(uint32_t n, const float* in, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 18 * i; float* o = out + 8 * i;
        const ot::ray_t r{ { a[0], a[1], a[2] }, { a[3], a[4], a[5] } };
        const ot::v3 A{ a[6], a[7], a[8] }, B{ a[9], a[10], a[11] }, Cc{ a[12], a[13], a[14] };
        const ot::range_t range{ a[15], a[16] };
        const auto h = ot::intersect_ray_tri(r, A, B, Cc, range);
        o[0] = h ? 1.f : 0.f; o[1] = h ? h->dist : 0.f; o[2] = h ? h->bary.x : 0.f; o[3] = h ? h->bary.y : 0.f;
        o[4] = ot::test_ray_tri(r, A, B, Cc, range) ? 1.f : 0.f;
        o[5] = ot::test_ray_tri(r, A, B, Cc, range, a[17]) ? 1.f : 0.f;
        const auto lp = ot::intersect_line_plane(A, B, Cc, r.d);
        o[6] = lp ? 1.f : 0.f; o[7] = lp ? *lp : 0.f;
    }
}
extern "C" void oracle_ray_tri_w(uint32_t n, const float* in, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 17 * i; const float* a0 = in + 17 * (i & ~7u); float* o = out + 4 * i;
        const ot::range_t range{ a0[15], a0[16] };
        const ot::v3 ro{ a[0], a[1], a[2] }, rd{ a[3], a[4], a[5] }, A{ a[6], a[7], a[8] }, B{ a[9], a[10], a[11] }, Cc{ a[12], a[13], a[14] };
        const auto r = ot::intersect_ray_tri_w(ro, rd, A, B, Cc, range);
        o[0] = r.result; o[1] = r.baryx; o[2] = r.baryy; o[3] = ot::test_ray_tri_w(ro, rd, A, B, Cc, range) ? 1.f : 0.f;
    }
}
extern "C" void oracle_ray_aabb_fast(uint32_t n, const float* in, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 14 * i; const float* a0 = in + 14 * (i & ~7u); float* o = out + 3 * i;
        const auto r = ot::ray_aabb_fast({ a[0], a[1], a[2] }, { a[3], a[4], a[5] }, { a[6], a[7], a[8] }, { a[9], a[10], a[11] }, { a0[12], a0[13] });
        o[0] = r.mask ? 1.f : 0.f; o[1] = r.min; o[2] = r.max;
    }
}
// per query in: o[3] d[3] x[3] tan_alpha eccentricity x0 tmin tmax z_scale; out as oracle/ref_traverse.cpp's ref_traverse_cones: the accepted triangles in
// traversal order, their number, the closest distance, the face flag
extern "C" void oracle_cone_work_lists(const wtgpu_scene_desc* desc, uint32_t n, const float* q, uint32_t cap, uint32_t* counts, uint32_t* tuids, float* dist, uint32_t* front) {
    ads_t ads(desc);
    for (uint32_t i = 0; i < n; ++i) {
        const float* c = q + 15 * i;
        const auto rec = intersect_cone(ads, kat_cone(c), { c[12], c[13] }, c[14], false);
        counts[i] = (uint32_t)rec.tris.size(); dist[i] = rec.tris.empty() ? inf : rec.dist; front[i] = rec.front_face ? 1u : 0u;
        for (uint32_t k = 0; k < cap; ++k) tuids[(size_t)i * cap + k] = k < rec.tris.size() ? rec.tris[k] : 0xffffffffu;
    }
}
static void kat_put_cone(const ot::elliptic_cone_t& c, float* o) {
    o[0] = c.x().x; o[1] = c.x().y; o[2] = c.x().z; o[3] = c.x0; o[4] = c.e; o[5] = c.one_over_e; o[6] = c.tan_alpha; o[7] = c.z_apex;
}
extern "C" void oracle_cone_through_ellipse_n(uint32_t n, const float* in, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 16 * i; float* o = out + 9 * i;
        f_t sid = 0;
        const auto c = ot::elliptic_cone_t::cone_through_ellipse({ a[0], a[1], a[2] }, { a[3], a[4], a[5] }, { a[6], a[7], a[8] }, ot::ray_t{ { a[9], a[10], a[11] }, { a[12], a[13], a[14] } }, a[15], &sid);
        kat_put_cone(c, o); o[8] = sid;
    }
}
extern "C" void oracle_cone_through_ellipsoid_n(uint32_t n, const float* in, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 19 * i;
        const ot::frame_t F{ { a[3], a[4], a[5] }, { a[6], a[7], a[8] }, { a[9], a[10], a[11] } };
        kat_put_cone(ot::elliptic_cone_t::cone_through_ellipsoid({ a[0], a[1], a[2] }, F, ot::ray_t{ { a[12], a[13], a[14] }, { a[15], a[16], a[17] } }, a[18]), out + 8 * i);
    }
}
// Mueller / Stokes algebra, laid out like oracle/ref_mueller.cpp
extern "C" void oracle_mueller(uint32_t n, const float* in, float* out) {
    auto load_m = [](const float* a) { ot::mueller_t M; for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) M.m[c][r] = a[4 * c + r]; return M; };
    auto put_m = [](const ot::mueller_t& M, float* o) { for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) o[4 * c + r] = M.m[c][r]; };
    auto load_f = [](const float* f) { return ot::frame_t{ { f[0], f[1], f[2] }, { f[3], f[4], f[5] }, { f[6], f[7], f[8] } }; };
    auto put_s = [](const ot::stokes_t& s, float* o) { for (int i = 0; i < 4; ++i) o[i] = s.S[i]; };
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 89 * i; float* o = out + 144 * i;
        const auto A = load_m(a), B = load_m(a + 16);
        const ot::stokes_t S{ { a[32], a[33], a[34], a[35] } };
        const ot::frame_t F1 = load_f(a + 36), F2 = load_f(a + 45), F3 = load_f(a + 54), F4 = load_f(a + 63);
        const ot::v2 t1{ a[72], a[73] }, t2{ a[74], a[75] };
        const c_t fs{ a[76], a[77] }, fp{ a[78], a[79] }, eta{ a[80], a[81] };
        const ot::v3 w{ a[82], a[83], a[84] };
        put_m(A * B, o); put_s(A * S, o + 16);
        put_m(ot::mueller_t::rotation(t1, t2), o + 20);
        put_m(ot::mueller_t::fresnel(fs, fp), o + 36);
        put_m(ot::mueller_fresnel_reflection(eta, w), o + 52);
        put_m(ot::mueller_fresnel_transmission(eta, w), o + 68);
        put_m(ot::change_incident_frame(A, F1, F2), o + 84);
        put_m(ot::change_exitant_frame(A, F1, F2), o + 100);
        put_m(ot::compose(A, B, F1, F2), o + 116);
        put_s(S.reorient(F1, F2), o + 132);
        put_s(ot::mueller_apply(A, S, F1, F2), o + 136);
        put_s(ot::mueller_apply(A, S, F1, F2, F3, F4), o + 140);
    }
}
// integrator::traverse, laid out like oracle/ref_traverse.cpp's ref_integrator_traverse
extern "C" void oracle_integrator_traverse(const wtgpu_scene_desc* desc, uint32_t n, const float* q, uint32_t cap, float* out, uint32_t* ntris, uint32_t* tris, uint32_t* nedges, uint32_t* edges) {
    scene_t sc(desc);
    for (uint32_t i = 0; i < n; ++i) {
        const float* c = q + 16 * i; float* o = out + 12 * i;
        const auto r = traverse(sc, kat_cone(c), c[12], c[13], c[14] != 0, c[15] != 0, nullptr);
        const bool rt = r.ballistic && !r.empty;
        o[0] = r.empty; o[1] = r.ballistic; o[2] = r.origin.x; o[3] = r.origin.y; o[4] = r.origin.z; o[5] = r.empty ? 0.f : r.distance(); o[6] = r.intersection_region_depth;
        o[7] = !r.empty && (r.ballistic ? r.ray.front_face : r.cone.front_face); o[8] = rt; o[9] = rt ? r.ray.bary.x : 0.f; o[10] = rt ? r.ray.bary.y : 0.f; o[11] = 0;
        std::vector<uint32_t> tl, el;
        if (rt) tl.push_back(r.ray.tuid); else if (!r.empty) { tl = r.cone.tris; el = r.cone.edges; }
        ntris[i] = (uint32_t)tl.size(); nedges[i] = (uint32_t)el.size();
        for (uint32_t k = 0; k < cap; ++k) { tris[(size_t)i * cap + k] = k < tl.size() ? tl[k] : 0xffffffffu; edges[(size_t)i * cap + k] = k < el.size() ? el[k] : 0xffffffffu; }
    }
}
// plt_path_t::find_closest_triangle, laid out like oracle/ref_traverse.cpp's ref_find_closest_triangle
extern "C" void oracle_find_closest_triangle(const wtgpu_scene_desc* desc, uint32_t n, const float* q, float* out, uint32_t* tuid) {
    scene_t sc(desc); film_t film(sc); path_stats_t st;
    plt_path_t integ(sc, film, &st);
    std::vector<uint32_t> list;
    for (uint32_t i = 0; i < n; ++i) {
        const float* c = q + 10 * i;
        list.clear(); for (uint32_t k = 0; k < (uint32_t)c[9]; ++k) list.push_back((uint32_t)c[8] + k);
        const auto id = integ.find_closest_triangle(list, { c[6], c[7] }, { c[0], c[1], c[2] }, { c[3], c[4], c[5] });
        const bool f = id.primary != WTGPU_INVALID_IDX;
        tuid[i] = f ? id.primary : 0xffffffffu;
        out[3 * i] = f ? id.dist : 0.f; out[3 * i + 1] = f ? id.bary.x : 0.f; out[3 * i + 2] = f ? id.bary.y : 0.f;
    }
}
// self-intersection offsets, laid out like oracle/ref_traverse.cpp's ref_edge_offsets
extern "C" void oracle_edge_offsets(const wtgpu_scene_desc* desc, uint32_t n, const float* q, float* out) {
    scene_t sc(desc);
    for (uint32_t i = 0; i < n; ++i) {
        const float* c = q + 7 * i; float* o = out + 6 * i;
        const uint32_t ei = (uint32_t)c[0];
        const ray_t ray{ { c[1], c[2], c[3] }, { c[4], c[5], c[6] } };
        const v3 p = sc.offseted_ray_origin_edge(ei, ray);
        const uint32_t t1 = desc->edges[ei].tri1;
        const v3 err = scene_t::tri_fp_errors(sc.ads.tri_a(t1), sc.ads.tri_b(t1), sc.ads.tri_c(t1), ray.o);
        o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = err.x; o[4] = err.y; o[5] = err.z;
    }
}
// plt_bdpt_t::find_closest_triangle, laid out like oracle/ref_traverse.cpp's ref_bd_find_closest_triangle
extern "C" void oracle_bd_find_closest_triangle(const wtgpu_scene_desc* desc, uint32_t n, const float* q, float* out, uint32_t* tuid) {
    scene_t sc(desc); film_t film(sc); bdpt_stats_t st;
    plt_bdpt_t integ(sc, film, &st);
    std::vector<uint32_t> list;
    for (uint32_t i = 0; i < n; ++i) {
        const float* c = q + 28 * i; float* o = out + 4 * i;
        list.clear(); for (uint32_t k = 0; k < (uint32_t)c[9]; ++k) list.push_back((uint32_t)c[8] + k);
        const v3 dir{ c[3], c[4], c[5] };
        const frame_t bf{ { c[10], c[11], c[12] }, { c[13], c[14], c[15] }, dir };
        const auto env = elliptic_cone_t::make_ecc(ray_t{ { c[16], c[17], c[18] }, dir }, { c[19], c[20], c[21] }, c[22], c[23], c[24]);
        const wavefront_t wf(v2{ c[25], c[26] });
        const auto id = integ.find_closest_triangle(list, { c[6], c[7] }, { c[0], c[1], c[2] }, dir, bf, env, wf, c[27] != 0);
        const bool f = id.primary != WTGPU_INVALID_IDX;
        tuid[i] = f ? id.primary : 0xffffffffu;
        o[0] = f ? id.dist : 0.f; o[1] = f ? id.bary.x : 0.f; o[2] = f ? id.bary.y : 0.f; o[3] = id.integrated_radiant_flux;
    }
}
// fraunhofer_fsd_t's constructor, laid out like oracle/ref_traverse.cpp's ref_ffsd_aperture
extern "C" void oracle_ffsd_aperture(const wtgpu_scene_desc* desc, uint32_t n, const float* q, uint32_t cap, uint32_t* counts, float* summary, float* edges) {
    scene_t sc(desc);
    for (uint32_t i = 0; i < n; ++i) {
        const float* c = q + 27 * i;
        const auto beam = kat_cone(c);
        const frame_t fr{ { c[12], c[13], c[14] }, { c[15], c[16], c[17] }, { c[18], c[19], c[20] } };
        const wavefront_t wf(v2{ c[23], c[24] });
        std::vector<uint32_t> es; for (uint32_t k = 0; k < (uint32_t)c[26]; ++k) es.push_back((uint32_t)c[25] + k);
        const fraunhofer_fsd_t f(sc, nullptr, fr, c[21], c[22], beam, es, wf);
        const auto& ap = f.ap;
        counts[i] = (uint32_t)ap.edges.size();
        summary[4 * i] = ap.recp_I; summary[4 * i + 1] = ap.psi02; summary[4 * i + 2] = ap.P0; summary[4 * i + 3] = ap.P0_pdf;
        for (uint32_t k = 0; k < cap; ++k) {
            float* o = edges + ((size_t)i * cap + k) * 9;
            if (k < ap.edges.size()) { const auto& e = ap.edges[k]; o[0] = e.e.x; o[1] = e.e.y; o[2] = e.v.x; o[3] = e.v.y; o[4] = e.a_b.real(); o[5] = e.a_b.imag(); o[6] = e.iab_2.real(); o[7] = e.iab_2.imag(); o[8] = ap.edge_pdfs[k]; }
            else for (int j = 0; j < 9; ++j) o[j] = 0.f;
        }
    }
}
// fsd_t::build + f(), laid out like oracle/ref_traverse.cpp's ref_utd_fsd
extern "C" void oracle_utd_fsd(const wtgpu_scene_desc* desc, uint32_t n, const float* q, uint32_t cap, uint32_t* nap, float* ap, uint32_t* nf, float* fo) {
    scene_t sc(desc);
    for (uint32_t i = 0; i < n; ++i) {
        const float* c = q + 28 * i;
        const frame_t fr{ { c[3], c[4], c[5] }, { c[6], c[7], c[8] }, { c[9], c[10], c[11] } };
        std::vector<uint32_t> es; for (uint32_t k = 0; k < (uint32_t)c[21]; ++k) es.push_back((uint32_t)c[20] + k);
        const fsd_t f = fsd_t::build(sc, { c[0], c[1], c[2] }, fr, { c[12], c[13], c[14] }, { c[15], c[16], c[17] }, c[18], es);
        nap[i] = (uint32_t)f.edges.size();
        for (uint32_t k = 0; k < cap; ++k) {
            float* o = ap + ((size_t)i * cap + k) * 15; for (int j = 0; j < 15; ++j) o[j] = 0.f;
            if (k >= f.edges.size()) continue;
            const auto& e = f.edges[k];
            o[0] = e.v.x; o[1] = e.v.y; o[2] = e.v.z; o[3] = e.l; o[4] = e.nff.x; o[5] = e.nff.y; o[6] = e.nff.z; o[7] = e.tff.x; o[8] = e.tff.y; o[9] = e.tff.z;
            o[10] = e.nbf.x; o[11] = e.nbf.y; o[12] = e.nbf.z; o[13] = e.alpha; o[14] = (float)e.ads_edge_idx;
        }
        const auto r = f.f({ c[22], c[23], c[24] }, { c[25], c[26], c[27] });
        nf[i] = (uint32_t)r.size();
        for (uint32_t k = 0; k < cap; ++k) {
            float* o = fo + ((size_t)i * cap + k) * 10; for (int j = 0; j < 10; ++j) o[j] = 0.f;
            if (k >= r.size()) continue;
            const auto& d = r[k];
            o[0] = (float)d.edge_idx; o[1] = d.p.x; o[2] = d.p.y; o[3] = d.p.z; o[4] = d.ri; o[5] = d.ro;
            o[6] = d.utd.Ds.real(); o[7] = d.utd.Ds.imag(); o[8] = d.utd.Dh.real(); o[9] = d.utd.Dh.imag();
        }
    }
}
extern "C" void oracle_cone_cluster(uint32_t n, const float* in, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 20 * i; const float* a0 = in + 20 * (i & ~7u);
        const auto cone = kat_cone(a0);
        f_t tmin = 0;
        const bool hit = ot::cone_cluster_lane(cone.o(), cone.d(), cone.r.invd, cone.tan_alpha, cone.x0, { a[12], a[13], a[14] }, { a[15], a[16], a[17] }, { a0[18], a0[19] }, tmin);
        out[2 * i] = hit ? 1.f : 0.f; out[2 * i + 1] = tmin;
    }
}
extern "C" void oracle_stack_sorter(uint32_t n, uint32_t run, float* io) {
    std::vector<ot::stack_node_ptr_t> st(run);
    for (uint32_t i = 0; i + run <= n; i += run) {
        for (uint32_t k = 0; k < run; ++k) st[k] = { io[2 * (i + k)], (int32_t)io[2 * (i + k) + 1] };
        ot::stack_sorter(st.data(), (int)run);
        for (uint32_t k = 0; k < run; ++k) { io[2 * (i + k)] = st[k].min_range; io[2 * (i + k) + 1] = (float)st[k].ptr; }
    }
}
extern "C" void oracle_cone_basics(uint32_t n, const float* in, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 13 * i; float* o = out + 5 * i;
        const auto cone = kat_cone(a);
        const auto ax = cone.axes(a[12]);
        o[0] = ax.x; o[1] = ax.y; o[2] = cone.z_apex; o[3] = cone.e; o[4] = cone.one_over_e;
    }
}
extern "C" void oracle_point_in_triangle3(uint32_t n, const float* in, float* out) {
    for (uint32_t i = 0; i < n; ++i) { const float* a = in + 12 * i; out[i] = ot::is_point_in_triangle({ a[0], a[1], a[2] }, { a[3], a[4], a[5] }, { a[6], a[7], a[8] }, { a[9], a[10], a[11] }) ? 1.f : 0.f; }
}
extern "C" void oracle_point_in_triangle2(uint32_t n, const float* in, float* out) {
    for (uint32_t i = 0; i < n; ++i) { const float* a = in + 8 * i; out[i] = ot::g2d::point_in_triangle2({ a[0], a[1] }, { a[2], a[3] }, { a[4], a[5] }, { a[6], a[7] }) ? 1.f : 0.f; }
}
