// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// ot_bdpt.h: plt_bdpt (bidirectional) -- restating
//   src/integrator/plt_bdpt.cpp:43-148, include/wt/integrator/plt_bdpt/{plt_bdpt_detail.hpp,vertex.hpp},
//   Fraunhofer free-space diffraction (include/wt/interaction/fsd/fraunhofer/*.hpp, src/interaction/fsd/fraunhofer/*.cpp),
//   gaussian2d_t::integrate_triangle (src/math/gaussian2d.cpp:96-192), clip_triangle_z (include/wt/math/intersect/clip.hpp).
#pragma once
#ifdef OT_DEBUG_COUNTERS
#include <atomic>
namespace ot { inline std::atomic<unsigned long long> g_dbg[8]; }
#define OT_DBG(i, n) (ot::g_dbg[i] += (n))
#define OT_DBG_MAX(i, n) do { unsigned long long o_ = ot::g_dbg[i]; while (o_ < (n) && !ot::g_dbg[i].compare_exchange_weak(o_, (n))) {} } while (0)
#else
#define OT_DBG(i, n) ((void)0)
#define OT_DBG_MAX(i, n) ((void)0)
#endif
#include "ot_integrator.h"

namespace ot {

static constexpr f_t sqrt_pi = 1.77245385090551602730f;
static constexpr f_t inv_sqrt_pi = 0.56418958354775628695f;

// boost-style sinc (include/wt/math/common.hpp:414-434)
inline f_t sinc(f_t x) {
    const f_t t0 = std::numeric_limits<f_t>::epsilon(), t2 = 0.00034526698300124390839884978618400831996329879769945f, tn = 0.018581361171917516667460937040007436176452688944747f;
    if (std::fabs(x) >= tn) return lm::sin(x) / x;
    f_t r = 1;
    if (std::fabs(x) >= t0) { const f_t x2 = x * x; r -= x2 / 6.f; if (std::fabs(x) >= t2) r += (x2 * x2) / 120.f; }
    return r;
}

// ---- gaussian2d_t with x = (1,0), mu = 0 (include/wt/math/distribution/gaussian2d.hpp; all wavefronts are built that way: beam_generic.hpp:130-139)
struct gaussian2d_t {
    v2 sigma{ 1, 1 }, recp_sigma{ 1, 1 };
    f_t norm = inv_two_pi;
    explicit gaussian2d_t(v2 s = { 1, 1 }) : sigma(s), recp_sigma{ 1.f / s.x, 1.f / s.y }, norm(inv_two_pi * (1.f / s.x) * (1.f / s.y)) {}
    bool is_dirac() const { return sigma.x == 0 || sigma.y == 0; }
    f_t pdf(v2 p) const {
        const v2 u = p * recp_sigma;
        return !is_dirac() ? norm * lm::exp(-dot(u, u) / 2) : ((p.x == 0 && p.y == 0) ? inf : 0.f);
    }
    v2 to_canonical(v2 v) const {
        const v2 p{ dot(v2{ 1, 0 }, v), dot(v2{ -0.f, 1 }, v) };
        if (!is_dirac()) return p * recp_sigma;
        return { p.x == 0 ? 0.f : inf, p.y == 0 ? 0.f : inf };
    }
    f_t integrate_triangle(v2 a, v2 b, v2 c) const;
};

namespace g2d {      // src/math/gaussian2d.cpp:24-94
inline f_t erf_(f_t x) { return erf_lut()(x); }
inline f_t I_gauss_gauss0(f_t a, f_t b, f_t c, f_t d) {
    const f_t n2 = 1 / (a + 2 * c * c), n = std::sqrt(n2);
    return -sqrt_pi / 2 * n * lm::exp(-2 * a * sqr(d - b * c) * n2) * (erf_((a * b + 2 * c * d) * n) - erf_((a * (1 + b) + 2 * c * (c + d)) * n));
}
inline f_t I_gauss_gauss1(f_t a, f_t b, f_t c, f_t d) {
    const f_t n2 = 1 / (a + 2 * c * c), n = std::sqrt(n2);
    return -sqrt_pi / 2 * n * lm::exp(-2 * a * sqr(d - b * c) * n2) * (2 * erf_(a * (d / c - b) * n) + erf_((a * b + 2 * c * d) * n) + erf_((a * (1 + b) + 2 * c * (c + d)) * n));
}
inline f_t I_gauss0(f_t a, f_t b) { const f_t n = std::sqrt(1 / a); return -sqrt_pi / 2 * n * (erf_(a * b * n) - erf_(a * (1 + b) * n)); }
inline f_t I_gauss1(f_t a, f_t b, f_t c, f_t d) {
    const f_t sa = std::sqrt(a), n = 1 / sa, d_c = d / c;
    return -sqrt_pi / 2 * n * (sign(b) * erf_(sa * std::fabs(b)) + sign(1 + b) * erf_(sa * std::fabs(1 + b)) - 2 * sign(b - d_c) * erf_(sa * std::fabs(b - d_c)));
}
inline f_t I_gauss_erf(f_t a, f_t b, f_t c, f_t d) {
    const f_t d_c = d / c;
    const bool in = c != 0 && -d_c > 0 && -d_c < 1;
    auto gg = [&](f_t s) { const f_t q = std::sqrt(s); return in ? I_gauss_gauss1(a, b, c * q, d * q) : I_gauss_gauss0(a, b, c * q, d * q); };
    return ((d != 0 && c != 0) ? sign(d) : (d == 0 && c != 0) ? sign(c) : 1.f) *
           ((in ? I_gauss1(a, b, c, d) : I_gauss0(a, b)) -
            2 * (0.2936683276537767f * gg(0.6517755981618476f) + 0.135758042187825f * gg(3.250040490513459f) +
                 0.05245255757691102f * gg(31.86882707224491f) + 0.01673209873360605f * gg(778.6613983601425f)));
}
inline bool point_in_triangle2(v2 p, v2 a, v2 b, v2 c) {        // math/util.hpp:69-82
    auto sgn = [](v2 p1, v2 p2, v2 p3) { return diff_prod(p1.x - p3.x, p2.y - p3.y, p2.x - p3.x, p1.y - p3.y); };
    const f_t s1 = sgn(p, a, b), s2 = sgn(p, b, c), s3 = sgn(p, c, a);
    const bool neg = s1 < 0 || s2 < 0 || s3 < 0, pos = s1 > 0 || s2 > 0 || s3 > 0;
    return !(neg && pos);
}
}

// src/math/gaussian2d.cpp:96-192
inline f_t gaussian2d_t::integrate_triangle(v2 a, v2 b, v2 c) const {
    if (is_dirac()) {   // barycentric_if_point_inside(a,b,c, mu=0) (math/barycentric.hpp:96-116)
        const f_t A = (a.x * (b.y - c.y) - a.y * (b.x - c.x)) + (b.x * c.y - c.x * b.y);   // det [[a,1],[b,1],[c,1]]
        const f_t sA = sign(A);
        const f_t bx = sA * diff_prod(b.x, c.y, c.x, b.y), by = sA * diff_prod(c.x, a.y, a.x, c.y);
        return (bx >= 0 && by >= 0 && bx + by <= std::fabs(A)) ? 1.f : 0.f;
    }
    const f_t L = 3;
    a = to_canonical(a); b = to_canonical(b); c = to_canonical(c);
    if (min3(a.x, b.x, c.x) >= L || max3(a.x, b.x, c.x) <= -L || min3(a.y, b.y, c.y) >= L || max3(a.y, b.y, c.y) <= -L) return 0;
    const bool ain = length2(a) <= sqr(L), bin = length2(b) <= sqr(L), cin = length2(c) <= sqr(L);
    if (!ain && !bin && !cin) {
        const bool iab = intersect_edge_ellipse(a, b, L, L).points > 0, iac = intersect_edge_ellipse(a, c, L, L).points > 0, ibc = intersect_edge_ellipse(b, c, L, L).points > 0;
        if (!iab && !iac && !ibc) return g2d::point_in_triangle2({ 0, 0 }, a, b, c) ? 1.f : 0.f;
    }
    const f_t min_len = min3(length2(a - b), length2(a - c), length2(b - c));
    if (min_len < 1e-3f) {
        const f_t delta = .002f;
        OT_DBG(0, 1); unsigned long long dbg_it = 0;
        if (b.y < a.y) std::swap(a, b);
        if (c.y < a.y) std::swap(a, c);
        const f_t ab = b.y == a.y ? inf : (b.x - a.x) / (b.y - a.y);
        const f_t ac = c.y == a.y ? inf : (c.x - a.x) / (c.y - a.y);
        const f_t bc = c.y == b.y ? inf : (c.x - b.x) / (c.y - b.y);
        f_t ret = 0;
        for (f_t y = std::max(-L, a.y + delta / 2); y < std::min(L, std::max(b.y, c.y)); y += delta) {
            f_t x0 = y < b.y ? ab * (y - a.y) + a.x : bc * (y - b.y) + b.x;
            f_t x1 = y < c.y ? ac * (y - a.y) + a.x : bc * (y - b.y) + b.x;
            if (x0 > x1) std::swap(x0, x1);
            for (f_t x = std::max(-L, x0) + delta / 2; x < std::min(L, x1); x += delta) { ret += lm::exp(-(sqr(x) + sqr(y)) / 2); OT_DBG(1, 1); ++dbg_it; }
        }
        OT_DBG_MAX(6, dbg_it); (void)dbg_it;
        return ret * inv_two_pi * sqr(delta);
    }
    OT_DBG(2, 1);
    // analytic approximation: T = mat2(b-a, c-a) (columns)
    const mat2 T{ b - a, c - a };
    const f_t detT = T.m[0][0] * T.m[1][1] - T.m[1][0] * T.m[0][1];
    // glm::inverse(mat2): 1/det * [[d,-b],[-c,a]] in column-major terms
    const f_t od = 1.f / detT;
    const mat2 Ti{ T.m[1][1] * od, -T.m[0][1] * od, -T.m[1][0] * od, T.m[0][0] * od };
    const v2 mu0 = Ti * a;
    mat2 Tt{ T.m[0][0], T.m[1][0], T.m[0][1], T.m[1][1] };
    const mat2 A = Tt * T;
    const f_t detA = A.m[0][0] * A.m[1][1] - A.m[1][0] * A.m[0][1];
    const f_t Sxy = A.m[0][1], Syy = A.m[1][1];
    if (Syy <= 0 || detA <= 0) return 0;
    const f_t denom = 1 / std::sqrt(2 * Syy);
    const f_t pa = detA * sqr(denom), pb = mu0.x;
    const f_t c0 = Sxy * denom, d0 = (Sxy * mu0.x + Syy * mu0.y) * denom, q = .5f / denom;
    const f_t I0 = g2d::I_gauss_erf(pa, pb, c0 - q, d0 + q), I1 = g2d::I_gauss_erf(pa, pb, c0, d0);
    return inv_sqrt_pi / 2 * std::fabs(detT * denom) * std::max(0.f, I0 - I1);
}

// gaussian_wavefront_t (include/wt/beam/gaussian_wavefront.hpp:22-119), built by beam_generic_t::wavefront (beam_generic.hpp:130-139)
struct wavefront_t {
    gaussian2d_t dist;
    explicit wavefront_t(const beam_t& b, f_t d) {
        const v3 fp = b.footprint(d);
        gaussian2d_t g(v2{ fp.x / beam_cross_section_envelope, fp.y / beam_cross_section_envelope });
        dist = g.is_dirac() ? gaussian2d_t(v2{ 0, 0 }) : g;
    }
    explicit wavefront_t(v2 sigma) { gaussian2d_t g(sigma); dist = g.is_dirac() ? gaussian2d_t(v2{ 0, 0 }) : g; }      // (for the pins: a wavefront of given standard deviations)
    v2 envelope() const { return dist.sigma * beam_cross_section_envelope; }
    f_t amplitude_magnitude(v2 x) const { return std::sqrt(dist.pdf(x)); }
    f_t integrate_triangle(v2 a, v2 b, v2 c) const { return dist.integrate_triangle(a, b, c); }
};

// include/wt/math/intersect/clip.hpp:20-88
struct clip_ret_t { v3 vs[5]; int tris = 0; void triangle(int idx, v3 o[3]) const {
    if (idx == 0) { o[0] = vs[0]; o[1] = vs[1]; o[2] = vs[2]; } else if (idx == 1) { o[0] = vs[2]; o[1] = vs[0]; o[2] = vs[tris == 2 ? 3 : 4]; } else { o[0] = vs[4]; o[1] = vs[2]; o[2] = vs[3]; } } };
inline clip_ret_t clip_triangle_z(v3 a, v3 b, v3 c, range_t zr) {
    const v3 ppmax{ 0, 0, zr.max }, ppmin{ 0, 0, zr.min }, n{ 0, 0, 1 };
    const v3 tri[3] = { a, b, c };
    int cls[3];
    for (int i = 0; i < 3; ++i) cls[i] = tri[i].z > zr.max ? +1 : tri[i].z < zr.min ? -1 : 0;
    clip_ret_t ret; int idx = 0;
    auto add = [&](v3 v) { if (idx < 5) ret.vs[idx++] = v; };
    for (int i = 0; i < 3; ++i) {
        const int next = i == 2 ? 0 : i + 1;
        if (cls[i] == 0) add(tri[i]);
        if (cls[next] != cls[i]) {
            const auto pt = intersect_edge_plane(tri[i], tri[next], cls[i] == -1 ? ppmin : (cls[i] == 1 || cls[next] == 1) ? ppmax : ppmin, n);
            add(pt ? *pt : (cls[i] != 0 ? tri[i] : tri[next]));
            if (cls[next] != 0 && cls[i] != 0) {
                const auto pt2 = intersect_edge_plane(tri[i], tri[next], cls[next] == 1 ? ppmax : ppmin, n);
                add(pt2 ? *pt2 : tri[next]);
            }
        }
    }
    ret.tris = idx < 3 ? 0 : idx == 3 ? 1 : idx == 4 ? 2 : 3;
    return ret;
}

// ================================================================================================ Fraunhofer FSD
namespace ffsd {
struct edge_t { v2 e, v; c_t a_b, iab_2; };
static constexpr f_t PA1 = 0.0049361075794549872500f, PA2 = 0.21899789398059305541f, P0_sigma = 0.288675134594813f / 4;
inline f_t alpha1(f_t x, f_t y) { return x == 0 ? 0.f : inv_two_pi * y / (x * (x * x + y * y)) * (lm::cos(x / 2) - sinc(x / 2)); }
inline f_t alpha2(f_t x, f_t y) { return x == 0 ? 0.f : inv_two_pi * y / (x * x + y * y) * sinc(x / 2); }
inline f_t chi_e(v2 xi) { const f_t chi = 0.830092714835359f; const f_t t = 1 + chi * dot(xi, xi), t2 = t * t, t3 = t2 * t; return std::max(0.f, 1 - (3 / t2 - 2 / t3)); }
inline f_t chi_0(v2 xi) { xi = xi / P0_sigma; return lm::exp(-.5f * dot(xi, xi)); }
// zeta = xi * Xi, Xi = mat2(e, m) columns, row-vector times matrix: (dot(xi,e), dot(xi,m)), m = (e.y,-e.x) (fsd.hpp:27-33, 108)
inline v2 zeta_of(const edge_t& e, v2 xi) { return { xi.x * e.e.x + xi.y * e.e.y, xi.x * e.e.y + xi.y * (-e.e.x) }; }
inline c_t Psi(const edge_t& e, v2 xi) {
    const v2 z = zeta_of(e, xi);
    const c_t a1 = e.a_b * alpha1(z.x, z.y), a2 = e.iab_2 * alpha2(z.x, z.y);
    return lm::polar(length2(e.e), -dot(e.v, xi)) * (a1 + a2);
}
inline f_t Psi2(const edge_t& e, v2 xi) {
    const v2 z = zeta_of(e, xi);
    const c_t a1 = e.a_b * alpha1(z.x, z.y), a2 = e.iab_2 * alpha2(z.x, z.y);
    return sqr(length2(e.e)) * std::norm(a1 + a2);
}
inline f_t Pj(const edge_t& e) { return sqr(length2(e.e)) * PA1 * std::norm(e.a_b) + sqr(length2(e.e)) * PA2 * std::norm(e.iab_2); }

struct aperture_t {
    std::vector<edge_t> edges; std::vector<f_t> edge_pdfs;
    f_t P0 = 0, P0_pdf = 0, psi02 = 0, recp_I = 0;
    f_t ASF_unclamped(v2 xi) const { c_t a{}; for (const auto& e : edges) a += Psi(e, xi); return std::norm(a); }
    f_t ASF(v2 xi) const { return ASF_unclamped(xi) * chi_e(xi) + psi02 * chi_0(xi); }
    f_t sampling_density(v2 xi) const { f_t d = 0; for (const auto& e : edges) d += Psi2(e, xi); return d * chi_e(xi) + P0 * inv_two_pi / sqr(P0_sigma) * chi_0(xi); }
};

// LUT sampling (fsd_lut.hpp:27-69): tables from the scene description (regenerated: the originals are LFS stubs)
struct lut_t {
    uint32_t N, M; const float *th1, *th2, *c1, *c2;
    static f_t lerp1(f_t x, const float* tbl, uint32_t S) {
        x *= (f_t)(S - 1);
        const size_t l = std::min((size_t)x, (size_t)S - 1), h = std::min(l + 1, (size_t)S - 1);
        const f_t f = x - std::floor(x);
        return f * tbl[h] + (1 - f) * tbl[l];
    }
    f_t lerp2(f_t x, f_t rx, const float* tbl) const {
        x *= (f_t)(M - 1);
        const size_t l = std::min((size_t)x, (size_t)M - 1), h = std::min(l + 1, (size_t)M - 1);
        const f_t f = x - std::floor(x);
        return f * lerp1(rx, tbl + h * M, M) + (1 - f) * lerp1(rx, tbl + l * M, M);
    }
    v2 sample(v3 r3, const float* th, const float* cd) const {
        const f_t theta = lerp1(r3.x, th, N);
        const f_t tf = theta * 2 / pi;
        const f_t r = std::max(0.f, lerp2(tf, r3.y, cd));
        v2 z = r * v2{ lm::cos(theta), lm::sin(theta) };
        const int q = std::min(3, (int)(r3.z * 4));
        z.x *= (((q + 1) / 2) % 2 == 0 ? 1.f : -1.f);
        z.y *= ((q / 2) % 2 == 0 ? 1.f : -1.f);
        return z;
    }
};
}

struct fraunhofer_fsd_t {       // fraunhofer::free_space_diffraction_t
    ffsd::aperture_t ap; f_t k; frame_t frame; const ffsd::lut_t* lut;
    static constexpr f_t wo2_cutoff = .85f;
    bool empty() const { return ap.edges.empty(); }
    struct for_test_t {};
    fraunhofer_fsd_t(for_test_t, const ffsd::lut_t* l) : k(1), frame(frame_t::canonical()), lut(l) {}      // aperture filled by the caller (oracle.cpp test hooks)

    // src/interaction/fsd/fraunhofer/free_space_diffraction.cpp:22-129 (fsd_unit = 1 mm; lengths arrive in metres)
    fraunhofer_fsd_t(const scene_t& sc, const ffsd::lut_t* l, const frame_t& fr, f_t k_, f_t total_power, const elliptic_cone_t& beam,
                     const std::vector<uint32_t>& edges, const wavefront_t& wf) : k(k_), frame(fr), lut(l) {
        const v2 cse = wf.envelope();
        const f_t r = std::max(cse.x, cse.y);
        const f_t max_edge_length = .33f * r;
        ap.recp_I = total_power > 0 ? 1 / total_power : 0;
        f_t P_total = 0;
        for (uint32_t ed : edges) {
            const wtgpu_edge& E = sc.d->edges[ed];
            const v3 n1{ E.n1[0], E.n1[1], E.n1[2] }, n2{ E.n2[0], E.n2[1], E.n2[2] };
            if (dot(beam.d(), n1) * dot(beam.d(), n2) >= 0) continue;
            const v3 l1 = frame.to_local(v3{ E.a[0], E.a[1], E.a[2] } - beam.o()), l2 = frame.to_local(v3{ E.b[0], E.b[1], E.b[2] } - beam.o());
            const v2 u1{ l1.x, l1.y }, u2{ l2.x, l2.y };
            f_t t1 = 0, t2 = 1;
            auto in_ell = [&](v2 p) { const v2 q{ p.x / cse.x, p.y / cse.y }; return dot(q, q) <= 1; };
            if (!in_ell(u1) || !in_ell(u2)) {
                const auto intr = intersect_edge_ellipse(u1, u2, cse.x, cse.y);
                if (intr.points == 0) continue;
                t1 = std::max(0.f, intr.t1); t2 = std::min(1.f, intr.t2);
            }
            auto mix2 = [](v2 a, v2 b, f_t t) { return v2{ mix(a.x, b.x, t), mix(a.y, b.y, t) }; };
            const f_t len = length(mix2(u1, u2, t1) - mix2(u1, u2, t2));
            const int segments = std::max(1, int(std::round(len / max_edge_length) + .5f));
            const f_t seg = 1.f / segments;
            v2 v1 = mix2(u1, u2, t1);
            f_t a = wf.amplitude_magnitude(v1);
            for (int i = 0; i < segments; ++i) {
                const f_t tt = mix(t1, t2, (f_t)(i + 1) * seg);
                const v2 v2p = mix2(u1, u2, tt);
                const f_t b = wf.amplitude_magnitude(v2p);
                if (a > 0 || b > 0) {
                    const v2 v = (v1 + v2p) / 2.f, e = v2p - v1;
                    // divide by fsd_unit (1 mm): metres -> mm
                    const ffsd::edge_t fe{ { e.x * 1000.f, e.y * 1000.f }, { v.x * 1000.f, v.y * 1000.f }, c_t{ a - b, 0 }, c_t{ 0, 1 } * c_t{ a + b, 0 } / 2.f };
                    const f_t pdf = ffsd::Pj(fe);
                    if (pdf > 0) { ap.edges.push_back(fe); ap.edge_pdfs.push_back(pdf); P_total += pdf; }
                }
                v1 = v2p; a = b;
            }
        }
        const f_t r0 = 3 * ffsd::P0_sigma;
        const v2 dirs[8] = { { -inv_sqrt_two, -inv_sqrt_two }, { -1, 0 }, { -inv_sqrt_two, inv_sqrt_two }, { 0, 1 }, { inv_sqrt_two, inv_sqrt_two }, { 1, 0 }, { inv_sqrt_two, -inv_sqrt_two }, { 0, -1 } };
        f_t acc = 0;
        for (int i = 0; i < 8; ++i) acc = acc + ap.ASF_unclamped(r0 * dirs[i]);
        ap.psi02 = acc / 8.f;
        ap.P0 = (two_pi * sqr(ffsd::P0_sigma) * ap.psi02) / sqr(k);      // k * fsd_unit is dimensionless k[1/mm]*1mm
        P_total += ap.P0;
        if (P_total > 0) { const f_t rp = 1.f / P_total; ap.P0_pdf = ap.P0 * rp; for (auto& p : ap.edge_pdfs) p *= rp; }
        else { ap.P0_pdf = 1; ap.edges.clear(); ap.edge_pdfs.clear(); }
    }

    // fsd_sampler.cpp:37-113
    v2 sampleN(sampler_t& s) const {
        // sampler.discrete<true>(n+1, pb) (sampler.hpp:52-71)
        const size_t count = ap.edges.size() + 1;
        const f_t p = s.r() * 1.f;
        f_t cdf = 0; size_t sel = count - 1;
        for (size_t i = 0; i + 1 < count; ++i) { const f_t ep = i == 0 ? ap.P0_pdf : ap.edge_pdfs[i - 1]; cdf += ep; if (p < cdf) { sel = i; break; } }
        if (sel == 0) return ffsd::P0_sigma * normal2d(s.r2());
        const ffsd::edge_t& e = ap.edges[sel - 1];
        // sample1: invXi = inverse(mat2(e, m)); zeta * invXi (row vector)
        const v2 m{ e.e.y, -e.e.x };
        const f_t det = e.e.x * m.y - m.x * e.e.y, od = 1.f / det;
        // glm::inverse column-major: inv = 1/det * mat2(d,-b,-c,a) with A=mat2(a,b,c,d) -> a=e.x,b=e.y,c=m.x,d=m.y
        const mat2 inv{ m.y * od, -e.e.y * od, -m.x * od, e.e.x * od };
        const f_t A = std::norm(e.a_b), B = std::norm(e.iab_2);
        // discrete<2>({A,B}) non-normalised
        const f_t P = A + B; const f_t pp = s.r() * P;
        const int tosample = pp < A ? 0 : 1;
        const v2 zeta = tosample == 0 ? lut->sample(s.r3(), lut->th1, lut->c1) : lut->sample(s.r3(), lut->th2, lut->c2);
        // row-vector * matrix: (dot(zeta, col0), dot(zeta, col1))
        return { zeta.x * inv.m[0][0] + zeta.y * inv.m[0][1], zeta.x * inv.m[1][0] + zeta.y * inv.m[1][1] };
    }
    struct xi_sample_t { v2 xi{}; f_t pdf = 0, weight = 0; };
    xi_sample_t sample_rejection(sampler_t& s) const {
        const size_t edge_count = ap.edges.size();
        const bool rej = edge_count > 1;
        const size_t M = edge_count, max_tries = M * 1024ul;
        const f_t recp_M = 1.f / (f_t)M;
        OT_DBG(3, 1); OT_DBG(5, M);
        for (size_t tr = 0; tr < max_tries; ++tr) {
            OT_DBG(4, 1);
            const v2 xi = sampleN(s);
            const f_t g = ap.sampling_density(xi), f = ap.ASF(xi);
            const bool done = rej ? s.r() * g < f * recp_M : true;
            if (done) return { xi, f * ap.recp_I, 1 };
        }
        return {};
    }
    struct sample_ret_t { v3 wo{ 0, 0, 1 }; f_t dpd = 0; f_t weight = 0; };
    // free_space_diffraction.hpp:68-93
    sample_ret_t sample(sampler_t& s) const {
        const auto smp = sample_rejection(s);
        const f_t scale = k;
        if (smp.pdf > 0) {
            const v2 zeta = smp.xi / scale;
            const v2 wl{ zeta.x / std::sqrt(1 + sqr(zeta.x)), zeta.y / std::sqrt(1 + sqr(zeta.y)) };
            const f_t wo2 = length2(wl);
            if (wo2 < wo2_cutoff) return { { wl.x, wl.y, std::sqrt(1 - wo2) }, smp.pdf, smp.weight };
        }
        return {};
    }
    // free_space_diffraction.hpp:99-115
    f_t pdf(v3 wl) const {
        const f_t wo2 = length2(v2{ wl.x, wl.y });
        if (wl.z <= 0 || wo2 >= wo2_cutoff) return 0;
        const v2 zeta{ wl.x / std::sqrt(1 - sqr(wl.x)), wl.y / std::sqrt(1 - sqr(wl.y)) };
        const v2 xi = k * zeta;
        const f_t p = ap.ASF(xi) * ap.recp_I;
        return (0 <= p && p < 1e+2f) ? p : 0.f;
    }
    f_t f(v3 wl) const { return pdf(wl); }
};

// ================================================================================================ plt_bdpt
inline f_t shading_normals_correction_scale(bool forward, f_t wig, f_t wog, f_t wis, f_t wos) {   // integrator/common.hpp:22-33
    if (forward) return std::min(std::fabs(wis * wog / (wos * wig)), 1e+2f);
    return 1;
}

struct bdpt_stats_t { uint64_t vertices = 0, connections = 0, splats = 0; ads_counters_t ads; };

struct plt_bdpt_t {
    const scene_t& sc; bsdf_eval_t bsdfs; emitters_t emitters; sensor_eval_t sensor; film_t& film; bdpt_stats_t* stats;
    ffsd::lut_t lut;
    uint32_t max_depth; bool RR, FSD, use_MIS, sensor_direct, emitter_direct, force_rt;

    plt_bdpt_t(const scene_t& s, film_t& f, bdpt_stats_t* st) : sc(s), bsdfs(s), emitters(s), sensor(s), film(f), stats(st) {
        const auto& it = s.d->integrator;
        max_depth = it.max_depth; RR = it.russian_roulette != 0; FSD = it.fsd != 0; use_MIS = it.mis != 0; sensor_direct = it.sensor_direct != 0; emitter_direct = it.emitter_direct != 0;
        force_rt = s.d->sensor.ray_trace_only != 0;
        lut = { s.d->fsd_lut_n, s.d->fsd_lut_m, s.d->fsd_icdf_theta1, s.d->fsd_icdf_theta2, s.d->fsd_icdf1, s.d->fsd_icdf2 };
    }
    ads_counters_t* ctr() const { return stats ? &stats->ads : nullptr; }

    enum vtype_e { V_SENSOR, V_EMITTER, V_SURFACE, V_FSD };
    struct vertex_t {           // vertex.hpp:49-81
        vtype_e type; bool forward = false; bool delta = false; bool fraunhofer_fsd = false;
        f_t pdf_fwd = -1, pdf_bwd = -1, rr_weight = 1;
        bool has_beam = false; beam_t beam;
        int32_t emitter = -1, bsdf = -1; int fsd = -1;
        geo_t geo;
        v3 wp() const { return geo.p; }
        f_t& pdf() { return forward ? pdf_fwd : pdf_bwd; }
        f_t& pdf_reversed() { return !forward ? pdf_fwd : pdf_bwd; }
    };
    struct arena_t { std::vector<vertex_t> sv, ev; std::vector<fraunhofer_fsd_t> fsds; };

    bool v_is_area_emitter(const vertex_t& v) const { return v.type == V_EMITTER && emitters.is_area(v.emitter); }
    bool is_on_surface(const vertex_t& v) const { return v.type == V_SURFACE || v_is_area_emitter(v) || (v.type == V_SENSOR && v.geo.kind == geo_t::SURFACE); }
    v3 v_ng(const vertex_t& v) const { return (v.type == V_SURFACE || v_is_area_emitter(v)) ? v.geo.s.geo.n : v3{ 0, 0, 1 }; }
    v3 v_ns(const vertex_t& v) const { return (v.type == V_SURFACE || v_is_area_emitter(v)) ? v.geo.s.shading.n : v3{ 0, 0, 1 }; }
    const surface_t* surface_if_any(const vertex_t& v) const { return (v.type == V_SURFACE || v_is_area_emitter(v)) ? &v.geo.s : nullptr; }
    int32_t v_emitter(const vertex_t& v) const { return v.type == V_EMITTER ? v.emitter : sc.d->shapes[sc.d->tri_meta[v.geo.s.tuid].shape_idx].emitter; }
    bool on_emitter(const vertex_t& v) const { return v.type == V_EMITTER || (v.type == V_SURFACE && sc.d->shapes[sc.d->tri_meta[v.geo.s.tuid].shape_idx].emitter >= 0); }
    bool is_delta_emitter(const vertex_t& v) const { return v.type == V_EMITTER && (emitters.is_delta_direction(v.emitter) || emitters.is_delta_position(v.emitter)); }
    bool is_delta_sensor(const vertex_t& v) const { return v.type == V_SENSOR && (sensor.is_delta_direction() || sensor.is_delta_position()); }
    bool is_nondelta_interaction(const vertex_t& v) const { return (v.type == V_SURFACE || v.type == V_FSD) && !v.delta; }
    bool is_connectible(const vertex_t& v) const {      // vertex.hpp:415-425
        switch (v.type) {
        case V_FSD: return true;
        case V_EMITTER: return !emitters.is_delta_direction(v.emitter);
        case V_SENSOR: return !sensor.is_delta_direction();
        case V_SURFACE: return !bsdfs.is_delta_only(v.bsdf, v.beam.k);
        }
        return false;
    }
    // vertex.hpp:224-243
    f_t dir_to_area(pd_t dpdf, v3 p, const vertex_t& next) const {
        if (dpdf.density_or_zero() == 0) return 0;
        const v3 d = next.wp() - p;
        const f_t d2 = length2(d);
        if (d2 == 0) return inf;
        f_t ppdf = dpdf.density_or_zero() * (1 / d2);
        if (is_on_surface(next)) ppdf *= std::fabs(dot(v_ng(next), normalize(d)));
        return ppdf;
    }
    // vertex.hpp:489-506 / 522-547 / 508-513 / 549-564
    f_t pdf_next_from_sensor(const vertex_t& v, const vertex_t& next) const {
        const v3 dl = next.wp() - v.wp();
        const f_t rd2 = 1 / length2(dl);
        const v3 d = dl * std::sqrt(rd2);
        f_t ppdf = sensor.pdf_direction_density(d) * rd2;
        if (is_on_surface(next)) ppdf *= std::fabs(dot(v_ng(next), d));
        return ppdf;
    }
    f_t pdf_sensor(const vertex_t&) const { return sensor.pdf_position_density(); }
    f_t pdf_next_from_emitter(const vertex_t& v, const vertex_t& next) const {
        const v3 dl = next.wp() - v.wp();
        const f_t rd2 = 1 / length2(dl);
        const v3 d = dl * std::sqrt(rd2);
        const int32_t em = v_emitter(v);
        if (emitters.is_infinite(em)) {     // directional_t::pdf_target_position (directional.hpp:172-179)
            const wtgpu_emitter& e = emitters.em(em);
            const frame_t fr = frame_t::build_orthogonal_frame({ e.dir[0], e.dir[1], e.dir[2] });
            const v3 pl = fr.to_local(next.wp() - v3{ e.world_centre[0], e.world_centre[1], e.world_centre[2] });
            return length2(v2{ pl.x, pl.y }) <= sqr(e.world_radius) ? 1.f / (pi * sqr(e.world_radius)) : 0.f;
        }
        f_t ppdf = emitters.pdf_direction_density(em, d, surface_if_any(v)) * rd2;
        if (is_on_surface(next)) ppdf *= std::fabs(dot(v_ng(next), d));
        return ppdf;
    }
    f_t pdf_emitter(const vertex_t& v) const {
        const int32_t em = v_emitter(v);
        if (emitters.is_infinite(em)) return 0;
        return emitters.pdf_emitter(em) * emitters.pdf_position_density(em);
    }
    // vertex.hpp:444-487
    f_t v_pdf(const arena_t& ar, const vertex_t& v, const vertex_t* prev, const vertex_t& next, bool mode_forward) const {
        if (v.type == V_EMITTER) return pdf_next_from_emitter(v, next);
        if (v.type == V_SENSOR) return pdf_next_from_sensor(v, next);
        const v3 p = v.wp();
        const v3 wiw = normalize(prev->wp() - p), wow = normalize(next.wp() - p);
        pd_t pdf = pd_t::discrete(0);
        if (v.type == V_SURFACE) {
            const surface_t& srf = v.geo.s;
            const bsdf_query_t q{ &srf, v.beam.k, mode_forward };
            pdf = pd_t::density(bsdfs.pdf(v.bsdf, srf.shading.to_local(wiw), srf.shading.to_local(wow), q));
        } else if (v.type == V_FSD) {
            if (v.fraunhofer_fsd) pdf = pd_t::density(ar.fsds[v.fsd].pdf(ar.fsds[v.fsd].frame.to_local(wow)));
            else return 0;
        }
        return dir_to_area(pdf, p, next);
    }
    // vertex_t::interact (vertex.hpp:330-413)
    std::optional<beam_t> interact(const arena_t& ar, const vertex_t& v, const vertex_t& next, bool ignore_fsd) const {
        const v3 wiw = -v.beam.dir();
        const f_t k = v.beam.k;
        f_t f = 0;
        if (v.fraunhofer_fsd && !ignore_fsd) {
            const v3 wow = normalize(next.wp() - v.wp());
            f = ar.fsds[v.fsd].f(ar.fsds[v.fsd].frame.to_local(wow));
        }
        if (v.type == V_SURFACE) {
            const surface_t& srf = v.geo.s;
            const v3 wow = normalize(next.wp() - v.wp());
            const v3 wi = srf.shading.to_local(wiw), wo = srf.shading.to_local(wow);
            const v3 ng = v_ng(v), ns = v_ns(v);
            const f_t wig = dot(wiw, ng), wog = dot(wow, ng), wis = wi.z, wos = wo.z;
            if (wig * wis <= 0 || wog * wos <= 0) return std::nullopt;
            mueller_t fb = bsdfs.f(v.bsdf, wi, wo, bsdf_query_t{ &srf, k, v.forward });
            f_t scale = 1 / std::fabs(wos);
            if (ns != ng) scale *= shading_normals_correction_scale(v.forward, wig, wog, wis, wos);
            fb = fb * scale;
            if (f > 0) fb = fb + mueller_t::identity() * f;
            if (fb.mean_intensity() == 0) return std::nullopt;
            beam_t ret = v.beam;
            ret.transform_surface_interaction(srf, wow, fb, 1);
            return ret;
        }
        if (v.type == V_FSD) {
            const v3 p = v.wp();
            const f_t beam_dist = dot(p - v.beam.origin(), v.beam.dir());
            beam_t ret = v.beam;
            ret.transform_region_interaction(p, beam_dist, normalize(next.wp() - p), f);
            return ret;
        }
        return std::nullopt;
    }

    struct walk_t {         // bdpt_walk_data_t (plt_bdpt_detail.hpp:70-183)
        beam_t beam; bool forward; pd_t pdf_from_prev; f_t throughput = 1, rr_weight = 1;
        std::vector<vertex_t>* vertices; arena_t* arena; sampler_t* sampler;
    };
    bool append_vertex(walk_t& d, vertex_t& v, pd_t pdf_fwd, pd_t pdf_revr) const {     // :96-122
        vertex_t& prev = d.vertices->back();
        if (prev.wp() == v.wp()) return false;
        v.pdf() = dir_to_area(d.pdf_from_prev, prev.wp(), v);
        v.beam = d.beam; v.has_beam = true;
        prev.pdf_reversed() = dir_to_area(pdf_revr, v.wp(), prev);
        d.pdf_from_prev = pdf_fwd;
        d.vertices->push_back(v);
        if (stats) stats->vertices++;
        return true;
    }
    bool continue_walk(walk_t& d, bool allow_RR) const {       // :167-182
        if (d.vertices->size() > max_depth + 1) return false;
        if (!allow_RR || !RR) return true;
        d.vertices->back().rr_weight = d.rr_weight;
        const f_t r = d.throughput < 1 ? std::max(d.throughput, .5f) : 1.f;
        if (d.sampler->r() <= r) { const f_t s = 1 / r; d.rr_weight *= s; d.throughput *= s; return true; }
        return false;
    }

    // find_closest_triangle (:362-419)
    struct wfi_t { uint32_t primary = WTGPU_INVALID_IDX; f_t dist = inf; v2 bary{ -1, -1 }; f_t integrated_radiant_flux = 0; };
    wfi_t find_closest_triangle(const std::vector<uint32_t>& tris, range_t zr, v3 origin, v3 dir, const frame_t& bf, const elliptic_cone_t& env,
                                const wavefront_t& wf, bool integrate_front_facing) const {
        wfi_t id;
        for (uint32_t t : tris) {
            const v3 a = sc.ads.tri_a(t), b = sc.ads.tri_b(t), c = sc.ads.tri_c(t);
            const f_t tol = cone_intersection_tolerance(origin, aabb_t::from_points(a, b, c));
            const auto intr = intersect_ray_tri(ray_t{ origin, dir }, a, b, c, zr.grow(tol));
            if (intr && intr->dist < id.dist) { id.primary = t; id.dist = intr->dist; id.bary = intr->bary; }
        }
        if (id.primary != WTGPU_INVALID_IDX) return id;
        for (uint32_t t : tris) {
            const bool front = dot(sc.ads.tri_n(t), -dir) > 0;
            if (front != integrate_front_facing) continue;
            const auto cl = clip_triangle_z(bf.to_local(sc.ads.tri_a(t) - env.o()), bf.to_local(sc.ads.tri_b(t) - env.o()), bf.to_local(sc.ads.tri_c(t) - env.o()), zr);
            const f_t csz = zr.centre();
            for (int i = 0; i < cl.tris; ++i) {
                v3 ct[3]; cl.triangle(i, ct);
                id.integrated_radiant_flux += wf.integrate_triangle(env.project_local(ct[0], csz), env.project_local(ct[1], csz), env.project_local(ct[2], csz));
            }
        }
        return id;
    }

    // random_walk (:421-526), iterative
    void random_walk(walk_t& data) const {
        for (;;) {
            beam_t& beam = data.beam;
            const auto intersection = traverse(sc, beam.envelope, data.vertices->back().geo, wavenum_to_wavelen(beam.k), force_rt, FSD, ctr());
            if (intersection.empty) return;
            const f_t beam_dist = intersection.distance();
            const range_t zr{ beam_dist, beam_dist + intersection.intersection_region_depth };
            const bool is_ballistic = intersection.ballistic || beam.is_ray();
            const v3 origin_wp = intersection.origin;
            const v3 interaction_wp = origin_wp + zr.min * beam.dir();
            const frame_t beam_frame = beam.envelope.frame();
            const elliptic_cone_t envelope = beam.envelope;
            const wavefront_t wf(beam, beam_dist);
            wfi_t ct;
            if (is_ballistic) { ct.primary = intersection.ray.tuid; ct.dist = intersection.ray.dist; ct.bary = intersection.ray.bary; }
            else ct = find_closest_triangle(intersection.cone.tris, zr, origin_wp, beam.dir(), beam_frame, envelope, wf, intersection.cone.front_face);
            bool do_RR = true;
            if (ct.primary != WTGPU_INVALID_IDX) {
                // sample_surface_interaction (:193-270)
                const f_t k = beam.k;
                surface_t srf = sc.make_surface(ct.primary, ct.bary, origin_wp + envelope.d() * ct.dist);
                srf.footprint = beam.surface_footprint_static(srf, beam_dist);
                const int32_t bsdf = sc.d->shapes[sc.d->tri_meta[ct.primary].shape_idx].bsdf;
                const v3 ng = srf.ng(), ns = srf.ns();
                const v3 wiw = -beam.dir();
                const v3 wi = srf.shading.to_local(wiw);
                const f_t wig = dot(wiw, ng), wis = wi.z;
                if (wig * wis <= 0) return;
                const auto bs = bsdfs.sample(bsdf, wi, bsdf_query_t{ &srf, k, data.forward }, *data.sampler);
                if (!bs || bs->dpd.is_zero()) return;
                const bool is_delta = bs->dpd.is_discrete;
                const v3 wo = bs->wo;
                const v3 wow = normalize(srf.shading.to_world(wo));
                const f_t wog = dot(wow, ng), wos = wo.z;
                if (wog * wos <= 0) return;
                const pd_t pdf_fwd = bs->dpd;
                const pd_t pdf_revr = pd_t::density(bsdfs.pdf(bsdf, wo, wi, bsdf_query_t{ &srf, k, !data.forward }));
                vertex_t v; v.type = V_SURFACE; v.forward = data.forward; v.delta = is_delta; v.bsdf = bsdf; v.geo = geo_t::surface(srf);
                if (!append_vertex(data, v, pdf_fwd, pdf_revr)) return;
                f_t w = 1;
                if (ns != ng) w *= shading_normals_correction_scale(data.forward, wig, wog, wis, wos);
                // transform_surface_interaction (:125-137)
                data.beam.transform_surface_interaction(srf, wow, bs->M, w);
                data.throughput *= w * bs->M.mean_intensity();
                const f_t eta = bs->eta.real();
                if (!data.forward && eta != 1) data.throughput /= sqr(eta);
            } else if (!is_ballistic && !intersection.cone.edges.empty()) {
                // sample_fraunhofer_fsd_interaction (:288-346)
                const f_t I = 1 - ct.integrated_radiant_flux;
                fraunhofer_fsd_t fs(sc, &lut, beam.envelope.frame(), beam.k, I, beam.envelope, intersection.cone.edges, wf);
                if (fs.empty()) { data.beam.transform_restart(interaction_wp, beam_dist); do_RR = false; }
                else {
                    data.arena->fsds.push_back(fs);
                    const int fi = (int)data.arena->fsds.size() - 1;
                    const auto smp = data.arena->fsds[fi].sample(*data.sampler);
                    if (smp.dpd == 0 || smp.weight == 0) return;
                    const v3 wow = fs.frame.to_world(smp.wo);
                    vertex_t v; v.type = V_FSD; v.forward = data.forward; v.delta = false; v.fraunhofer_fsd = true; v.fsd = fi; v.geo = geo_t::point(interaction_wp);
                    if (!append_vertex(data, v, pd_t::density(smp.dpd), pd_t::density(smp.dpd))) return;
                    data.beam.transform_region_interaction(interaction_wp, beam_dist, wow, smp.weight);
                    data.throughput *= smp.weight;
                }
            } else {
                do_RR = false;
                data.beam.transform_restart(interaction_wp, beam_dist);
            }
            if (!continue_walk(data, do_RR)) return;
        }
    }

    struct connect_ret_t { vertex_t tmp; bool has_element = false; element_sample_t element; stokes_t L{}; };

    stokes_t connect_and_integrate(const beam_t& db, const geo_t& dg, const beam_t& eb, const geo_t& eg) const {     // :725-745
        if (db.intensity() == 0 || eb.intensity() == 0) return {};
        if (shadow(sc, dg, eg, ctr())) return {};
        return integrate_beams(db, eb);
    }

    // connect_subpaths (:747-923)
    connect_ret_t connect_subpaths(const arena_t& ar, int s, int t, sampler_t& sampler) const {
        const auto& sv = ar.sv; const auto& ev = ar.ev;
        connect_ret_t ret;
        ret.tmp.type = V_SENSOR;
        if (s == 0) {
            const vertex_t& last = sv[t - 1];
            if (on_emitter(last)) {
                beam_t QE = last.beam; QE.mul(last.rr_weight);
                ret.L = emitters.Li(v_emitter(last), QE, is_on_surface(last) ? &last.geo.s : nullptr);
            }
        } else if (t == 0) {
            const vertex_t& last = ev[s - 1];
            if (sensor.is_virtual()) {
                const vertex_t& current = ev[s - 2];
                const beam_t& beam = last.beam;
                const f_t dist = length(last.wp() - beam.origin());
                auto dc = sensor.Si(beam, { 0, dist });
                if (dc) {
                    ret.has_element = true; ret.element = dc->element;
                    beam_t& db = dc->beam;
                    f_t w = current.rr_weight;
                    if (is_on_surface(current) && is_nondelta_interaction(current)) w /= std::fabs(dot(db.dir(), v_ns(current)));
                    if (dc->surface) w /= std::fabs(dot(db.dir(), dc->surface->ng()));
                    db.mul(w);
                    ret.tmp = vertex_t{}; ret.tmp.type = V_SENSOR; ret.tmp.forward = false; ret.tmp.pdf_bwd = 0;
                    ret.tmp.geo = dc->surface ? geo_t::surface(*dc->surface) : geo_t::point(db.origin());
                    ret.L = integrate_beams(db, beam);
                }
            }
        } else if (s == 1) {
            const vertex_t& last = sv[t - 1];
            if (is_connectible(last)) {
                auto ed = emitters.sample_emitter_direct(sampler, last.wp(), last.beam.k);
                if ((ed.dpd.is_discrete || !ed.dpd.is_zero()) && ed.beam.intensity() > 0) {
                    f_t w = last.rr_weight;
                    if (is_on_surface(last)) w *= std::fabs(dot(ed.beam.dir(), v_ns(last)));
                    ed.beam.mul(w);
                    ret.tmp = vertex_t{}; ret.tmp.type = V_EMITTER; ret.tmp.forward = true; ret.tmp.emitter = ed.emitter;
                    ret.tmp.geo = ed.surface ? geo_t::surface(*ed.surface) : geo_t::point(ed.beam.origin());
                    const auto db = interact(ar, last, ret.tmp, false);
                    if (db) ret.L = connect_and_integrate(*db, last.geo, ed.beam, ret.tmp.geo);
                }
            }
        } else if (t == 1) {
            const vertex_t& last = ev[s - 1];
            const bool do_direct = (sensor.is_virtual() || last.type != V_FSD) && is_connectible(last);
            if (do_direct) {
                auto sd = sensor.sample_direct(sampler, last.wp(), last.beam.k);
                if ((sd.dpd.is_discrete || !sd.dpd.is_zero()) && sd.beam.intensity() > 0) {
                    f_t w = last.rr_weight;
                    if (is_on_surface(last)) w *= std::fabs(dot(sd.beam.dir(), v_ns(last)));
                    sd.beam.mul(w);
                    ret.tmp = vertex_t{}; ret.tmp.type = V_SENSOR; ret.tmp.forward = false;
                    ret.tmp.geo = sd.surface ? geo_t::surface(*sd.surface) : geo_t::point(sd.beam.origin());
                    const auto eb = interact(ar, last, ret.tmp, false);
                    if (eb) { ret.L = connect_and_integrate(sd.beam, ret.tmp.geo, *eb, last.geo); ret.has_element = true; ret.element = sd.element; }
                }
            }
        } else {
            const vertex_t& evx = ev[s - 1]; const vertex_t& svx = sv[t - 1];
            const v3 dl = evx.wp() - svx.wp();
            if (is_connectible(evx) && is_connectible(svx) && !(dl.x == 0 && dl.y == 0 && dl.z == 0)) {
                const auto eb = interact(ar, evx, svx, true);
                const auto db = interact(ar, svx, evx, true);
                if (eb && db) {
                    const f_t recp_d2 = 1 / length2(dl);
                    const v3 d = dl * std::sqrt(recp_d2);
                    f_t wev = evx.rr_weight, wsv = svx.rr_weight * recp_d2;
                    if (is_on_surface(svx)) wev *= std::fabs(dot(v_ns(svx), d));
                    if (is_on_surface(evx)) wsv *= std::fabs(dot(v_ns(evx), d));
                    beam_t dbw = *db; dbw.mul(wsv);
                    beam_t ebw = *eb; ebw.mul(wev);
                    ret.L = connect_and_integrate(dbw, svx.geo, ebw, evx.geo);
                }
            }
        }
        return ret;
    }

    // bdpt_compute_mis_weight (:604-720)
    f_t mis_weight(const arena_t& ar, int s, int t, const connect_ret_t& cr) const {
        if (s + t <= 2) return 1;
        struct P { f_t pdf, pdf_rev; bool delta; };
        std::vector<P> sp(t), ep(s);
        for (int i = 0; i < t; ++i) sp[i] = { ar.sv[i].pdf_bwd, ar.sv[i].pdf_fwd, ar.sv[i].delta };
        for (int i = 0; i < s; ++i) ep[i] = { ar.ev[i].pdf_fwd, ar.ev[i].pdf_bwd, ar.ev[i].delta };
        const vertex_t& tmp = cr.tmp;
        if (s == 0) {
            const vertex_t& last = ar.sv[t - 1]; const vertex_t& prev = ar.sv[t - 2];
            sp[t - 1].pdf_rev = pdf_emitter(last);
            sp[t - 2].pdf_rev = pdf_next_from_emitter(last, prev);
        } else if (t == 0) {
            const vertex_t& last = sensor.is_virtual() ? tmp : ar.ev[s - 1]; const vertex_t& prev = ar.ev[s - 2];
            ep[s - 1].pdf_rev = pdf_sensor(last);
            ep[s - 2].pdf_rev = pdf_next_from_sensor(last, prev);
        } else if (s == 1) {
            const vertex_t& last = ar.sv[t - 1];
            sp[t - 1].pdf_rev = pdf_next_from_emitter(tmp, last);
            ep[0].pdf_rev = v_pdf(ar, last, &ar.sv[t - 2], tmp, false);
            ep[0].pdf = pdf_emitter(tmp);
        } else if (t == 1) {
            const vertex_t& last = ar.ev[s - 1];
            ep[s - 1].pdf_rev = pdf_next_from_sensor(tmp, last);
            sp[0].pdf_rev = v_pdf(ar, last, &ar.ev[s - 2], tmp, true);
            sp[0].pdf = pdf_sensor(tmp);
        } else {
            const vertex_t& e = ar.ev[s - 1]; const vertex_t& sv_ = ar.sv[t - 1]; const vertex_t& ep_ = ar.ev[s - 2]; const vertex_t& sp_ = ar.sv[t - 2];
            ep[s - 1].pdf_rev = v_pdf(ar, sv_, &sp_, e, false);
            ep[s - 2].pdf_rev = v_pdf(ar, e, &sv_, ep_, false);
            sp[t - 1].pdf_rev = v_pdf(ar, e, &ep_, sv_, true);
            sp[t - 2].pdf_rev = v_pdf(ar, sv_, &e, sp_, true);
        }
        if (t > 0) sp[t - 1].delta = false;
        if (s > 0) ep[s - 1].delta = false;
        const bool delta_emitter = s == 1 ? is_delta_emitter(tmp) : s > 1 ? is_delta_emitter(ar.ev[0]) : true;
        const bool delta_sensor = t == 1 ? is_delta_sensor(tmp) : t > 1 ? is_delta_sensor(ar.sv[0]) : true;
        auto one = [](f_t p) { return std::isfinite(p) && p > std::numeric_limits<f_t>::epsilon() ? p : 1.f; };
        f_t sum = 0, ri = 1;
        for (int i = t - 1; i >= 0; --i) { ri *= one(sp[i].pdf_rev) / one(sp[i].pdf); if (!sp[i].delta && !(i > 0 ? sp[i - 1].delta : delta_sensor)) sum += ri; }
        ri = 1;
        for (int i = s - 1; i >= 0; --i) { ri *= one(ep[i].pdf_rev) / one(ep[i].pdf); if (!ep[i].delta && !(i > 0 ? ep[i - 1].delta : delta_emitter)) sum += ri; }
        return 1 / (1 + sum);
    }

    // plt_bdpt_t::integrate, one sample (src/integrator/plt_bdpt.cpp:54-147)
    void integrate(uint32_t ex, uint32_t ey, sampler_t& sampler) const {
        const int32_t em = emitters.sample_emitter(sampler);
        const f_t emitter_pdf = emitters.pdf_emitter(em);
        const auto ws = emitters.sample_wavenumber(em, sampler);
        const f_t k = ws.k;
        const auto es = emitters.sample(em, sampler, k);
        const f_t recp_spectral_pd = ws.wpd.is_discrete ? 1.f / ws.wpd.v : 1.f / emitters.sum_spectral_pdf_for_all_emitters(k);
        const auto ss = sensor.sample(sampler, ex, ey, k);
        arena_t ar;
        {   // generate_sensor_subpath (:528-555)
            vertex_t v; v.type = V_SENSOR; v.forward = false; v.pdf_bwd = ss.ppd.density_or_zero(); v.beam = ss.beam; v.has_beam = true;
            v.geo = ss.surface ? geo_t::surface(*ss.surface) : geo_t::point(ss.beam.origin());
            ar.sv.push_back(v);
            sampler.set_stream(1);
            walk_t d{ ss.beam, false, ss.dpd, 1, 1, &ar.sv, &ar, &sampler };
            random_walk(d);
        }
        {   // generate_emitter_subpath (:557-581)
            vertex_t v; v.type = V_EMITTER; v.forward = true; v.pdf_fwd = es.ppd.density_or_zero() * emitter_pdf; v.beam = es.beam; v.has_beam = true; v.emitter = em;
            v.geo = es.surface ? geo_t::surface(*es.surface) : geo_t::point(es.beam.origin());
            ar.ev.push_back(v);
            sampler.set_stream(2);
            walk_t d{ es.beam, true, es.dpd, 1, 1, &ar.ev, &ar, &sampler };
            random_walk(d);
        }
        stokes_t L{};
        const f_t k_density = ws.wpd.v;     // mass per (1/mm) or density per (1/mm)
        for (int t = 0; t <= (int)ar.sv.size(); ++t)
            for (int s = 0; s <= (int)ar.ev.size(); ++s) {
                const int depth = t + s - 2;
                if ((t == 1 && s == 1) || depth < 0) continue;
                if (!emitter_direct && s == 1) continue;
                if (!sensor_direct && t == 1) continue;
                if (depth > (int)max_depth) break;
                sampler.set_stream(3u + 4096u * (uint32_t)t + (uint32_t)s);
                const auto ret = connect_subpaths(ar, s, t, sampler);
                if (stats) stats->connections++;
                if (ret.L.intensity() <= 0) continue;
                const f_t mis = use_MIS ? mis_weight(ar, s, t, ret) * recp_spectral_pd : 1 / ((f_t)(s + t + 1) * k_density);
                const stokes_t flux = ret.L * mis;
                if (t > 1) L = L + flux;
                else { film.splat_direct(ret.element, flux, k); if (stats) stats->splats++; }
            }
        film.splat(ss.element, L, k);
        if (stats) stats->splats++;
    }
};

} // namespace ot
