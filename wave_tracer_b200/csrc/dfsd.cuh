// dfsd.cuh -- UTD free-space diffraction on the device.
// Follows src/interaction/fsd/free_space_diffraction.cpp:23-234 and include/wt/interaction/fsd/utd.hpp:26-172 (under /root/reference).
// The aperture is stored as (interaction geometry, surviving edge ids); wedge parameters are rebuilt from the
// 96-B ads edge record on use instead of keeping a heap-allocated vector<wedge_edge_t> per path.
#pragma once
#include "dscene.cuh"

namespace wt {

constexpr float kUtdMinSinBeta = 1e-3f;
constexpr float kUtdSigmaScale = 45.f;
// The aperture of a UTD vertex: interaction geometry + the ids of the edges that take part.  The edge list is a row of an HBM array
// (sc.cap.edges entries per path, two rows per path: the aperture a path carries from its previous vertex and the one it builds now).
struct ApHead { V3 wp; Frame fr; V3 size; V3 wi; float k; };        // an Aperture without its edge list
struct Aperture {
    V3 wp; Frame fr; V3 size; V3 wi; float k;
    uint32_t n; uint32_t* edges;
};
struct Wedge { V3 v; float l; V3 nff, tff, nbf, e; float alpha; uint32_t idx; };

// free_space_diffraction_t ctor body for one edge (free_space_diffraction.cpp:36-78); false: edge does not take part
template <class AP> WT_D bool wedge_build(const DScene& sc, const AP& ap, uint32_t ed, Wedge& w) {
    const wtgpu_edge E = sc.edges[ed];
    const V3 n1 = mk3(E.n1), n2 = mk3(E.n2), t1 = mk3(E.t1), t2 = mk3(E.t2), ea = mk3(E.a), eb = mk3(E.b);
    const bool f1 = dot(ap.wi, n1) > 0.f;
    w.nff = f1 ? n1 : n2; w.tff = f1 ? t1 : t2; w.nbf = f1 ? n2 : n1;
    if (dot(ap.wi, w.nff) <= 0.f) return false;
    V3 v1 = ea, v2 = eb;
    if (vfinite(ap.size)) {
        float a, b;
        intersect_edge_ellipsoid(ea, eb, ap.wp, ap.fr.t, ap.fr.b, ap.size, a, b);
        a = clampf_(a, 0.f, 1.f); b = clampf_(b, 0.f, 1.f);
        v1 = mk3(mixf(ea.x, eb.x, a), mixf(ea.y, eb.y, a), mixf(ea.z, eb.z, a));
        v2 = mk3(mixf(ea.x, eb.x, b), mixf(ea.y, eb.y, b), mixf(ea.z, eb.z, b));
    }
    if (veq(v1, v2)) return false;
    w.v = (v1 + v2) / 2.f; w.l = length(v2 - v1); w.alpha = E.alpha; w.idx = ed;
    w.e = cross(w.nff, w.tff);
    return true;
}

// complex erfc(z) for |z| < 2.5: Maclaurin series of erf in binary64 (libcerf, the reference's cerfc, is a missing submodule).  The only call
// of the path is UTDF's cerfc(exp(i pi/4) * sqrt(x)) (utd.hpp:42), whose argument the reference forms in complex<float>: the point is off the
// 45-degree ray by f32 rounding and the series is evaluated AT THAT POINT, term for term as the CPU checker of the test suite does, which is pinned
// bit for bit against the reference's own utd.hpp.  A handful of calls per diffracting edge; never on the BVH-bound part.
WT_D void cerfc_series(float zre, float zim, float& re, float& im) {
    const double zr = (double)zre, zi = (double)zim;
    const double z2r = zr * zr - zi * zi, z2i = zr * zi + zi * zr;
    double tr = zr, ti = zi, sr = zr, si = zi;      // term = (-1)^n z^(2n+1) / n!
    for (int n = 1; n < 200; ++n) {
        const double fr = -z2r / (double)n, fi = -z2i / (double)n;
        const double nr = tr * fr - ti * fi, ni = tr * fi + ti * fr;
        tr = nr; ti = ni;
        const double d = 2.0 * n + 1.0;
        sr += tr / d; si += ti / d;
        if (tr * tr + ti * ti < 1e-60 && n > 4) break;
    }
    const double q = 2.0 / 1.7724538509055160273;
    re = (float)(-(q * sr) + 1.0); im = (float)(-(q * si));
}
WT_D C2 UTDF(float x) {                                         // utd.hpp:36-57
    const float ax = fabsf(x);
    C2 res;
    if (ax < 6.f) {
        const float sx = sqrtf(ax);
        float e4s, e4c; pm::sincosf(kPi4, &e4s, &e4c);      // std::exp(c_t{0, pi/4}) in complex<float>
        float cr, ci; cerfc_series(e4c * sx, e4s * sx, cr, ci);
        res = ((mkc(1.f, 1.f) * kSqrtPi2) * sx) * cexpi(ax) * mkc(cr, ci);
    } else {
        const float r = 1.f / (2.f * ax);
        const float r2 = r * r, r3 = r2 * r, r4 = r2 * r2;
        res = mkc(1.f - 3.f * r2 + 75.f * r4, r - 15.f * r3);
    }
    return x < 0.f ? cconj(res) : res;
}
WT_D float UTDa(float sgn, float phi, float n) {                // utd.hpp:26-31
    const float N = roundf((sgn * kPi + phi) * kInvTwoPi / n);
    return 2.f * sqrf(pm::cosf(kPi * n * N - phi / 2.f));
}
WT_D float fmod_pos(float a, float b) { return a - b * floorf(a / b); }
WT_D float cotf_(float x) { return 1.f / pm::tanf(x); }

WT_D bool wedge_diffraction_point(const Wedge& w, V3 src, V3 dst, V3& p) {     // utd.hpp:62-80
    const float sl = length(mk2(dot(src - w.v, w.tff), dot(src - w.v, w.nff)));
    const float dl = length(mk2(dot(dst - w.v, w.tff), dot(dst - w.v, w.nff)));
    const float dist = dot(w.e, src - w.v) + dot(dst - src, w.e) * sl / (sl + dl);
    if (fabsf(dist) > w.l / 2.f) return false;
    p = w.v + w.e * dist;
    return !(veq(p, src) || veq(p, dst));
}
WT_D bool wedge_diffraction_point_dir(const Wedge& w, V3 src, V3 wo, V3& p) {  // utd.hpp:85-110
    const float cb = dot(wo, w.e);
    const float sb = sqrtf(fmaxf(0.f, 1.f - sqrf(cb)));
    if (sb < kUtdMinSinBeta) return false;
    const float sl = length(mk2(dot(src - w.v, w.tff), dot(src - w.v, w.nff)));
    const V3 prj = w.v + dot(src - w.v, w.e) * w.e;
    p = prj + sl * (cb / sb) * w.e;
    if (length2(p - w.v) > sqrf(w.l / 2.f)) return false;
    return !veq(p, src);
}
WT_DN void wedge_UTD(const Wedge& w, float k, V3 wi, V3 wo, float ro, C2& Ds, C2& Dh) {   // utd.hpp:115-172
    const float n = 2.f - w.alpha * kInvPi;
    const float sb2 = fmaxf(0.f, 1.f - sqrf(dot(wi, w.e)));
    const float sb = sqrtf(sb2);
    const float phii = pm::atan2f(dot(w.nff, wi), dot(w.tff, wi));
    const float phio = pm::atan2f(dot(w.nff, wo), dot(w.tff, wo));
    const float kL = k_times_len(k, ro * sb2);
    const C2 F1 = UTDF(kL * UTDa(1.f, phii - phio, n)), F2 = UTDF(kL * UTDa(-1.f, phii - phio, n));
    const C2 F3 = UTDF(kL * UTDa(1.f, phii + phio, n)), F4 = UTDF(kL * UTDa(-1.f, phii + phio, n));
    const C2 D1 = (-cotf_((kPi + (phii - phio)) / (2.f * n))) * F1;
    const C2 D2 = (-cotf_((kPi - (phii - phio)) / (2.f * n))) * F2;
    const C2 D3 = (-cotf_((kPi + (phii + phio)) / (2.f * n))) * F3;
    const C2 D4 = (-cotf_((kPi - (phii + phio)) / (2.f * n))) * F4;
    const float kro = k_times_len(k, ro);
    const C2 D = (1.f / (2.f * n * sqrtf(kro) * sb) * kInvSqrtTwoPi) * cexpi(-kPi4);
    const float t1 = fmod_pos(phii + phio, kPi2), t2 = fmod_pos(phii - phio, kPi2);
    const bool z = fabsf(t1) < 1e-5f || fabsf(t2) < 1e-5f;
    const C2 s = z ? mkc(0.f, 0.f) : (D1 + D2) - (D3 + D4);
    const C2 h = z ? mkc(0.f, 0.f) : (D1 + D2) + (D3 + D4);
    Ds = (-D) * s; Dh = (-D) * h;
}

// free_space_diffraction_t::pdf (free_space_diffraction.cpp:152-194)
WT_DN float fsd_pdf(const DScene& sc, const Aperture& ap, V3 src, V3 wo) {
    if (ap.n == 0u) return 0.f;
    float ret = 0.f;
    for (uint32_t j = 0; j < ap.n; ++j) {
        Wedge w; if (!wedge_build(sc, ap, ap.edges[j], w)) continue;
        V3 p; if (!wedge_diffraction_point_dir(w, src, wo, p)) continue;
        const V3 ui = src - p;
        if ((dot(wo, w.nff) <= 0.f && dot(wo, w.nbf) <= 0.f) || (dot(ui, w.nff) <= 0.f && dot(ui, w.nbf) <= 0.f)) continue;
        const float ri = length(src - p);
        const V3 wi = (src - p) / ri;
        const float phii = pm::atan2f(dot(w.nff, wi), dot(w.tff, wi)), phio = pm::atan2f(dot(w.nff, wo), dot(w.tff, wo));
        const float sigma = sqrtf(kUtdSigmaScale / k_times_len(ap.k, ri));
        float x1 = fabsf(fmod_pos(phio - (kPi + phii), kTwoPi)), x2 = fabsf(fmod_pos(phio - (kPi - phii), kTwoPi));
        if (x1 > kPi) x1 -= kTwoPi;
        if (x2 > kPi) x2 -= kTwoPi;
        ret += kInvSqrtTwoPi / sigma * (pm::expf(-.5f * sqrf(x1 / sigma)) + pm::expf(-.5f * sqrf(x2 / sigma))) / 2.f;
    }
    return ret / (float)(ap.n + 1u);
}
// free_space_diffraction_t::sample (free_space_diffraction.cpp:84-150); invalid samples carry weight 0
WT_DN void fsd_sample(const DScene& sc, const Aperture& ap, V3 src, Sampler& smp, V3& wo, float& weight) {
    wo = mk3(0.f, 0.f, 1.f); weight = 0.f;
    const int eidx = uniform_int_interval(smp, 0, (int)ap.n + 1);
    if (eidx == (int)ap.n) { wo = -normalize(src - ap.wp); weight = (float)(ap.n + 1u); return; }
    Wedge w; if (!wedge_build(sc, ap, ap.edges[eidx], w)) return;
    const V3 p = w.v + (rnd(smp) - .5f) * w.l * w.e;
    const V3 ui = src - p;
    if (dot(ui, w.nff) <= 0.f && dot(ui, w.nbf) <= 0.f) return;
    const float ri = length(src - p);
    const V3 wi = (src - p) / ri;
    const float phii = pm::atan2f(dot(w.nff, wi), dot(w.tff, wi));
    const float sigma = sqrtf(kUtdSigmaScale / k_times_len(ap.k, ri));
    const float sm = sigma * normal2d(rnd2(smp)).x;
    const float phio = (rnd(smp) < .5f ? kPi + phii : kPi - phii) + sm;
    const float cb = dot(wi, w.e);
    const float sb = sqrtf(fmaxf(0.f, 1.f - sqrf(cb)));
    const V3 d = sb * (pm::cosf(phio) * w.tff + pm::sinf(phio) * w.nff) - cb * w.e;
    if (dot(d, w.nff) <= 0.f && dot(d, w.nbf) <= 0.f) return;
    if (sb < kUtdMinSinBeta) return;
    const float dpd = fsd_pdf(sc, ap, src, d);
    if (dpd == 0.f) return;
    wo = d; weight = 1.f / dpd;
}

// plt_path do_fsd (integrator/plt_path/plt_path_detail.hpp:311-346): returns (|ts|^2+|th|^2)/2
WT_DN float do_fsd(const DScene& sc, const Cone& cone_from_src, const Geo& src_geo, V3 dst, const Aperture& ap, float k, Counters& ctr, uint32_t& edges_fetched) {
    const V3 src = cone_from_src.o;
    const Geo dst_geo = geo_point(dst);
    C2 ts = mkc(0.f, 0.f), th = mkc(0.f, 0.f);
    for (uint32_t j = 0; j < ap.n; ++j) {
        Wedge w; if (!wedge_build(sc, ap, ap.edges[j], w)) continue;
        ++edges_fetched;
        V3 p; if (!wedge_diffraction_point(w, src, dst, p)) continue;
        const V3 ui = src - p, uo = dst - p;
        if ((dot(uo, w.nff) <= 0.f && dot(uo, w.nbf) <= 0.f) || (dot(ui, w.nff) <= 0.f && dot(ui, w.nbf) <= 0.f)) continue;
        const float ri = length(ui), ro = length(uo);
        C2 Ds, Dh; wedge_UTD(w, ap.k, ui / ri, uo / ro, ro, Ds, Dh);
        if (Dh.re == 0.f && Dh.im == 0.f && Ds.re == 0.f && Ds.im == 0.f) continue;
        const Geo eg = geo_edge(p, w.idx);
        if (shadow_between(sc, eg, src_geo, ctr) || shadow_between(sc, eg, dst_geo, ctr)) continue;
        const C2 phase = cexpi(-k_times_len(k, ro + ri));
        ts = ts + phase * Ds; th = th + phase * Dh;
    }
    if (cone_contains(cone_from_src, dst)) {
        if (!shadow_between(sc, src_geo, dst_geo, ctr)) {
            const C2 phase = cexpi(-k_times_len(k, length(dst - src)));
            ts = ts + phase; th = th + phase;
        }
    }
    return (cnorm(ts) + cnorm(th)) / 2.f;
}

// ------------------------------------------------------------------------------------------------ do_fsd, one warp for its 32 paths
// do_fsd per thread runs a loop over <= 48 edges in which most edges fail a cheap geometric test and a few reach the expensive part
// (four UTD transition functions + two shadow rays): the warp ends up executing the expensive part one or two lanes at a time
// (ncu: 1.2 active threads per instruction on the etoile-like scene).  Here the warp's 32 paths are served together:
//   filter   every lane walks ITS path's edges through the cheap tests (wedge_build, diffraction point, side tests) in lockstep and appends
//            the survivors to a shared (owner lane, edge) list;
//   evaluate whenever 32 items are queued, every lane evaluates one item -- UTD coefficients, both shadow rays, phase -- with the owner's
//            inputs fetched by shuffles;
//   reduce   every owner adds its items' terms in list order, which is its own edge order: sums are bit-identical to the sequential loop.
// All 32 lanes must call (need = false for lanes without a request).
struct UtdShared { uint32_t item[64]; float4 res[32]; uint32_t ok[32]; };
WT_D V3 shfl3(V3 v, int src) { return mk3(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src), __shfl_sync(0xffffffffu, v.z, src)); }
__device__ __noinline__ float warp_do_fsd(const DScene& sc, UtdShared& sh, bool need, const Cone& cone_from_src, const Geo& src_geo, V3 dst, const Aperture& ap, float k,
                        Counters& ctr, uint32_t& edges_fetched) {
    const unsigned lane = threadIdx.x & 31u;
    if (!__ballot_sync(0xffffffffu, need)) return 0.f;
    const V3 src = cone_from_src.o;
    const uint32_t n = need ? ap.n : 0u;
    const uint32_t nmax = __reduce_max_sync(0xffffffffu, n);
    C2 ts = mkc(0.f, 0.f), th = mkc(0.f, 0.f);
    uint32_t cnt = 0u;
    auto evaluate = [&](uint32_t m) {
        const bool has = lane < m;
        const uint32_t it = sh.item[has ? lane : 0u];
        const int owner = (int)(it >> 26); const uint32_t ed = it & 0x3ffffffu;
        ApHead ah; ah.wp = shfl3(ap.wp, owner); ah.fr.t = shfl3(ap.fr.t, owner); ah.fr.b = shfl3(ap.fr.b, owner); ah.fr.n = shfl3(ap.fr.n, owner);
        ah.size = shfl3(ap.size, owner); ah.wi = shfl3(ap.wi, owner); ah.k = __shfl_sync(0xffffffffu, ap.k, owner);
        const V3 osrc = shfl3(src, owner), odst = shfl3(dst, owner);
        Geo osg; osg.kind = __shfl_sync(0xffffffffu, src_geo.kind, owner); osg.p = shfl3(src_geo.p, owner); osg.id = __shfl_sync(0xffffffffu, src_geo.id, owner);
        const float ok_ = __shfl_sync(0xffffffffu, k, owner);
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f); uint32_t okf = 0u;
        if (has) {
            Wedge w; V3 p;
            if (wedge_build(sc, ah, ed, w) && wedge_diffraction_point(w, osrc, odst, p)) {      // passed in the filter: recomputed, not re-decided
                const V3 ui = osrc - p, uo = odst - p;
                const float ri = length(ui), ro = length(uo);
                C2 Ds, Dh; wedge_UTD(w, ah.k, ui / ri, uo / ro, ro, Ds, Dh);
                if (!(Dh.re == 0.f && Dh.im == 0.f && Ds.re == 0.f && Ds.im == 0.f)) {
                    const Geo eg = geo_edge(p, w.idx);
                    if (!(shadow_between(sc, eg, osg, ctr) || shadow_between(sc, eg, geo_point(odst), ctr))) {
                        const C2 phase = cexpi(-k_times_len(ok_, ro + ri));
                        const C2 a = phase * Ds, b = phase * Dh;
                        r = make_float4(a.re, a.im, b.re, b.im); okf = 1u;
                    }
                }
            }
        }
        sh.res[lane] = r; sh.ok[lane] = okf;
        __syncwarp();
        for (uint32_t q = 0; q < m; ++q) {
            if ((sh.item[q] >> 26) == lane && sh.ok[q]) { const float4 v = sh.res[q]; ts = ts + mkc(v.x, v.y); th = th + mkc(v.z, v.w); }
        }
        __syncwarp();
    };
    for (uint32_t j = 0; j < nmax; ++j) {
        bool pass = false; uint32_t ed = 0u;
        if (j < n) {
            ed = ap.edges[j];
            Wedge w;
            if (wedge_build(sc, ap, ed, w)) {
                ++edges_fetched;
                V3 p;
                if (wedge_diffraction_point(w, src, dst, p)) {
                    const V3 ui = src - p, uo = dst - p;
                    pass = !((dot(uo, w.nff) <= 0.f && dot(uo, w.nbf) <= 0.f) || (dot(ui, w.nff) <= 0.f && dot(ui, w.nbf) <= 0.f));
                }
            }
        }
        const unsigned pm = __ballot_sync(0xffffffffu, pass);
        if (pass) sh.item[cnt + (uint32_t)__popc(pm & ((1u << lane) - 1u))] = (lane << 26) | ed;
        cnt += (uint32_t)__popc(pm);
        __syncwarp();
        if (cnt >= 32u) {
            evaluate(32u);
            const uint32_t rest = cnt - 32u;
            const uint32_t mv = lane < rest ? sh.item[32u + lane] : 0u;
            __syncwarp();
            if (lane < rest) sh.item[lane] = mv;
            cnt = rest;
            __syncwarp();
        }
    }
    if (cnt) evaluate(cnt);
    if (!need) return 0.f;
    if (cone_contains(cone_from_src, dst)) {
        if (!shadow_between(sc, src_geo, geo_point(dst), ctr)) {
            const C2 phase = cexpi(-k_times_len(k, length(dst - src)));
            ts = ts + phase; th = th + phase;
        }
    }
    return (cnorm(ts) + cnorm(th)) / 2.f;
}

} // namespace wt
