// dbdpt.cuh -- plt_bdpt on the device (included by wavefront.cu after traverse()).
//
// Follows src/integrator/plt_bdpt.cpp:43-148, include/wt/integrator/plt_bdpt/{plt_bdpt_detail.hpp,vertex.hpp}, the Fraunhofer FSD
// (include/wt/interaction/fsd/fraunhofer/*.hpp, src/interaction/fsd/fraunhofer/*.cpp), gaussian2d_t::integrate_triangle
// (src/math/gaussian2d.cpp:96-192) and clip_triangle_z (include/wt/math/intersect/clip.hpp)   [paths under /root/reference].
//
// Execution model (round 1): persistent threads, one sample per thread at a time; the two subpaths' vertices and the
// Fraunhofer apertures live in a per-thread arena in HBM laid out word-interleaved across threads (word w of thread s at
// arena[w*P + s]) so that lock-step accesses coalesce.  A wavefront split (walk kernels + connection kernel) is the next step.
#pragma once

namespace wt {

#define WT_NI __device__ __noinline__      // BDPT is one big per-thread kernel: keep its large pieces out of line (compile time, i-cache)

constexpr float kSqrtPi = 1.77245385090551602730f;
constexpr float kInvSqrtPi = 0.56418958354775628695f;
constexpr int kFsdChunk = 48;               // Fraunhofer aperture segments staged in shared memory at a time by the sampler kernel
// Per-sample record sizes are run-time (dtrav.cuh Caps): cap.verts vertices per subpath (max_depth + 2), cap.ap_walk apertures per subpath of
// cap.seg segments each.  Defaults (wavefront.cu): 4 apertures of 48 segments; grown by wtgpu_render when a sample needs more.

WT_D float sincf_(float x) {                // include/wt/math/common.hpp:414-434
    const float t0 = 1.1920929e-7f, t2 = 0.00034526698300124390839884978618400831996329879769945f, tn = 0.018581361171917516667460937040007436176452688944747f;
    if (fabsf(x) >= tn) return pm::sinf(x) / x;
    float r = 1.f;
    if (fabsf(x) >= t0) { const float x2 = x * x; r -= x2 / 6.f; if (fabsf(x) >= t2) r += (x2 * x2) / 120.f; }
    return r;
}

// ---- gaussian2d with x=(1,0), mu=0 (include/wt/math/distribution/gaussian2d.hpp)
struct G2 { V2 s, rs; float norm; };
WT_D G2 g2_make(V2 sg) { G2 g; g.s = sg; g.rs = mk2(1.f / sg.x, 1.f / sg.y); g.norm = kInvTwoPi * (1.f / sg.x) * (1.f / sg.y); return g; }
WT_D bool g2_dirac(const G2& g) { return g.s.x == 0.f || g.s.y == 0.f; }
WT_D float g2_pdf(const G2& g, V2 p) { const V2 u = p * g.rs; return !g2_dirac(g) ? g.norm * pm::expf(-dot(u, u) / 2.f) : ((p.x == 0.f && p.y == 0.f) ? WT_INF : 0.f); }
WT_D V2 g2_canon(const G2& g, V2 v) {
    const V2 p = mk2(dot(mk2(1.f, 0.f), v), dot(mk2(-0.f, 1.f), v));
    if (!g2_dirac(g)) return p * g.rs;
    return mk2(p.x == 0.f ? 0.f : WT_INF, p.y == 0.f ? 0.f : WT_INF);
}
namespace g2d {     // src/math/gaussian2d.cpp:24-94
WT_D float Igg0(const DScene& sc, float a, float b, float c, float d) {
    const float n2 = 1.f / (a + 2.f * c * c), n = sqrtf(n2);
    return -kSqrtPi / 2.f * n * pm::expf(-2.f * a * sqrf(d - b * c) * n2) * (erf_lut(sc, (a * b + 2.f * c * d) * n) - erf_lut(sc, (a * (1.f + b) + 2.f * c * (c + d)) * n));
}
WT_D float Igg1(const DScene& sc, float a, float b, float c, float d) {
    const float n2 = 1.f / (a + 2.f * c * c), n = sqrtf(n2);
    return -kSqrtPi / 2.f * n * pm::expf(-2.f * a * sqrf(d - b * c) * n2) * (2.f * erf_lut(sc, a * (d / c - b) * n) + erf_lut(sc, (a * b + 2.f * c * d) * n) + erf_lut(sc, (a * (1.f + b) + 2.f * c * (c + d)) * n));
}
WT_D float Ig0(const DScene& sc, float a, float b) { const float n = sqrtf(1.f / a); return -kSqrtPi / 2.f * n * (erf_lut(sc, a * b * n) - erf_lut(sc, a * (1.f + b) * n)); }
WT_D float Ig1(const DScene& sc, float a, float b, float c, float d) {
    const float sa = sqrtf(a), n = 1.f / sa, dc = d / c;
    return -kSqrtPi / 2.f * n * (signf_(b) * erf_lut(sc, sa * fabsf(b)) + signf_(1.f + b) * erf_lut(sc, sa * fabsf(1.f + b)) - 2.f * signf_(b - dc) * erf_lut(sc, sa * fabsf(b - dc)));
}
// Out of line and rolled: inlined twice with its four-term loop unrolled this was 16 copies of Igg0/Igg1 (~2500 SASS instructions), and
// k_bd_resolve -- which runs it with 2-3 lanes per warp -- stalled 6.3 cycles per issue on instruction fetch (profiles/r01s3_ncu_full_digest.txt).
// Scalar arguments and result: the call moves nothing through local memory.  Same terms, same order of additions.
__device__ __noinline__ float Ige(const DScene& sc, float a, float b, float c, float d) {
    const float dc = d / c;
    const bool in = c != 0.f && -dc > 0.f && -dc < 1.f;
    const float sv[4] = { 0.6517755981618476f, 3.250040490513459f, 31.86882707224491f, 778.6613983601425f };
    const float wv[4] = { 0.2936683276537767f, 0.135758042187825f, 0.05245255757691102f, 0.01673209873360605f };
    float acc = 0.f;
#pragma unroll 1
    for (int i = 0; i < 4; ++i) { const float q = sqrtf(sv[i]); const float t = wv[i] * (in ? Igg1(sc, a, b, c * q, d * q) : Igg0(sc, a, b, c * q, d * q)); acc = i == 0 ? t : acc + t; }
    return ((d != 0.f && c != 0.f) ? signf_(d) : (d == 0.f && c != 0.f) ? signf_(c) : 1.f) * ((in ? Ig1(sc, a, b, c, d) : Ig0(sc, a, b)) - 2.f * acc);
}
WT_D bool pit2(V2 p, V2 a, V2 b, V2 c) {        // math/util.hpp:69-82
    const float s1 = diff_prod(p.x - b.x, a.y - b.y, a.x - b.x, p.y - b.y), s2 = diff_prod(p.x - c.x, b.y - c.y, b.x - c.x, p.y - c.y), s3 = diff_prod(p.x - a.x, c.y - a.y, c.x - a.x, p.y - a.y);
    const bool neg = s1 < 0.f || s2 < 0.f || s3 < 0.f, pos = s1 > 0.f || s2 > 0.f || s3 > 0.f;
    return !(neg && pos);
}
}
// gaussian2d_t::integrate_triangle (src/math/gaussian2d.cpp:96-192) in three parts, so that the rare but long quadrature branch can be served by
// a whole warp (k_bd_resolve): g2_classify takes the early-outs and decides the branch, g2_quadrature / g2_quadrature_warp is the 0.002-step
// Riemann sum of :141-164 for triangles with a short edge, g2_analytic the four-Gaussian erf approximation of :166-192.
enum : int { G2_DONE = 0, G2_QUADRATURE = 1, G2_ANALYTIC = 2 };
WT_D int g2_classify(const G2& g, V2& a, V2& b, V2& c, float& value) {
    if (g2_dirac(g)) {
        const float A = (a.x * (b.y - c.y) - a.y * (b.x - c.x)) + (b.x * c.y - c.x * b.y);
        const float sA = signf_(A);
        const float bx = sA * diff_prod(b.x, c.y, c.x, b.y), by = sA * diff_prod(c.x, a.y, a.x, c.y);
        value = (bx >= 0.f && by >= 0.f && bx + by <= fabsf(A)) ? 1.f : 0.f;
        return G2_DONE;
    }
    const float L = 3.f;
    a = g2_canon(g, a); b = g2_canon(g, b); c = g2_canon(g, c);
    if (min3f(a.x, b.x, c.x) >= L || max3f(a.x, b.x, c.x) <= -L || min3f(a.y, b.y, c.y) >= L || max3f(a.y, b.y, c.y) <= -L) { value = 0.f; return G2_DONE; }
    const bool ain = length2(a) <= sqrf(L), bin = length2(b) <= sqrf(L), cin = length2(c) <= sqrf(L);
    if (!ain && !bin && !cin) {
        const bool iab = intersect_edge_ellipse_points(a, b, L, L) > 0, iac = intersect_edge_ellipse_points(a, c, L, L) > 0, ibc = intersect_edge_ellipse_points(b, c, L, L) > 0;
        if (!iab && !iac && !ibc) { value = g2d::pit2(mk2(0.f, 0.f), a, b, c) ? 1.f : 0.f; return G2_DONE; }
    }
    const float min_len = min3f(length2(a - b), length2(a - c), length2(b - c));
    return min_len < 1e-3f ? G2_QUADRATURE : G2_ANALYTIC;
}
// The quadrature's (y, x) sample sequence, shared by the serial and the warp form: float accumulation of y and x exactly as the reference's loops.
template <class F> WT_D void g2_quadrature_points(V2 a, V2 b, V2 c, F&& f) {
    const float L = 3.f, delta = .002f;
    if (b.y < a.y) { const V2 t = a; a = b; b = t; }
    if (c.y < a.y) { const V2 t = a; a = c; c = t; }
    const float ab = b.y == a.y ? WT_INF : (b.x - a.x) / (b.y - a.y);
    const float ac = c.y == a.y ? WT_INF : (c.x - a.x) / (c.y - a.y);
    const float bc = c.y == b.y ? WT_INF : (c.x - b.x) / (c.y - b.y);
    for (float y = fmaxf(-L, a.y + delta / 2.f); y < fminf(L, fmaxf(b.y, c.y)); y += delta) {
        float x0 = y < b.y ? ab * (y - a.y) + a.x : bc * (y - b.y) + b.x;
        float x1 = y < c.y ? ac * (y - a.y) + a.x : bc * (y - b.y) + b.x;
        if (x0 > x1) { const float t = x0; x0 = x1; x1 = t; }
        for (float x = fmaxf(-L, x0) + delta / 2.f; x < fminf(L, x1); x += delta) f(x, y);
    }
}
WT_D float g2_quadrature(V2 a, V2 b, V2 c) {
    float ret = 0.f;
    g2_quadrature_points(a, b, c, [&](float x, float y) { ret += pm::expf(-(sqrf(x) + sqrf(y)) / 2.f); });
    return ret * kInvTwoPi * sqrf(.002f);
}
// The same sum by a whole warp (all 32 lanes call with the same triangle): the sample sequence is enumerated by every lane (cheap, warp-uniform),
// lane j keeps sample j of each batch of 32, the exponentials of a batch are evaluated in parallel and added IN SEQUENCE ORDER -- the result is
// bit-identical to g2_quadrature.  One thread walking up to ~10^4 samples (each a binary64-internal exp) used to decide the duration of the
// whole k_bd_resolve launch (profiles/r02_phases.txt).
WT_D float g2_quadrature_warp(V2 a, V2 b, V2 c) {
    const unsigned lane = threadIdx.x & 31u;
    float ret = 0.f, mx = 0.f, my = 0.f; uint32_t cnt = 0u;
    auto flush = [&](uint32_t n) {
        const float e = lane < n ? pm::expf(-(sqrf(mx) + sqrf(my)) / 2.f) : 0.f;
        for (uint32_t j = 0; j < n; ++j) ret += __shfl_sync(0xffffffffu, e, (int)j);
    };
    g2_quadrature_points(a, b, c, [&](float x, float y) {
        if ((cnt & 31u) == lane) { mx = x; my = y; }
        ++cnt;
        if ((cnt & 31u) == 0u) flush(32u);
    });
    if (cnt & 31u) flush(cnt & 31u);
    return ret * kInvTwoPi * sqrf(.002f);
}
// A warp's quadrature pieces (all 32 lanes call; `mine`: this lane holds one).  Under a beam much wider than the mesh's triangles EVERY piece takes the
// quadrature branch with ~10^2 samples: then each lane sums its own (32 pieces side by side).  Only a piece with very many samples -- a sliver
// across the whole 6-sigma window -- is worth the warp's joint evaluation, where the samples of ONE piece are spread over the lanes and the other
// lanes' pieces wait.  Same sample sequence, same order of additions either way.
constexpr float kQuadCoopSamples = 4096.f, kQuadQueueSamples = 1024.f;
// The flat Gaussian-power kernel goes one step further: a piece with more than ~10^3 samples becomes a task of its own (one warp per piece,
// k_bd_quad_tasks), so that a 32-triangle chunk full of long slivers does not decide the duration of the launch.
struct QuadQ { float4* tasks; int* n; int cap; };        // task = (a.xy, b.xy) (c.xy, destination index as two words)
WT_D float g2_quadrature_mixed(bool mine, V2 pa, V2 pb, V2 pc, const QuadQ* q = nullptr, size_t dst = 0) {
    const unsigned lane = threadIdx.x & 31u, FULL = 0xffffffffu;
    float val = 0.f; bool big = false;
    if (mine) {
        const float L = 3.f, delta = .002f;
        const float h = fminf(L, max3f(pa.y, pb.y, pc.y)) - fmaxf(-L, min3f(pa.y, pb.y, pc.y)), w = fminf(L, max3f(pa.x, pb.x, pc.x)) - fmaxf(-L, min3f(pa.x, pb.x, pc.x));
        const float est = fmaxf(h, 0.f) * fmaxf(w, 0.f) * (.5f / (delta * delta));
        if (q && est > kQuadQueueSamples) {
            const int at = atomicAdd(q->n, 1);
            if (at < q->cap) {
                q->tasks[2 * (size_t)at] = make_float4(pa.x, pa.y, pb.x, pb.y);
                q->tasks[2 * (size_t)at + 1] = make_float4(pc.x, pc.y, __uint_as_float((uint32_t)dst), __uint_as_float((uint32_t)(dst >> 32)));
                mine = false;       // (the task's warp writes the value)
            }
        }
        if (mine) { big = est > kQuadCoopSamples; if (!big) val = g2_quadrature(pa, pb, pc); }
    }
    unsigned m = __ballot_sync(FULL, mine && big);
    while (m) {
        const int src = __ffs(m) - 1; m &= m - 1u;
        const V2 qa = mk2(__shfl_sync(FULL, pa.x, src), __shfl_sync(FULL, pa.y, src));
        const V2 qb = mk2(__shfl_sync(FULL, pb.x, src), __shfl_sync(FULL, pb.y, src));
        const V2 qc = mk2(__shfl_sync(FULL, pc.x, src), __shfl_sync(FULL, pc.y, src));
        const float r = g2_quadrature_warp(qa, qb, qc);
        if ((int)lane == src) val = r;
    }
    return val;
}
WT_NI float g2_analytic(const DScene& sc, V2 a, V2 b, V2 c) {
    const V2 t0 = b - a, t1 = c - a;         // T = mat2(t0, t1) columns
    const float detT = t0.x * t1.y - t1.x * t0.y;
    const float od = 1.f / detT;
    M2 Ti; Ti.c0x = t1.y * od; Ti.c0y = -t0.y * od; Ti.c1x = -t1.x * od; Ti.c1y = t0.x * od;
    const V2 mu0 = m2mul(Ti, a);
    M2 T; T.c0x = t0.x; T.c0y = t0.y; T.c1x = t1.x; T.c1y = t1.y;
    M2 Tt; Tt.c0x = t0.x; Tt.c0y = t1.x; Tt.c1x = t0.y; Tt.c1y = t1.y;
    const M2 A = m2mm(Tt, T);
    const float detA = A.c0x * A.c1y - A.c1x * A.c0y;
    const float Sxy = A.c0y, Syy = A.c1y;
    if (Syy <= 0.f || detA <= 0.f) return 0.f;
    const float denom = 1.f / sqrtf(2.f * Syy);
    const float pa = detA * sqrf(denom), pb = mu0.x;
    const float c0 = Sxy * denom, d0 = (Sxy * mu0.x + Syy * mu0.y) * denom, q = .5f / denom;
    const float I0 = g2d::Ige(sc, pa, pb, c0 - q, d0 + q), I1 = g2d::Ige(sc, pa, pb, c0, d0);
    return kInvSqrtPi / 2.f * fabsf(detT * denom) * fmaxf(0.f, I0 - I1);
}
WT_NI float g2_integrate_triangle(const DScene& sc, const G2& g, V2 a, V2 b, V2 c) {
    float v = 0.f;
    const int kind = g2_classify(g, a, b, c, v);
    return kind == G2_DONE ? v : kind == G2_QUADRATURE ? g2_quadrature(a, b, c) : g2_analytic(sc, a, b, c);
}
WT_D G2 wavefront_of(const Beam& b, float d) {       // beam_generic.hpp:130-139 + gaussian_wavefront.hpp:34-40
    const V3 fp = beam_footprint(b, d);
    const G2 g = g2_make(mk2(fp.x / kEnvelope, fp.y / kEnvelope));
    return g2_dirac(g) ? g2_make(mk2(0.f, 0.f)) : g;
}

// clip_triangle_z (include/wt/math/intersect/clip.hpp:34-88)
struct Clip { V3 vs[5]; int tris; };
WT_D void clip_tri(const Clip& c, int idx, V3 o[3]) {
    if (idx == 0) { o[0] = c.vs[0]; o[1] = c.vs[1]; o[2] = c.vs[2]; } else if (idx == 1) { o[0] = c.vs[2]; o[1] = c.vs[0]; o[2] = c.vs[c.tris == 2 ? 3 : 4]; } else { o[0] = c.vs[4]; o[1] = c.vs[2]; o[2] = c.vs[3]; }
}
WT_D Clip clip_triangle_z(V3 a, V3 b, V3 c, Range zr) {
    const V3 ppmax = mk3(0.f, 0.f, zr.mx), ppmin = mk3(0.f, 0.f, zr.mn), n = mk3(0.f, 0.f, 1.f);
    // the three edges go through ONE copy of the clipping code (vertices rotate through registers): same edges, same order, a third of the code
    V3 ti = a, tn = b, tl = c;
    auto zcls = [&](const V3& v) { return v.z > zr.mx ? 1 : v.z < zr.mn ? -1 : 0; };
    int ci = zcls(a), cn = zcls(b), cl = zcls(c);
    Clip r; int idx = 0;
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
        if (ci == 0) { if (idx < 5) r.vs[idx++] = ti; }
        if (cn != ci) {
            V3 pt;
            const bool ok = intersect_edge_plane(ti, tn, ci == -1 ? ppmin : (ci == 1 || cn == 1) ? ppmax : ppmin, n, pt);
            if (idx < 5) r.vs[idx++] = ok ? pt : (ci != 0 ? ti : tn);
            if (cn != 0 && ci != 0) {
                V3 p2; const bool ok2 = intersect_edge_plane(ti, tn, cn == 1 ? ppmax : ppmin, n, p2);
                if (idx < 5) r.vs[idx++] = ok2 ? p2 : tn;
            }
        }
        const V3 tv = ti; ti = tn; tn = tl; tl = tv;
        const int tc = ci; ci = cn; cn = cl; cl = tc;
    }
    r.tris = idx < 3 ? 0 : idx == 3 ? 1 : idx == 4 ? 2 : 3;
    return r;
}
WT_D V2 cone_project_local(const Cone& c, V3 p, float z) {       // elliptic_cone.hpp:205-214
    const V2 xy = mk2(p.x, p.y);
    const float scale = (c.ta * z + c.x0) / fabsf(c.ta * p.z + c.x0);
    return (c.x0 == 0.f && c.ta == 0.f) ? xy : xy * scale;
}

// ================================================================================================ Fraunhofer FSD
struct FEdge { V2 e, v; C2 a_b, iab_2; };
struct FHead { float P0, P0_pdf, psi02, recp_I, k; Frame frame; uint32_t n; };
constexpr float kPA1 = 0.0049361075794549872500f, kPA2 = 0.21899789398059305541f, kP0s = 0.288675134594813f / 4.f;
WT_D float falpha1(float x, float y) { return x == 0.f ? 0.f : kInvTwoPi * y / (x * (x * x + y * y)) * (pm::cosf(x / 2.f) - sincf_(x / 2.f)); }
WT_D float falpha2(float x, float y) { return x == 0.f ? 0.f : kInvTwoPi * y / (x * x + y * y) * sincf_(x / 2.f); }
WT_D float fchi_e(V2 xi) { const float t = 1.f + 0.830092714835359f * dot(xi, xi), t2 = t * t, t3 = t2 * t; return fmaxf(0.f, 1.f - (3.f / t2 - 2.f / t3)); }
WT_D float fchi_0(V2 xi) { xi = xi / kP0s; return pm::expf(-.5f * dot(xi, xi)); }
WT_D V2 fzeta(const FEdge& e, V2 xi) { return mk2(xi.x * e.e.x + xi.y * e.e.y, xi.x * e.e.y + xi.y * (-e.e.x)); }
WT_D C2 fPsi(const FEdge& e, V2 xi) {
    const V2 z = fzeta(e, xi);
    const C2 s = e.a_b * falpha1(z.x, z.y) + e.iab_2 * falpha2(z.x, z.y);
    const float rho = length2(e.e), th = -dot(e.v, xi);
    float sn, cs; pm::sincosf(th, &sn, &cs);
    return mkc(rho * cs, rho * sn) * s;
}
WT_D float fPsi2(const FEdge& e, V2 xi) { const V2 z = fzeta(e, xi); return sqrf(length2(e.e)) * cnorm(e.a_b * falpha1(z.x, z.y) + e.iab_2 * falpha2(z.x, z.y)); }
WT_D float fPj(const FEdge& e) { return sqrf(length2(e.e)) * kPA1 * cnorm(e.a_b) + sqrf(length2(e.e)) * kPA2 * cnorm(e.iab_2); }

// per-thread arena accessor: word w of this thread at base[w*P + slot]
// this sample's records, contiguous (AoS): a vertex or an aperture is read by one thread at a time.  Layout: 2 x cap.verts vertices (sensor
// subpath, emitter subpath), then 2 x cap.ap_walk apertures of ap_words = 16 + 9 cap.seg words.
struct Arena { float* base; uint32_t ap0, ap_words; };
constexpr uint32_t kVertWords = 68;
WT_D float& aw(const Arena& A, uint32_t w) { return A.base[w]; }
WT_D Arena mk_arena(const DScene& sc, float* arena, uint32_t slot) {
    Arena A; A.base = arena + (size_t)slot * sc.cap.arena_words; A.ap0 = 2u * sc.cap.verts * kVertWords; A.ap_words = sc.cap.ap_words; return A;
}
WT_D uint32_t ap_base(const Arena& A, int ai) { return A.ap0 + (uint32_t)ai * A.ap_words; }
WT_D FEdge ap_edge(const Arena& A, int ai, uint32_t j) {
    const uint32_t b = ap_base(A, ai) + 16 + j * 9;
    FEdge e; e.e = mk2(aw(A, b), aw(A, b + 1)); e.v = mk2(aw(A, b + 2), aw(A, b + 3)); e.a_b = mkc(aw(A, b + 4), aw(A, b + 5)); e.iab_2 = mkc(aw(A, b + 6), aw(A, b + 7));
    return e;
}
WT_D float ap_edge_pdf(const Arena& A, int ai, uint32_t j) { return aw(A, ap_base(A, ai) + 16 + j * 9 + 8); }
WT_D FHead ap_head(const Arena& A, int ai) {
    const uint32_t b = ap_base(A, ai);
    FHead h; h.P0 = aw(A, b); h.P0_pdf = aw(A, b + 1); h.psi02 = aw(A, b + 2); h.recp_I = aw(A, b + 3); h.k = aw(A, b + 4);
    h.frame.t = mk3(aw(A, b + 5), aw(A, b + 6), aw(A, b + 7)); h.frame.b = mk3(aw(A, b + 8), aw(A, b + 9), aw(A, b + 10)); h.frame.n = mk3(aw(A, b + 11), aw(A, b + 12), aw(A, b + 13));
    h.n = __float_as_uint(aw(A, b + 14));
    return h;
}
WT_D float ap_ASF_unclamped(const Arena& A, int ai, uint32_t n, V2 xi) { C2 a = mkc(0.f, 0.f); for (uint32_t j = 0; j < n; ++j) a = a + fPsi(ap_edge(A, ai, j), xi); return cnorm(a); }
WT_D float ap_ASF(const Arena& A, int ai, const FHead& h, V2 xi) { return ap_ASF_unclamped(A, ai, h.n, xi) * fchi_e(xi) + h.psi02 * fchi_0(xi); }
WT_D float ap_sampling_density(const Arena& A, int ai, const FHead& h, V2 xi) {
    float d = 0.f; for (uint32_t j = 0; j < h.n; ++j) d += fPsi2(ap_edge(A, ai, j), xi);
    return d * fchi_e(xi) + h.P0 * kInvTwoPi / sqrf(kP0s) * fchi_0(xi);
}

// fraunhofer::free_space_diffraction_t ctor (src/interaction/fsd/fraunhofer/free_space_diffraction.cpp:22-129) into aperture slot `ai`.
// Returns the number of segments (0: empty aperture); an aperture of more than sc.cap.seg segments sets overflow and reports its size in need_seg.
WT_NI uint32_t fraunhofer_build(const DScene& sc, const Arena& A, int ai, const Frame& frame, float k, float total_power, const Cone& beam,
                                const uint32_t* edges, uint32_t n_edges, const G2& wf, bool& overflow, uint32_t& need_seg) {
    const V2 cse = wf.s * kEnvelope;
    const float r = fmaxf(cse.x, cse.y);
    const float max_len = .33f * r;
    float P_total = 0.f;
    uint32_t n = 0, n_all = 0;
    const uint32_t b0 = ap_base(A, ai);
    for (uint32_t ei = 0; ei < n_edges; ++ei) {
        const wtgpu_edge E = sc.edges[edges[ei]];
        if (dot(beam.d, mk3(E.n1)) * dot(beam.d, mk3(E.n2)) >= 0.f) continue;
        const V3 l1 = to_local(frame, mk3(E.a) - beam.o), l2 = to_local(frame, mk3(E.b) - beam.o);
        const V2 u1 = mk2(l1.x, l1.y), u2 = mk2(l2.x, l2.y);
        float t1 = 0.f, t2 = 1.f;
        const V2 q1 = mk2(u1.x / cse.x, u1.y / cse.y), q2 = mk2(u2.x / cse.x, u2.y / cse.y);
        if (!(dot(q1, q1) <= 1.f) || !(dot(q2, q2) <= 1.f)) {
            // intersect_edge_ellipse (misc.hpp:77-127): need points, t1, t2
            const V2 rs = mk2(1.f / cse.x, 1.f / cse.y);
            const V2 p0 = u1 * rs, p1 = u2 * rs, d = p1 - p0;
            const float aa = dot(d, d), bb = 2.f * dot(p0, d), cc = dot(p0, p0) - 1.f;
            const float det2 = bb * bb - 4.f * aa * cc;
            if (det2 <= 0.f || aa == 0.f) continue;
            const float ra = 1.f / aa, det = sqrtf(det2);
            float s1 = .5f * (-bb - signf_(bb) * det) * ra;
            float s2 = s1 == 0.f ? -bb * ra : cc * ra / s1;
            if (s1 > s2) { const float t = s1; s1 = s2; s2 = t; }
            const bool v1 = s1 >= 0.f && 1.f >= s1, v2 = s2 >= 0.f && 1.f >= s2;
            if (!v1 && !v2) continue;
            float rt1 = s1, rt2 = s2;
            if (!(v1 && v2)) { rt1 = v1 ? s1 : s2; rt2 = v1 ? s2 : s1; }
            t1 = fmaxf(0.f, rt1); t2 = fminf(1.f, rt2);
        }
        const V2 m1 = mk2(mixf(u1.x, u2.x, t1), mixf(u1.y, u2.y, t1)), m2 = mk2(mixf(u1.x, u2.x, t2), mixf(u1.y, u2.y, t2));
        const float len = length(m1 - m2);
        const int segments = max(1, int(roundf(len / max_len) + .5f));
        const float seg = 1.f / (float)segments;
        V2 v1 = m1;
        float a = sqrtf(g2_pdf(wf, v1));
        for (int i = 0; i < segments; ++i) {
            const float tt = mixf(t1, t2, (float)(i + 1) * seg);
            const V2 v2 = mk2(mixf(u1.x, u2.x, tt), mixf(u1.y, u2.y, tt));
            const float b = sqrtf(g2_pdf(wf, v2));
            if (a > 0.f || b > 0.f) {
                const V2 v = (v1 + v2) / 2.f, e = v2 - v1;
                FEdge fe; fe.e = mk2(e.x * 1000.f, e.y * 1000.f); fe.v = mk2(v.x * 1000.f, v.y * 1000.f); fe.a_b = mkc(a - b, 0.f);
                fe.iab_2 = (mkc(0.f, 1.f) * mkc(a + b, 0.f)) * (1.f / 2.f);
                // (c_t{0,1}*(ca+cb))/2 : complex / float
                fe.iab_2 = mkc((0.f * (a + b) - 1.f * 0.f) / 2.f, (0.f * 0.f + 1.f * (a + b)) / 2.f);
                const float pdf = fPj(fe);
                if (pdf > 0.f) {
                    ++n_all;
                    if (n < sc.cap.seg) {
                        const uint32_t w = b0 + 16 + n * 9;
                        aw(A, w) = fe.e.x; aw(A, w + 1) = fe.e.y; aw(A, w + 2) = fe.v.x; aw(A, w + 3) = fe.v.y;
                        aw(A, w + 4) = fe.a_b.re; aw(A, w + 5) = fe.a_b.im; aw(A, w + 6) = fe.iab_2.re; aw(A, w + 7) = fe.iab_2.im; aw(A, w + 8) = pdf;
                        ++n; P_total += pdf;
                    } else overflow = true;
                }
            }
            v1 = v2; a = b;
        }
    }
    if (n_all > sc.cap.seg) need_seg = max(need_seg, n_all);
    const float r0 = 3.f * kP0s;
    const V2 dirs[8] = { mk2(-kInvSqrtTwo, -kInvSqrtTwo), mk2(-1.f, 0.f), mk2(-kInvSqrtTwo, kInvSqrtTwo), mk2(0.f, 1.f), mk2(kInvSqrtTwo, kInvSqrtTwo), mk2(1.f, 0.f), mk2(kInvSqrtTwo, -kInvSqrtTwo), mk2(0.f, -1.f) };
    float acc = 0.f;
    for (int i = 0; i < 8; ++i) acc = acc + ap_ASF_unclamped(A, ai, n, r0 * dirs[i]);
    const float psi02 = acc / 8.f;
    const float P0 = (kTwoPi * sqrf(kP0s) * psi02) / sqrf(k);
    P_total += P0;
    float P0_pdf;
    if (P_total > 0.f) { const float rp = 1.f / P_total; P0_pdf = P0 * rp; for (uint32_t j = 0; j < n; ++j) aw(A, b0 + 16 + j * 9 + 8) *= rp; }
    else { P0_pdf = 1.f; n = 0; }
    aw(A, b0) = P0; aw(A, b0 + 1) = P0_pdf; aw(A, b0 + 2) = psi02; aw(A, b0 + 3) = total_power > 0.f ? 1.f / total_power : 0.f; aw(A, b0 + 4) = k;
    aw(A, b0 + 5) = frame.t.x; aw(A, b0 + 6) = frame.t.y; aw(A, b0 + 7) = frame.t.z; aw(A, b0 + 8) = frame.b.x; aw(A, b0 + 9) = frame.b.y; aw(A, b0 + 10) = frame.b.z;
    aw(A, b0 + 11) = frame.n.x; aw(A, b0 + 12) = frame.n.y; aw(A, b0 + 13) = frame.n.z; aw(A, b0 + 14) = __uint_as_float(n);
    return n;
}

// fsd_lut_t (include/wt/interaction/fsd/fraunhofer/fsd_lut.hpp:37-69)
struct FLut { uint32_t N, M; const float *th1, *th2, *c1, *c2; };
WT_D float flerp1(float x, const float* tbl, uint32_t S) {
    x *= (float)(S - 1u);
    const uint32_t l = min((uint32_t)x, S - 1u), h = min(l + 1u, S - 1u);
    const float f = x - floorf(x);
    return f * __ldg(tbl + h) + (1.f - f) * __ldg(tbl + l);
}
WT_D V2 flut_sample(const FLut& L, V3 r3, const float* th, const float* cd) {
    const float theta = flerp1(r3.x, th, L.N);
    const float tf = theta * 2.f / kPi;
    float x = tf * (float)(L.M - 1u);
    const uint32_t l = min((uint32_t)x, L.M - 1u), h = min(l + 1u, L.M - 1u);
    const float f = x - floorf(x);
    const float r = fmaxf(0.f, f * flerp1(r3.y, cd + (size_t)h * L.M, L.M) + (1.f - f) * flerp1(r3.y, cd + (size_t)l * L.M, L.M));
    V2 z = r * mk2(pm::cosf(theta), pm::sinf(theta));
    const int q = min(3, (int)(r3.z * 4.f));
    z.x *= (((q + 1) / 2) % 2 == 0 ? 1.f : -1.f);
    z.y *= ((q / 2) % 2 == 0 ? 1.f : -1.f);
    return z;
}
// fsd_sampler (src/interaction/fsd/fraunhofer/fsd_sampler.cpp:37-113) + free_space_diffraction_t::sample (free_space_diffraction.hpp:68-93)
WT_NI void fraunhofer_sample(const Arena& A, int ai, const FLut& lut, Sampler& smp, V3& wo, float& dpd, float& weight) {
    wo = mk3(0.f, 0.f, 1.f); dpd = 0.f; weight = 0.f;
    const FHead h = ap_head(A, ai);
    const bool rej = h.n > 1u;
    const uint32_t max_tries = h.n * 1024u;
    const float recp_M = 1.f / (float)h.n;
    for (uint32_t tr = 0; tr < max_tries; ++tr) {
        // sampleN
        const float p = rnd(smp) * 1.f;
        float cdf = 0.f; uint32_t sel = h.n;
        for (uint32_t i = 0; i < h.n; ++i) { cdf += i == 0 ? h.P0_pdf : ap_edge_pdf(A, ai, i - 1u); if (p < cdf) { sel = i; break; } }
        V2 xi;
        if (sel == 0u) xi = kP0s * normal2d(rnd2(smp));
        else {
            const FEdge e = ap_edge(A, ai, sel - 1u);
            const V2 m = mk2(e.e.y, -e.e.x);
            const float od = 1.f / (e.e.x * m.y - m.x * e.e.y);
            const float i00 = m.y * od, i01 = -e.e.y * od, i10 = -m.x * od, i11 = e.e.x * od;     // glm::inverse, [col][row]
            const float Aa = cnorm(e.a_b), Bb = cnorm(e.iab_2);
            const float pp = rnd(smp) * (Aa + Bb);
            const V3 r3 = rnd3(smp);
            const V2 z = pp < Aa ? flut_sample(lut, r3, lut.th1, lut.c1) : flut_sample(lut, r3, lut.th2, lut.c2);
            xi = mk2(z.x * i00 + z.y * i01, z.x * i10 + z.y * i11);
        }
        const float g = ap_sampling_density(A, ai, h, xi), f = ap_ASF(A, ai, h, xi);
        const bool done = rej ? rnd(smp) * g < f * recp_M : true;
        if (done) {
            const float pdf = f * h.recp_I;
            if (pdf > 0.f) {
                const V2 zeta = xi / h.k;
                const V2 wl = mk2(zeta.x / sqrtf(1.f + sqrf(zeta.x)), zeta.y / sqrtf(1.f + sqrf(zeta.y)));
                const float wo2 = length2(wl);
                if (wo2 < .85f) { wo = mk3(wl.x, wl.y, sqrtf(1.f - wo2)); dpd = pdf; weight = 1.f; }
            }
            return;
        }
    }
}
WT_D float fraunhofer_pdf(const Arena& A, int ai, const FHead& h, V3 wl) {     // free_space_diffraction.hpp:99-115
    const float wo2 = length2(mk2(wl.x, wl.y));
    if (wl.z <= 0.f || wo2 >= .85f) return 0.f;
    const V2 xi = h.k * mk2(wl.x / sqrtf(1.f - sqrf(wl.x)), wl.y / sqrtf(1.f - sqrf(wl.y)));
    const float p = ap_ASF(A, ai, h, xi) * h.recp_I;
    return (0.f <= p && p < 1e+2f) ? p : 0.f;
}

// ================================================================================================ vertices
enum : uint32_t { BV_SENSOR = 0u, BV_EMITTER = 1u, BV_SURFACE = 2u, BV_FSD = 3u };
enum : uint32_t { BG_NONE = 0u, BG_POINT = 1u, BG_SURFACE = 2u, BG_DUMMY = 4u };
struct alignas(16) BVertex {            // vertex_t (integrator/plt_bdpt/vertex.hpp:49-81)
    uint32_t type, fwd, delta, ffsd;
    float pdf_fwd, pdf_bwd, rr;
    uint32_t gkind; V3 p; uint32_t tuid; V2 bary; Footprint fp; V3 dn;
    int32_t emitter, bsdf, fsd;
    uint32_t pad_;
    Beam beam;
};
static_assert(sizeof(BVertex) == kVertWords * 4, "vertex record size");
WT_D void bv_store(const Arena& A, uint32_t idx, const BVertex& v) {
    const float4* s = reinterpret_cast<const float4*>(&v); float4* d = reinterpret_cast<float4*>(A.base + idx * kVertWords);
#pragma unroll
    for (uint32_t i = 0; i < kVertWords / 4; ++i) d[i] = s[i];
}
WT_D void bv_load(const Arena& A, uint32_t idx, BVertex& v) {
    float4* d = reinterpret_cast<float4*>(&v); const float4* s = reinterpret_cast<const float4*>(A.base + idx * kVertWords);
#pragma unroll
    for (uint32_t i = 0; i < kVertWords / 4; ++i) d[i] = s[i];
}
// read-only view of a stored vertex: fields are fetched from the arena when used (no 272-B local copy)
WT_D const BVertex& bv_ref(const Arena& A, uint32_t idx) { return *reinterpret_cast<const BVertex*>(A.base + idx * kVertWords); }
// word offsets of the scalars the MIS walk touches
constexpr uint32_t kOffDelta = 2, kOffPdfFwd = 4, kOffPdfBwd = 5, kOffRr = 6;
WT_D Geo bv_geo(const BVertex& v) { return geo_surface(v.p, v.tuid, v.gkind == BG_SURFACE); }
WT_D Surface bv_surface(const DScene& sc, const BVertex& v) {
    if (v.gkind == BG_SURFACE) { Surface s = make_surface(sc, v.tuid, v.bary, v.p); s.fp = v.fp; return s; }
    return make_dummy_surface(v.dn, v.p);
}

struct BCtx { const DScene* sc; Arena A; FLut lut; Counters* ctr; bool overflow; uint32_t need_seg, need_ap, need_verts; };
WT_D BCtx mk_bctx(const DScene& sc, float* arena, uint32_t slot, const FLut& lut, Counters* ctr) {
    BCtx c; c.sc = &sc; c.A = mk_arena(sc, arena, slot); c.lut = lut; c.ctr = ctr; c.overflow = false; c.need_seg = 0u; c.need_ap = 0u; c.need_verts = 0u; return c;
}
WT_D bool bv_area_emitter(const BCtx& c, const BVertex& v) { return v.type == BV_EMITTER && c.sc->emitters[v.emitter].type == WTGPU_EMITTER_AREA; }
WT_D bool bv_on_surface(const BCtx& c, const BVertex& v) { return v.type == BV_SURFACE || bv_area_emitter(c, v) || (v.type == BV_SENSOR && (v.gkind == BG_SURFACE || v.gkind == BG_DUMMY)); }
WT_D bool bv_has_srf_normal(const BCtx& c, const BVertex& v) { return v.type == BV_SURFACE || bv_area_emitter(c, v); }
WT_D V3 bv_ng(const BCtx& c, const BVertex& v) { if (!bv_has_srf_normal(c, v)) return mk3(0.f, 0.f, 1.f); const Tri3 t = load_tri(*c.sc, v.tuid); return t.n; }
WT_D V3 bv_ns(const BCtx& c, const BVertex& v) { if (!bv_has_srf_normal(c, v)) return mk3(0.f, 0.f, 1.f); return bv_surface(*c.sc, v).shading.n; }
WT_D int32_t bv_emitter(const BCtx& c, const BVertex& v) { return v.type == BV_EMITTER ? v.emitter : c.sc->shapes[c.sc->tri_meta[v.tuid].shape_idx].emitter; }
WT_D bool bv_on_emitter(const BCtx& c, const BVertex& v) { return v.type == BV_EMITTER || (v.type == BV_SURFACE && c.sc->shapes[c.sc->tri_meta[v.tuid].shape_idx].emitter >= 0); }
WT_D bool em_delta_dir(const BCtx& c, int32_t e) { return c.sc->emitters[e].type == WTGPU_EMITTER_DIRECTIONAL; }
WT_D bool em_delta_pos(const BCtx& c, int32_t e) { const uint32_t t = c.sc->emitters[e].type; return t == WTGPU_EMITTER_POINT || t == WTGPU_EMITTER_SPOT; }
WT_D bool bv_delta_emitter(const BCtx& c, const BVertex& v) { return v.type == BV_EMITTER && (em_delta_dir(c, v.emitter) || em_delta_pos(c, v.emitter)); }
WT_D bool sensor_delta_pos(const BCtx& c) { return c.sc->sensor.type == WTGPU_SENSOR_PERSPECTIVE; }
WT_D bool bv_delta_sensor(const BCtx& c, const BVertex& v) { return v.type == BV_SENSOR && sensor_delta_pos(c); }
WT_D bool bv_nondelta_interaction(const BVertex& v) { return (v.type == BV_SURFACE || v.type == BV_FSD) && !v.delta; }
WT_D bool bv_connectible(const BCtx& c, const BVertex& v) {      // vertex.hpp:415-425
    if (v.type == BV_FSD) return true;
    if (v.type == BV_EMITTER) return !em_delta_dir(c, v.emitter);
    if (v.type == BV_SENSOR) return true;
    return !bsdf_is_delta_only(*c.sc, v.bsdf, v.beam.k);
}
WT_D float dir_to_area(const BCtx& c, Pd dpdf, V3 p, const BVertex& next) {     // vertex.hpp:224-243
    const float dv = dpdf.disc ? 0.f : dpdf.v;
    if (dv == 0.f) return 0.f;
    const V3 d = next.p - p;
    const float d2 = length2(d);
    if (d2 == 0.f) return WT_INF;
    float pp = dv * (1.f / d2);
    if (bv_on_surface(c, next)) pp *= fabsf(dot(bv_ng(c, next), normalize(d)));
    return pp;
}
WT_D float sensor_pdf_direction(const BCtx& c, V3 dir) {         // virtual_plane_sensor.cpp:182-186 / perspective.hpp:326-333
    const wtgpu_sensor& s = c.sc->sensor;
    if (s.type == WTGPU_SENSOR_VIRTUAL_PLANE) return cosine_hemisphere_pdf(fmaxf(dot(dir, mk3(s.frame_n)), 0.f));
    const V3 d = normalize(m3mul(s.inv_rot, dir));
    return d.z > 1.1920929e-7f ? 1.f / persp_recp_sa(s, d) : 0.f;
}
WT_D float sensor_pdf_position(const BCtx& c) { const wtgpu_sensor& s = c.sc->sensor; return s.type == WTGPU_SENSOR_VIRTUAL_PLANE ? 1.f / (s.extent[0] * s.extent[1]) : 0.f; }
WT_D float pdf_next_from_sensor(const BCtx& c, const BVertex& v, const BVertex& next) {   // vertex.hpp:489-506
    const V3 dl = next.p - v.p;
    const float rd2 = 1.f / length2(dl);
    const V3 d = dl * sqrtf(rd2);
    float pp = sensor_pdf_direction(c, d) * rd2;
    if (bv_on_surface(c, next)) pp *= fabsf(dot(bv_ng(c, next), d));
    return pp;
}
WT_D float emitter_pdf_position_density(const BCtx& c, int32_t e) { const wtgpu_emitter E = c.sc->emitters[e]; return E.type == WTGPU_EMITTER_AREA ? 1.f / c.sc->shapes[E.shape].surface_area : 0.f; }
WT_D float pdf_next_from_emitter(const BCtx& c, const BVertex& v, const BVertex& next) {  // vertex.hpp:522-547
    const V3 dl = next.p - v.p;
    const float rd2 = 1.f / length2(dl);
    const V3 d = dl * sqrtf(rd2);
    const int32_t em = bv_emitter(c, v);
    const wtgpu_emitter E = c.sc->emitters[em];
    if (E.type == WTGPU_EMITTER_DIRECTIONAL) {
        const Frame fr = orthogonal_frame(mk3(E.dir));
        const V3 pl = to_local(fr, next.p - mk3(E.world_centre));
        return length2(mk2(pl.x, pl.y)) <= sqrf(E.world_radius) ? 1.f / (kPi * sqrf(E.world_radius)) : 0.f;
    }
    float dd;
    if (E.type == WTGPU_EMITTER_POINT) dd = kInvFourPi;
    else if (E.type == WTGPU_EMITTER_SPOT) dd = 1.f / (kTwoPi * (1.f - pm::cosf(E.cutoff)));
    else dd = cosine_hemisphere_pdf(fmaxf(0.f, dot(d, bv_ng(c, v))));
    float pp = dd * rd2;
    if (bv_on_surface(c, next)) pp *= fabsf(dot(bv_ng(c, next), d));
    return pp;
}
WT_D float pdf_emitter_v(const BCtx& c, const BVertex& v) {       // vertex.hpp:549-564
    const int32_t em = bv_emitter(c, v);
    if (c.sc->emitters[em].type == WTGPU_EMITTER_DIRECTIONAL) return 0.f;
    return pdf_emitter(*c.sc, em) * emitter_pdf_position_density(c, em);
}
WT_NI float bv_pdf(const BCtx& c, const BVertex& v, const BVertex* prev, const BVertex& next, bool mode_fwd) {   // vertex.hpp:444-487
    if (v.type == BV_EMITTER) return pdf_next_from_emitter(c, v, next);
    if (v.type == BV_SENSOR) return pdf_next_from_sensor(c, v, next);
    const V3 wiw = normalize(prev->p - v.p), wow = normalize(next.p - v.p);
    Pd pdf = pd_disc(0.f);
    if (v.type == BV_SURFACE) {
        const Surface srf = bv_surface(*c.sc, v);
        BsdfQuery q; q.k = v.beam.k; q.fwd = mode_fwd; q.lobes = 0xffffffffu;
        pdf = pd_dens(bsdf_pdf(*c.sc, v.bsdf, to_local(srf.shading, wiw), to_local(srf.shading, wow), q));
    } else {
        if (!v.ffsd) return 0.f;
        const FHead h = ap_head(c.A, v.fsd);
        pdf = pd_dens(fraunhofer_pdf(c.A, v.fsd, h, to_local(h.frame, wow)));
    }
    return dir_to_area(c, pdf, v.p, next);
}
WT_D float snc_scale(bool fwd, float wig, float wog, float wis, float wos) { return fwd ? fminf(fabsf(wis * wog / (wos * wig)), 1e+2f) : 1.f; }   // integrator/common.hpp:22-33
// vertex_t::interact (vertex.hpp:330-413)
WT_NI bool bv_interact(const BCtx& c, const BVertex& v, V3 next_p, bool ignore_fsd, Beam& out) {
    const DScene& sc = *c.sc;
    const V3 wiw = -v.beam.env.d;
    float f = 0.f;
    if (v.ffsd && !ignore_fsd) { const FHead h = ap_head(c.A, v.fsd); f = fraunhofer_pdf(c.A, v.fsd, h, to_local(h.frame, normalize(next_p - v.p))); }
    if (v.type == BV_SURFACE) {
        const Surface srf = bv_surface(sc, v);
        const V3 wow = normalize(next_p - v.p);
        const V3 wi = to_local(srf.shading, wiw), wo = to_local(srf.shading, wow);
        const V3 ng = srf.geo.n, ns = srf.shading.n;
        const float wig = dot(wiw, ng), wog = dot(wow, ng);
        if (wig * wi.z <= 0.f || wog * wo.z <= 0.f) return false;
        BsdfQuery q; q.k = v.beam.k; q.fwd = v.fwd != 0u; q.lobes = 0xffffffffu;
        Mueller fb = bsdf_f(sc, v.bsdf, wi, wo, q);
        float scale = 1.f / fabsf(wo.z);
        if (!veq(ns, ng)) scale *= snc_scale(v.fwd != 0u, wig, wog, wi.z, wo.z);
        fb = mu_scale(fb, scale);
        if (f > 0.f) fb = mu_add(fb, mu_scale(mu_identity(), f));
        if (fb.m[0] == 0.f) return false;
        out = v.beam;
        beam_transform_surface(out, srf, wow, fb, 1.f);
        return true;
    }
    if (v.type == BV_FSD) {
        out = v.beam;
        beam_transform_region(out, v.p, dot(v.p - v.beam.env.o, v.beam.env.d), normalize(next_p - v.p), f);
        return true;
    }
    return false;
}

// ================================================================================================ walk
struct BWalk { Beam beam; bool fwd; Pd pdf_from_prev; float throughput, rr; uint32_t base, n; Geo prev_geo; uint32_t ap0, n_ap; };   // vertices at [base, base+n)

WT_D bool bd_append(const BCtx& c, BWalk& d, BVertex& v, Pd pdf_fwd, Pd pdf_revr) {     // plt_bdpt_detail.hpp:96-122
    const BVertex& prev = bv_ref(c.A, d.base + d.n - 1u);
    if (veq(prev.p, v.p)) return false;
    const float pa = dir_to_area(c, d.pdf_from_prev, prev.p, v);
    if (v.fwd) v.pdf_fwd = pa; else v.pdf_bwd = pa;
    v.beam = d.beam;
    const float pr = dir_to_area(c, pdf_revr, v.p, prev);
    const bool prev_fwd = prev.fwd != 0u;
    // prev.pdf_reversed(): transport backward -> pdf_fwd, forward -> pdf_bwd
    aw(c.A, (d.base + d.n - 1u) * kVertWords + (prev_fwd ? kOffPdfBwd : kOffPdfFwd)) = pr;
    d.pdf_from_prev = pdf_fwd;
    bv_store(c.A, d.base + d.n, v);
    d.n++;
    d.prev_geo = bv_geo(v);
    return true;
}

// What one traverse() found, reduced to what the vertex step needs (random_walk :421-470 + find_closest_triangle :362-419)
struct BHit { bool empty, ballistic, overflow; uint32_t primary; float pdist, bx, by, dist, region_depth, flux; V3 origin; uint32_t n_edges, need_edges; };
WT_D void bd_hit_init(BHit& h, const Beam& beam, const TravOut& tr, Range& zr) {
    h.empty = tr.empty; h.overflow = false; h.primary = WTGPU_INVALID_IDX; h.pdist = WT_INF; h.bx = h.by = -1.f; h.flux = 0.f; h.n_edges = 0u; h.need_edges = 0u;
    h.origin = tr.origin; h.region_depth = tr.region_depth; h.dist = 0.f; h.ballistic = true;
    zr = mkr(0.f, 0.f);
    if (tr.empty) return;
    if (tr.cone.overflow) h.overflow = true;
    h.dist = tr.ballistic ? tr.ray.dist : tr.cone.dist;
    zr = mkr(h.dist, h.dist + tr.region_depth);
    h.ballistic = tr.ballistic || cone_is_ray(beam.env);
    if (h.ballistic) { h.primary = tr.ray.tuid; h.pdist = tr.ray.dist; h.bx = tr.ray.bx; h.by = tr.ray.by; }
}
// one thread, start to finish (the cross-check drivers)
WT_D void bd_resolve_hit(const DScene& sc, const Beam& beam, const TravOut& tr, const TriList& tl, uint32_t* edges, BHit& h) {
    Range zr; bd_hit_init(h, beam, tr, zr);
    if (tr.empty || h.ballistic) return;
    const V3 dir = beam.env.d;
    find_closest(sc, tl, tr.origin, dir, zr, h.primary, h.pdist, h.bx, h.by);
    if (h.primary != WTGPU_INVALID_IDX) return;
    const Frame beam_frame = cone_frame(beam.env);
    const G2 wf = wavefront_of(beam, h.dist);
    const float csz = (zr.mx + zr.mn) / 2.f;
    for (uint32_t i = 0; i < tl.n; ++i) {
        const Tri3 t = load_tri(sc, tri_at(tl, i));
        if ((dot(t.n, -dir) > 0.f) != tr.cone.front) continue;
        const Clip cl = clip_triangle_z(to_local(beam_frame, t.a - beam.env.o), to_local(beam_frame, t.b - beam.env.o), to_local(beam_frame, t.c - beam.env.o), zr);
        for (int k = 0; k < cl.tris; ++k) {
            V3 ct[3]; clip_tri(cl, k, ct);
            h.flux += g2_integrate_triangle(sc, wf, cone_project_local(beam.env, ct[0], csz), cone_project_local(beam.env, ct[1], csz), cone_project_local(beam.env, ct[2], csz));
        }
    }
    if (sc.integrator.fsd) { bool eo = false; h.n_edges = collect_edges(sc, tl, edges, sc.cap.edges, eo); if (eo) { h.overflow = true; h.need_edges = 3u * tl.n; } }
}

// bd_resolve_hit with the flux integral organised for a warp (all 32 lanes call; `act` marks lanes that hold a walker): every lane walks its own
// clipped triangles, but the lanes advance piece by piece together, and a piece that needs the quadrature branch of integrate_triangle is
// evaluated by the whole warp (g2_quadrature_warp).  Same values, same order of additions as bd_resolve_hit.
WT_D void bd_resolve_hit_warp(const DScene& sc, bool act, const Beam& beam, const TravOut& tr, const TriList& tl, uint32_t* edges, BHit& h) {
    const unsigned lane = threadIdx.x & 31u;
    bool need = false;
    Range zr = mkr(0.f, 0.f);
    V3 dir = mk3(0.f, 0.f, 1.f);
    if (act) {
        bd_hit_init(h, beam, tr, zr);
        if (!tr.empty && !h.ballistic) {
            dir = beam.env.d;
            find_closest(sc, tl, tr.origin, dir, zr, h.primary, h.pdist, h.bx, h.by);
            need = h.primary == WTGPU_INVALID_IDX;
        }
    }
    const bool flux_lane = need;
    if (__any_sync(0xffffffffu, need)) {
        Frame beam_frame; G2 wf; float csz = 0.f;
        if (need) { beam_frame = cone_frame(beam.env); wf = wavefront_of(beam, h.dist); csz = (zr.mx + zr.mn) / 2.f; }
        uint32_t i = 0u; int k = 0; Clip cl; cl.tris = 0;
        while (__any_sync(0xffffffffu, need)) {
            int kind = G2_DONE; float val = 0.f; bool add = false;
            V2 pa = mk2(0.f, 0.f), pb = pa, pc = pa;
            if (need) {
                while (k >= cl.tris) {        // next triangle facing like the closest one (plt_bdpt_detail.hpp:391-416)
                    if (i >= tl.n) { need = false; break; }
                    const Tri3 t = load_tri(sc, tri_at(tl, i)); ++i;
                    if ((dot(t.n, -dir) > 0.f) != tr.cone.front) continue;
                    cl = clip_triangle_z(to_local(beam_frame, t.a - beam.env.o), to_local(beam_frame, t.b - beam.env.o), to_local(beam_frame, t.c - beam.env.o), zr);
                    k = 0;
                }
                if (need) {
                    V3 ct[3]; clip_tri(cl, k, ct); ++k;
                    pa = cone_project_local(beam.env, ct[0], csz); pb = cone_project_local(beam.env, ct[1], csz); pc = cone_project_local(beam.env, ct[2], csz);
                    kind = g2_classify(wf, pa, pb, pc, val);
                    if (kind == G2_ANALYTIC) val = g2_analytic(sc, pa, pb, pc);
                    add = true;
                }
            }
            { const float qv = g2_quadrature_mixed(kind == G2_QUADRATURE, pa, pb, pc); if (kind == G2_QUADRATURE) val = qv; }
            if (add) h.flux += val;
        }
    }
    if (flux_lane && sc.integrator.fsd) { bool eo = false; h.n_edges = collect_edges(sc, tl, edges, sc.cap.edges, eo); if (eo) { h.overflow = true; h.need_edges = 3u * tl.n; } }
}

// ---- long triangle lists: one WARP per list (k_bd_resolve_big), and the Gaussian power of the lists that need it -- no triangle on the beam's
// central ray -- as flat 32-triangle tasks over the whole GPU (k_bd_flux_chunks) whose piece values are then added IN LIST ORDER per list
// (k_bd_flux_finish).  Under a beam much wider than the mesh's triangles a list holds 10^3-10^5 entries and nearly every clipped piece takes the
// quadrature branch of integrate_triangle (~10^2 samples each, some 10^4): done inside one warp per list, a handful of lists decided the
// duration of the launch.  Same values, same order of additions as bd_resolve_hit.

// the clipped pieces of list entries [base, base + 32), one entry per lane (all 32 lanes call): values of pieces 0..2 and their number
WT_D void bd_flux_chunk(const DScene& sc, const Beam& beam, bool front, const TriList& tl, Range zr, const Frame& beam_frame, const G2& wf, float csz, uint32_t base,
                        float& v0, float& v1, float& v2, int& cnt, const QuadQ* q = nullptr, size_t dst = 0) {
    const unsigned lane = threadIdx.x & 31u;
    const V3 dir = beam.env.d;
    const uint32_t i = base + lane;
    Clip cl; cl.tris = 0;
    if (i < tl.n) {
        const Tri3 t = load_tri(sc, tri_at(tl, i));
        if ((dot(t.n, -dir) > 0.f) == front)
            cl = clip_triangle_z(to_local(beam_frame, t.a - beam.env.o), to_local(beam_frame, t.b - beam.env.o), to_local(beam_frame, t.c - beam.env.o), zr);
    }
    v0 = v1 = v2 = 0.f;
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {       // piece k of every lane's triangle
        int kind = G2_DONE; float val = 0.f; V2 pa = mk2(0.f, 0.f), pb = pa, pc = pa;
        if (k < cl.tris) {
            V3 ct[3]; clip_tri(cl, k, ct);
            pa = cone_project_local(beam.env, ct[0], csz); pb = cone_project_local(beam.env, ct[1], csz); pc = cone_project_local(beam.env, ct[2], csz);
            kind = g2_classify(wf, pa, pb, pc, val);
            if (kind == G2_ANALYTIC) val = g2_analytic(sc, pa, pb, pc);
        }
        { const float qv = g2_quadrature_mixed(kind == G2_QUADRATURE, pa, pb, pc, q, dst + (size_t)k); if (kind == G2_QUADRATURE) val = qv; }
        if (k == 0) v0 = val; else if (k == 1) v1 = val; else v2 = val;
    }
    cnt = cl.tris;
}
// ordered accumulation of one chunk: entry base, base + 1, ...; pieces 0, 1, 2 of each (all 32 lanes call; every lane returns the same sum)
WT_D float bd_flux_add(float flux, float v0, float v1, float v2, int cnt) {
    const unsigned FULL = 0xffffffffu;
    unsigned rest = __ballot_sync(FULL, cnt > 0);
    while (rest) {
        const int l = __ffs(rest) - 1; rest &= rest - 1u;
        const int c = __shfl_sync(FULL, cnt, l);
        const float a0 = __shfl_sync(FULL, v0, l), a1 = __shfl_sync(FULL, v1, l), a2 = __shfl_sync(FULL, v2, l);
        flux += a0; if (c > 1) flux += a1; if (c > 2) flux += a2;
    }
    return flux;
}
// One warp resolves one walker (all lanes call with the same arguments; every lane ends with the same BHit).  First half: the closest triangle, found
// by the flat search; returns true when there is none, i.e. the Gaussian power (and the edge set) are still to come.
WT_D bool bd_resolve_closest_big(const DScene& sc, const Beam& beam, const TravOut& tr, const TriList& tl, unsigned long long best, BHit& h) {
    Range zr; bd_hit_init(h, beam, tr, zr);
    if (tr.empty || h.ballistic) return false;
    if (best == ~0ull) return true;
    // the winner of the flat search (k_bd_closest_chunks): its triangle once more, for the barycentrics
    const uint32_t e = (uint32_t)best, tu = tri_at(tl, e);
    const Tri3 t = load_tri(sc, tu);
    const float tol = cone_intersection_tolerance(tr.origin, t.a, t.b, t.c);
    const RayTri rt = intersect_ray_tri(tr.origin, beam.env.d, t.a, t.b, t.c, mkr(zr.mn - tol, zr.mx + tol));
    h.primary = tu; h.pdist = rt.dist; h.bx = rt.bx; h.by = rt.by;
    return false;
}
// Second half, inside the warp (lists too short to be worth queueing, or no scratch left): the power chunk by chunk, the edge set through the scratch bitmap
WT_D void bd_resolve_flux_big(const DScene& sc, const Beam& beam, const TravOut& tr, const TriList& tl, uint32_t* edges, uint32_t* edge_bits, BHit& h) {
    const Range zr = mkr(h.dist, h.dist + tr.region_depth);
    const Frame beam_frame = cone_frame(beam.env);
    const G2 wf = wavefront_of(beam, h.dist);
    const float csz = (zr.mx + zr.mn) / 2.f;
    float flux = 0.f;
    for (uint32_t base = 0u; base < tl.n; base += 32u) {
        float v0, v1, v2; int cnt;
        bd_flux_chunk(sc, beam, tr.cone.front, tl, zr, beam_frame, wf, csz, base, v0, v1, v2, cnt);
        flux = bd_flux_add(flux, v0, v1, v2, cnt);
    }
    h.flux = flux;
    if (sc.integrator.fsd) { bool eo = false; uint32_t need = 0u; h.n_edges = w_collect_edges(sc, tl, edges, sc.cap.edges, edge_bits, eo, need); if (eo) { h.overflow = true; h.need_edges = need; } }
}

// continue_walk (plt_bdpt_detail.hpp:167-182)
WT_D bool bd_continue_walk(BCtx& c, BWalk& data, Sampler& smp, bool do_RR) {
    const DScene& sc = *c.sc;
    if (data.n > sc.integrator.max_depth + 1u) return false;
    if (data.n >= sc.cap.verts) { c.overflow = true; c.need_verts = max(c.need_verts, 2u * sc.cap.verts); return false; }      // the subpath's vertex row is full (capacity growth)
    if (do_RR && sc.integrator.russian_roulette) {
        aw(c.A, (data.base + data.n - 1u) * kVertWords + kOffRr) = data.rr;
        const float r = data.throughput < 1.f ? fmaxf(data.throughput, .5f) : 1.f;
        if (rnd(smp) <= r) { const float s = 1.f / r; data.rr *= s; data.throughput *= s; }
        else return false;
    }
    return true;
}
// second half of sample_fraunhofer_fsd_interaction (:318-346): the sampled direction becomes an FSD vertex
WT_D bool bd_walk_fsd_finish(BCtx& c, BWalk& data, Sampler& smp, int ai, V3 interaction_wp, float beam_dist, V3 wo, float dpd, float wgt, uint32_t& n_vert) {
    if (dpd == 0.f || wgt == 0.f) return false;
    const V3 wow = to_world(cone_frame(data.beam.env), wo);
    BVertex v; v.type = BV_FSD; v.fwd = data.fwd; v.delta = 0u; v.ffsd = 1u; v.pdf_fwd = v.pdf_bwd = -1.f; v.rr = 1.f;
    v.gkind = BG_POINT; v.p = interaction_wp; v.tuid = WTGPU_INVALID_IDX; v.bary = mk2(0.f, 0.f); v.fp.x = mk2(1.f, 0.f); v.fp.la = v.fp.lb = 0.f; v.dn = mk3(0.f, 0.f, 1.f);
    v.emitter = -1; v.bsdf = -1; v.fsd = ai; v.pad_ = 0u;
    if (!bd_append(c, data, v, pd_dens(dpd), pd_dens(dpd))) return false;
    ++n_vert;
    beam_transform_region(data.beam, interaction_wp, beam_dist, wow, wgt);
    data.throughput *= wgt;
    return bd_continue_walk(c, data, smp, true);
}
enum : int { BD_END = 0, BD_CONTINUE = 1, BD_FSD_DEFERRED = 2 };
// One vertex of a subpath: the body of plt_bdpt::random_walk (plt_bdpt_detail.hpp:470-526) after the traverse.
// defer_fsd: stop after the aperture is built (BD_FSD_DEFERRED, aperture index data.ap0 + data.n_ap - 1) so that the caller can run
// the rejection sampler elsewhere and resume with bd_walk_fsd_finish.
WT_NI int bd_walk_step(BCtx& c, BWalk& data, Sampler& smp, const BHit& h, const uint32_t* edges, uint32_t& n_vert, bool defer_fsd) {
    const DScene& sc = *c.sc;
    if (h.empty) return BD_END;
    if (h.overflow) c.overflow = true;
    Beam& beam = data.beam;
    const float beam_dist = h.dist;
    const V3 dir = beam.env.d;
    const V3 interaction_wp = h.origin + beam_dist * dir;
    const Frame beam_frame = cone_frame(beam.env);
    bool do_RR = true;
    if (h.primary != WTGPU_INVALID_IDX) {     // sample_surface_interaction (:193-270)
        Surface srf = make_surface(sc, h.primary, mk2(h.bx, h.by), h.origin + dir * h.pdist);
        srf.fp = surface_footprint_static(beam, srf, beam_dist);
        const int32_t bsdf = sc.shapes[sc.tri_meta[h.primary].shape_idx].bsdf;
        const V3 ng = srf.geo.n, ns = srf.shading.n;
        const V3 wiw = -dir;
        const V3 wi = to_local(srf.shading, wiw);
        const float wig = dot(wiw, ng);
        if (wig * wi.z <= 0.f) return BD_END;
        BsdfQuery q; q.k = beam.k; q.fwd = data.fwd; q.lobes = 0xffffffffu;
        const BsdfSample bs = bsdf_sample(sc, bsdf, wi, q, smp);
        if (!bs.valid || bs.dpd.v == 0.f) return BD_END;
        const V3 wow = normalize(to_world(srf.shading, bs.wo));
        const float wog = dot(wow, ng);
        if (wog * bs.wo.z <= 0.f) return BD_END;
        BsdfQuery qr = q; qr.fwd = !data.fwd;
        const Pd pdf_revr = pd_dens(bsdf_pdf(sc, bsdf, bs.wo, wi, qr));
        BVertex v; v.type = BV_SURFACE; v.fwd = data.fwd; v.delta = bs.dpd.disc; v.ffsd = 0u; v.pdf_fwd = v.pdf_bwd = -1.f; v.rr = 1.f;
        v.gkind = BG_SURFACE; v.p = srf.wp; v.tuid = h.primary; v.bary = mk2(h.bx, h.by); v.fp = srf.fp; v.dn = mk3(0.f, 0.f, 1.f); v.emitter = -1; v.bsdf = bsdf; v.fsd = -1; v.pad_ = 0u;
        if (!bd_append(c, data, v, bs.dpd, pdf_revr)) return BD_END;
        ++n_vert;
        float w = 1.f;
        if (!veq(ns, ng)) w *= snc_scale(data.fwd, wig, wog, wi.z, bs.wo.z);
        beam_transform_surface(data.beam, srf, wow, bs.M, w);
        data.throughput *= w * bs.M.m[0];
        if (!data.fwd && bs.eta.re != 1.f) data.throughput /= sqrf(bs.eta.re);
    } else if (!h.ballistic && sc.integrator.fsd && h.n_edges) {       // sample_fraunhofer_fsd_interaction (:288-346)
        if (data.n_ap >= sc.cap.ap_walk) { c.overflow = true; c.need_ap = max(c.need_ap, sc.cap.ap_walk + 2u); return BD_END; }      // (how many more the walk would need is not known: two more)
        const int ai = (int)(data.ap0 + data.n_ap);
        const G2 wf = wavefront_of(beam, beam_dist);
        const uint32_t nseg = fraunhofer_build(sc, c.A, ai, beam_frame, beam.k, 1.f - h.flux, beam.env, edges, h.n_edges, wf, c.overflow, c.need_seg);
        if (nseg == 0u) { beam_transform_restart(data.beam, interaction_wp, beam_dist); do_RR = false; }
        else {
            ++data.n_ap;
            if (defer_fsd) return BD_FSD_DEFERRED;
            V3 wo; float dpd, wgt;
            fraunhofer_sample(c.A, ai, c.lut, smp, wo, dpd, wgt);
            return bd_walk_fsd_finish(c, data, smp, ai, interaction_wp, beam_dist, wo, dpd, wgt, n_vert) ? BD_CONTINUE : BD_END;
        }
    } else { do_RR = false; beam_transform_restart(data.beam, interaction_wp, beam_dist); }
    return bd_continue_walk(c, data, smp, do_RR) ? BD_CONTINUE : BD_END;
}
// key for the material sort of subpath walkers: bsdf id | fsd | null | miss
WT_D uint32_t bd_hit_key(const DScene& sc, const BHit& h, uint32_t n_keys) {
    if (h.empty) return n_keys - 1u;
    if (h.primary != WTGPU_INVALID_IDX) return (uint32_t)sc.shapes[sc.tri_meta[h.primary].shape_idx].bsdf;
    return (!h.ballistic && sc.integrator.fsd && h.n_edges) ? n_keys - 3u : n_keys - 2u;
}

// ================================================================================================ connections + MIS
struct BConn { BVertex tmp; bool has_el; Element el; Stokes L; };

WT_D Stokes connect_and_integrate(const BCtx& c, const Beam& db, const Geo& dg, const Beam& eb, const Geo& eg) {   // plt_bdpt_detail.hpp:725-745
    if (beam_intensity(db) == 0.f || beam_intensity(eb) == 0.f) return stokes_zero();
    if (shadow_between(*c.sc, dg, eg, *c.ctr)) return stokes_zero();
    return integrate_beams(db, eb);
}
WT_D void tmp_init(BVertex& t, uint32_t type) {
    t.type = type; t.fwd = type == BV_EMITTER ? 1u : 0u; t.delta = 0u; t.ffsd = 0u; t.pdf_fwd = -1.f; t.pdf_bwd = -1.f; t.rr = 1.f;
    t.gkind = BG_POINT; t.p = mk3(0.f, 0.f, 0.f); t.tuid = WTGPU_INVALID_IDX; t.bary = mk2(0.f, 0.f); t.fp.x = mk2(1.f, 0.f); t.fp.la = t.fp.lb = 0.f; t.dn = mk3(0.f, 0.f, 1.f);
    t.emitter = -1; t.bsdf = -1; t.fsd = -1; t.pad_ = 0u;
}
// CLS = strategy class of (s,t) (bd_pair_class: the branch taken below), or -1 to choose at run time.  The wavefront driver launches one
// kernel per class; compiling only that class's branch into it keeps each kernel's code small (these kernels stall on instruction fetch).
template <int CLS>
WT_NI void bd_connect(BCtx& c, uint32_t nsv, uint32_t nev, int s, int t, Sampler& smp, BConn& ret) {       // plt_bdpt_detail.hpp:747-923
    const DScene& sc = *c.sc;
    const uint32_t SB = 0u, EB = sc.cap.verts;
    ret.has_el = false; ret.L = stokes_zero(); tmp_init(ret.tmp, BV_SENSOR);
    const bool virt = sc.sensor.type == WTGPU_SENSOR_VIRTUAL_PLANE;
    if ((CLS < 0 && s == 0) || CLS == 0) {
        const BVertex& last = bv_ref(c.A, SB + t - 1);
        if (bv_on_emitter(c, last)) {
            Beam QE = last.beam; beam_mul(QE, last.rr);
            if (last.type == BV_SURFACE) { const Surface srf = bv_surface(sc, last); ret.L = emitter_Li(sc, bv_emitter(c, last), QE, srf); }
        }
    } else if ((CLS < 0 && t == 0) || CLS == 1) {
        if (virt) {
            const BVertex& last = bv_ref(c.A, EB + s - 1); const BVertex& cur = bv_ref(c.A, EB + s - 2);
            const Beam& beam = last.beam;
            Beam db; Element el;
            if (sensor_Si(sc, beam, mkr(0.f, length(last.p - beam.env.o)), db, el)) {
                ret.has_el = true; ret.el = el;
                float w = cur.rr;
                if (bv_on_surface(c, cur) && bv_nondelta_interaction(cur)) w /= fabsf(dot(db.env.d, bv_ns(c, cur)));
                w /= fabsf(dot(db.env.d, mk3(sc.sensor.frame_n)));
                beam_mul(db, w);
                tmp_init(ret.tmp, BV_SENSOR); ret.tmp.pdf_bwd = 0.f; ret.tmp.gkind = BG_DUMMY; ret.tmp.p = db.env.o; ret.tmp.dn = mk3(sc.sensor.frame_n);
                ret.L = integrate_beams(db, beam);
            }
        }
    } else if ((CLS < 0 && s == 1) || CLS == 2) {
        const BVertex& last = bv_ref(c.A, SB + t - 1);
        if (bv_connectible(c, last)) {
            EmitterDirect ed = scene_sample_emitter_direct(sc, smp, last.p, last.beam.k);
            if ((ed.dpd.disc || ed.dpd.v != 0.f) && beam_intensity(ed.beam) > 0.f) {
                float w = last.rr;
                if (bv_on_surface(c, last)) w *= fabsf(dot(ed.beam.env.d, bv_ns(c, last)));
                beam_mul(ed.beam, w);
                tmp_init(ret.tmp, BV_EMITTER); ret.tmp.emitter = ed.emitter;
                if (ed.has_surface) { ret.tmp.gkind = BG_SURFACE; ret.tmp.p = ed.sp; ret.tmp.tuid = ed.stuid; }
                else { ret.tmp.gkind = BG_POINT; ret.tmp.p = ed.beam.env.o; }
                Beam db;
                if (bv_interact(c, last, ret.tmp.p, false, db)) ret.L = connect_and_integrate(c, db, bv_geo(last), ed.beam, bv_geo(ret.tmp));
            }
        }
    } else if ((CLS < 0 && t == 1) || CLS == 3) {
        const BVertex& last = bv_ref(c.A, EB + s - 1);
        if ((virt || last.type != BV_FSD) && bv_connectible(c, last)) {
            SensorDirect sd = sensor_sample_direct(sc, smp, last.p, last.beam.k);
            if ((sd.dpd.disc || sd.dpd.v != 0.f) && beam_intensity(sd.beam) > 0.f) {
                float w = last.rr;
                if (bv_on_surface(c, last)) w *= fabsf(dot(sd.beam.env.d, bv_ns(c, last)));
                beam_mul(sd.beam, w);
                tmp_init(ret.tmp, BV_SENSOR); ret.tmp.p = sd.beam.env.o;
                if (virt) { ret.tmp.gkind = BG_DUMMY; ret.tmp.dn = mk3(sc.sensor.frame_n); }
                Beam eb;
                if (bv_interact(c, last, ret.tmp.p, false, eb)) { ret.L = connect_and_integrate(c, sd.beam, bv_geo(ret.tmp), eb, bv_geo(last)); ret.has_el = true; ret.el = sd.el; }
            }
        }
    } else {
        const BVertex& ev = bv_ref(c.A, EB + s - 1); const BVertex& sv = bv_ref(c.A, SB + t - 1);
        const V3 dl = ev.p - sv.p;
        if (bv_connectible(c, ev) && bv_connectible(c, sv) && !(dl.x == 0.f && dl.y == 0.f && dl.z == 0.f)) {
            Beam eb, db;
            if (bv_interact(c, ev, sv.p, true, eb) && bv_interact(c, sv, ev.p, true, db)) {
                const float rd2 = 1.f / length2(dl);
                const V3 d = dl * sqrtf(rd2);
                float wev = ev.rr, wsv = sv.rr * rd2;
                if (bv_on_surface(c, sv)) wev *= fabsf(dot(bv_ns(c, sv), d));
                if (bv_on_surface(c, ev)) wsv *= fabsf(dot(bv_ns(c, ev), d));
                beam_mul(db, wsv); beam_mul(eb, wev);
                ret.L = connect_and_integrate(c, db, bv_geo(sv), eb, bv_geo(ev));
            }
        }
    }
}

WT_D float area_or_one(float p) { return (isfinite(p) && p > 1.1920929e-7f) ? p : 1.f; }
WT_NI float bd_mis(BCtx& c, int s, int t, const BConn& cr) {       // plt_bdpt_detail.hpp:604-720
    if (s + t <= 2) return 1.f;
    const DScene& sc = *c.sc;
    const uint32_t SB = 0u, EB = sc.cap.verts;
    // The reference copies the subpaths' area pdfs into scratch vectors and overwrites the entries the connection changes (:621-690): the
    // reverse pdfs of the last two vertices of each subpath, and vertex 0's forward pdf when it is the sampled endpoint.  Here the overrides
    // are kept as scalars and every other entry is read from the vertex records where it is used: no per-thread arrays, any path length.
    int sr_i1 = -1, sr_i2 = -1, er_i1 = -1, er_i2 = -1;
    float sr_v1 = 0.f, sr_v2 = 0.f, er_v1 = 0.f, er_v2 = 0.f, sp0 = 0.f, ep0 = 0.f;
    bool sp0_set = false, ep0_set = false;
    const BVertex& tmp = cr.tmp;
    const bool virt = sc.sensor.type == WTGPU_SENSOR_VIRTUAL_PLANE;
    if (s == 0) {
        const BVertex& last = bv_ref(c.A, SB + t - 1); const BVertex& prev = bv_ref(c.A, SB + t - 2);
        sr_i1 = t - 1; sr_v1 = pdf_emitter_v(c, last);
        sr_i2 = t - 2; sr_v2 = pdf_next_from_emitter(c, last, prev);
    } else if (t == 0) {
        const BVertex& prev = bv_ref(c.A, EB + s - 2);
        const BVertex& last = virt ? tmp : bv_ref(c.A, EB + s - 1);
        er_i1 = s - 1; er_v1 = sensor_pdf_position(c);
        er_i2 = s - 2; er_v2 = pdf_next_from_sensor(c, last, prev);
    } else if (s == 1) {
        const BVertex& last = bv_ref(c.A, SB + t - 1); const BVertex& lp = bv_ref(c.A, SB + t - 2);
        sr_i1 = t - 1; sr_v1 = pdf_next_from_emitter(c, tmp, last);
        er_i1 = 0; er_v1 = bv_pdf(c, last, &lp, tmp, false);
        ep0_set = true; ep0 = pdf_emitter_v(c, tmp);
    } else if (t == 1) {
        const BVertex& last = bv_ref(c.A, EB + s - 1); const BVertex& lp = bv_ref(c.A, EB + s - 2);
        er_i1 = s - 1; er_v1 = pdf_next_from_sensor(c, tmp, last);
        sr_i1 = 0; sr_v1 = bv_pdf(c, last, &lp, tmp, true);
        sp0_set = true; sp0 = sensor_pdf_position(c);
    } else {
        const BVertex& e = bv_ref(c.A, EB + s - 1); const BVertex& sv = bv_ref(c.A, SB + t - 1); const BVertex& epv = bv_ref(c.A, EB + s - 2); const BVertex& spv = bv_ref(c.A, SB + t - 2);
        er_i1 = s - 1; er_v1 = bv_pdf(c, sv, &spv, e, false);
        er_i2 = s - 2; er_v2 = bv_pdf(c, e, &sv, epv, false);
        sr_i1 = t - 1; sr_v1 = bv_pdf(c, e, &epv, sv, true);
        sr_i2 = t - 2; sr_v2 = bv_pdf(c, sv, &e, spv, true);
    }
    auto word = [&](uint32_t v, uint32_t off) { return aw(c.A, v * kVertWords + off); };
    auto sp_pdf = [&](int i) { return (i == 0 && sp0_set) ? sp0 : word(SB + i, kOffPdfBwd); };
    auto sp_rev = [&](int i) { return i == sr_i1 ? sr_v1 : i == sr_i2 ? sr_v2 : word(SB + i, kOffPdfFwd); };
    auto sp_d = [&](int i) { return i == t - 1 ? false : __float_as_uint(word(SB + i, kOffDelta)) != 0u; };
    auto ep_pdf = [&](int i) { return (i == 0 && ep0_set) ? ep0 : word(EB + i, kOffPdfFwd); };
    auto ep_rev = [&](int i) { return i == er_i1 ? er_v1 : i == er_i2 ? er_v2 : word(EB + i, kOffPdfBwd); };
    auto ep_d = [&](int i) { return i == s - 1 ? false : __float_as_uint(word(EB + i, kOffDelta)) != 0u; };
    bool delta_emitter = true, delta_sensor = true;
    if (s == 1) delta_emitter = bv_delta_emitter(c, tmp); else if (s > 1) delta_emitter = bv_delta_emitter(c, bv_ref(c.A, EB));
    if (t == 1) delta_sensor = bv_delta_sensor(c, tmp); else if (t > 1) delta_sensor = bv_delta_sensor(c, bv_ref(c.A, SB));
    float sum = 0.f, ri = 1.f;
    for (int i = t - 1; i >= 0; --i) { ri *= area_or_one(sp_rev(i)) / area_or_one(sp_pdf(i)); if (!sp_d(i) && !(i > 0 ? sp_d(i - 1) : delta_sensor)) sum += ri; }
    ri = 1.f;
    for (int i = s - 1; i >= 0; --i) { ri *= area_or_one(ep_rev(i)) / area_or_one(ep_pdf(i)); if (!ep_d(i) && !(i > 0 ? ep_d(i - 1) : delta_emitter)) sum += ri; }
    return 1.f / (1.f + sum);
}

// ================================================================================================ sample set-up shared by both drivers
struct BdSampleInit { uint32_t pixel, sample, ex, ey; float k, rspd, wpd_v; Element el; };
// plt_bdpt_t::integrate preamble + generate_{sensor,emitter}_subpath heads (src/integrator/plt_bdpt.cpp:54-75; plt_bdpt_detail.hpp:528-581):
// draws on sub-stream 0, writes vertex 0 of both subpaths, returns the two walk states
WT_D void bd_init_sample(const BCtx& c, uint32_t seed_lo, uint32_t seed_hi, uint32_t ex, uint32_t ey, uint32_t sample, BdSampleInit& si, BWalk& ws, BWalk& we) {
    const DScene& sc = *c.sc;
    Sampler smp; smp.k0 = seed_lo; smp.k1 = seed_hi; smp.pixel = ey * sc.sensor.width + ex; smp.sample = sample; smp.d = 0; smp.stream = sc.scene_stream;
    const int32_t em = sample_emitter(sc, smp);
    const float em_pdf = pdf_emitter(sc, em);
    const KSample ks = sample_wavenumber(sc, em, smp);
    const float k = ks.k;
    const EmitterSample es = emitter_sample(sc, em, smp, k);
    si.rspd = ks.wpd.disc ? 1.f / ks.wpd.v : 1.f / sum_spectral_pdf(sc, k);
    const SensorSample ss = sensor_sample(sc, smp, ex, ey, k);
    si.pixel = smp.pixel; si.sample = sample; si.ex = ex; si.ey = ey; si.k = k; si.wpd_v = ks.wpd.v; si.el = ss.el;
    {
        BVertex v; tmp_init(v, BV_SENSOR); v.pdf_bwd = ss.ppd.disc ? 0.f : ss.ppd.v; v.beam = ss.beam; v.p = ss.beam.env.o;
        if (ss.has_surface) { v.gkind = BG_DUMMY; v.dn = mk3(sc.sensor.frame_n); }
        bv_store(c.A, 0u, v);
        ws.beam = ss.beam; ws.fwd = false; ws.pdf_from_prev = ss.dpd; ws.throughput = 1.f; ws.rr = 1.f; ws.base = 0u; ws.n = 1u; ws.prev_geo = bv_geo(v); ws.ap0 = 0u; ws.n_ap = 0u;
    }
    {
        BVertex v; tmp_init(v, BV_EMITTER); v.pdf_fwd = (es.ppd.disc ? 0.f : es.ppd.v) * em_pdf; v.beam = es.beam; v.emitter = em; v.p = es.beam.env.o;
        if (es.has_surface) { v.gkind = BG_SURFACE; v.tuid = es.s.tuid; v.p = es.s.wp; }
        bv_store(c.A, sc.cap.verts, v);
        we.beam = es.beam; we.fwd = true; we.pdf_from_prev = es.dpd; we.throughput = 1.f; we.rr = 1.f; we.base = sc.cap.verts; we.n = 1u; we.prev_geo = bv_geo(v);
        we.ap0 = sc.cap.ap_walk; we.n_ap = 0u;
    }
}
// the (s,t) enumeration of plt_bdpt.cpp:96-110; f(s,t) for every strategy that is evaluated
template <class F> WT_D void bd_for_each_pair(const DScene& sc, uint32_t nsv, uint32_t nev, F&& f) {
    const int maxd = (int)sc.integrator.max_depth;
    for (int t = 0; t <= (int)nsv; ++t)
        for (int s = 0; s <= (int)nev; ++s) {
            const int depth = t + s - 2;
            if ((t == 1 && s == 1) || depth < 0) continue;
            if (!sc.integrator.emitter_direct && s == 1) continue;
            if (!sc.integrator.sensor_direct && t == 1) continue;
            if (depth > maxd) break;
            f(s, t);
        }
}
// one (s,t) strategy: connect, weight; returns the flux and where it goes (plt_bdpt.cpp:111-140)
template <int CLS = -1>
WT_D float bd_eval_pair(BCtx& c, const BdSampleInit& si, uint32_t seed_lo, uint32_t seed_hi, uint32_t nsv, uint32_t nev, int s, int t, BConn& cr) {
    const DScene& sc = *c.sc;
    Sampler smp; smp.k0 = seed_lo; smp.k1 = seed_hi; smp.pixel = si.pixel; smp.sample = si.sample; smp.d = 0; smp.stream = 3u + 4096u * (uint32_t)t + (uint32_t)s;
    bd_connect<CLS>(c, nsv, nev, s, t, smp, cr);
    if (cr.L.s[0] <= 0.f) return 0.f;
    const float mis = sc.integrator.mis ? bd_mis(c, s, t, cr) * si.rspd : 1.f / ((float)(s + t + 1) * si.wpd_v);
    return cr.L.s[0] * mis;
}

// ================================================================================================ driver 1: one thread per sample (cross-check path)
struct BdptArgs {
    DScene sc; FLut lut; float* arena; uint32_t P;
    uint32_t* trav_tris; uint32_t* hit_edges;      // per-thread rows (kTriRow / sc.cap.edges entries)
    DevCounters* ctr; float* film_block; float* film_light;
    uint32_t seed_lo, seed_hi, tile_x0, tile_y0, tile_w, tile_h, sample_begin;
    unsigned long long total;
};
WT_D void bd_sample_coords(unsigned long long id, uint32_t tile_x0, uint32_t tile_y0, uint32_t tile_w, uint32_t tile_h, uint32_t sample_begin, uint32_t& ex, uint32_t& ey, uint32_t& sample) {
    const uint64_t npix = (uint64_t)tile_w * tile_h;
    const uint32_t pi = (uint32_t)(id % npix), si = (uint32_t)(id / npix);
    ex = tile_x0 + pi % tile_w; ey = tile_y0 + pi / tile_w; sample = sample_begin + si;
}
WT_D void bd_flush_need(DevCounters* g, const BCtx& c) {
    if (c.need_seg) need_max(&g->need_seg, c.need_seg);
    if (c.need_ap) need_max(&g->need_ap, c.need_ap);
    if (c.need_verts) need_max(&g->need_verts, c.need_verts);
}
WT_D void bd_flush_stats(DevCounters* g, uint32_t n_splat, uint32_t n_vert, uint32_t n_conn, uint32_t n_samples, bool overflow) {
    const unsigned m = __activemask();
    const unsigned ns = __reduce_add_sync(m, n_splat), nv = __reduce_add_sync(m, n_vert), nc = __reduce_add_sync(m, n_conn), nn = __reduce_add_sync(m, n_samples);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) {
        if (ns) atomicAdd(&g->splats, (unsigned long long)ns);
        if (nv) atomicAdd(&g->segments, (unsigned long long)nv);
        if (nc) atomicAdd(&g->shaded, (unsigned long long)nc);
        if (nn) atomicAdd(&g->samples, (unsigned long long)nn);
    }
    count1(&g->overflow, overflow);
}
// plt_bdpt_t::integrate (src/integrator/plt_bdpt.cpp:43-148) start to finish in one thread.  Slow (divergent); kept as an
// independent second implementation of the same contract for the parity tests (WTGPU_RENDER_BDPT_MEGAKERNEL).
__global__ void __launch_bounds__(128) k_bdpt(const BdptArgs a) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    Counters ctr; counters_zero(ctr);
    const DScene& sc = a.sc;
    BCtx c = mk_bctx(sc, a.arena, tid, a.lut, &ctr);
    uint32_t* edges = a.hit_edges + (size_t)tid * sc.cap.edges;
    uint32_t n_samples = 0, n_vert = 0, n_conn = 0, n_splat = 0;
    const bool force_rt = sc.sensor.ray_trace_only != 0u;
    for (;;) {
        const unsigned long long id = atomicAdd(&a.ctr->next_sample, 1ull);
        if (id >= a.total) break;
        ++n_samples;
        uint32_t ex, ey, sample; bd_sample_coords(id, a.tile_x0, a.tile_y0, a.tile_w, a.tile_h, a.sample_begin, ex, ey, sample);
        BdSampleInit si; BWalk w[2];
        bd_init_sample(c, a.seed_lo, a.seed_hi, ex, ey, sample, si, w[0], w[1]);
        for (int wi = 0; wi < 2; ++wi) {
            Sampler smp; smp.k0 = a.seed_lo; smp.k1 = a.seed_hi; smp.pixel = si.pixel; smp.sample = si.sample; smp.d = 0; smp.stream = 1u + (uint32_t)wi;
            for (;;) {
                TravOut tr; BHit h;
                // (one launch: the triangle-list arena is not recycled here -- this cross-check driver is for small scenes; the arena grows to the whole render's demand)
                TriWriter tw = tri_writer(sc, a.trav_tris, tid);
                traverse(sc, w[wi].beam.env, w[wi].prev_geo, wavenum_to_wavelen(w[wi].beam.k), force_rt, tw, tr, ctr);
                bd_resolve_hit(sc, w[wi].beam, tr, tri_list(sc, a.trav_tris, tid, tr.cone.n_tris, tr.cone.overflow), edges, h);
                if (h.need_edges) need_max(&a.ctr->need_edges, h.need_edges);
                if (bd_walk_step(c, w[wi], smp, h, edges, n_vert, false) != BD_CONTINUE) break;
            }
        }
        float L0 = 0.f;
        const uint32_t nsv = w[0].n, nev = w[1].n;
        bd_for_each_pair(sc, nsv, nev, [&](int s, int t) {
            BConn cr;
            const float flux = bd_eval_pair(c, si, a.seed_lo, a.seed_hi, nsv, nev, s, t, cr);
            ++n_conn;
            if (cr.L.s[0] <= 0.f) return;
            if (t > 1) L0 += flux;
            else n_splat += film_splat(sc, a.film_block, a.film_light, true, cr.el, flux, si.k);
        });
        n_splat += film_splat(sc, a.film_block, a.film_light, false, si.el, L0, si.k);
    }
    flush_counters(a.ctr, ctr);
    bd_flush_need(a.ctr, c);
    bd_flush_stats(a.ctr, n_splat, n_vert, n_conn, n_samples, c.overflow);
}

// ================================================================================================ driver 2: wavefront
// Sample slots hold both subpaths' vertices (AoS records in the arena).  The two walks of a sample are independent "walkers"
// (id = 2 slot + which) that advance one vertex per iteration through  traverse -> material sort -> vertex step; when both have
// ended, the slot's (s,t) strategies are expanded into a task list and evaluated one strategy per thread; the last strategy to
// finish splats the sample's block contribution and frees the slot.
constexpr int kPairClasses = 5;
struct alignas(16) BdWalker { Beam beam; Geo prev_geo; float pdf_v; uint32_t pdf_disc; float throughput, rr; uint32_t n, rng_d, n_ap, pad0; };
struct alignas(16) BdHeader { uint32_t pixel, sample, ex, ey; float k, rspd, wpd_v, ox, oy; uint32_t elx, ely, pad0; };
struct BdArgs {
    RenderArgs r;               // sort buffers sized for 2P walkers; r.hit = walker hit records; r.alive = slot flags; r.pool = 2P
    FLut lut; float* arena; uint32_t P;
    float4* walkers; float4* headers;
    int* pending; float* L0; uint32_t* nverts; unsigned long long* pairs;      // pairs: strategy tasks, slot | s << 32 | t << 48
    TravRec* trav_rec; uint32_t* trav_tris;  // group traversal results (per walker), consumed by k_bd_resolve
    uint32_t* fsd_list; float4* fsd_out;    // walkers waiting for a Fraunhofer direction sample (two lists, ping-pong); its result / carried state
    size_t pair_off[kPairClasses];          // start of each strategy class's segment of `pairs`
    uint32_t fl_cur, fl_next, fl_fin;       // list fed by this iteration's vertex step and consumed by its sampler; carry-over list; list being finished
    float tag, tag_fin;                     // marks results written by this iteration's sampler / by the sampler whose list is being finished
};
WT_D void bd_walker_to_state(const DScene& sc, const BdWalker& w, uint32_t which, BWalk& d) {
    d.beam = w.beam; d.fwd = which == 1u; d.pdf_from_prev.v = w.pdf_v; d.pdf_from_prev.disc = w.pdf_disc != 0u; d.throughput = w.throughput; d.rr = w.rr;
    d.base = which * sc.cap.verts; d.n = w.n; d.prev_geo = w.prev_geo; d.ap0 = which * sc.cap.ap_walk; d.n_ap = w.n_ap;
}
WT_D void bd_state_to_walker(const BWalk& d, uint32_t rng_d, BdWalker& w) {
    w.beam = d.beam; w.prev_geo = d.prev_geo; w.pdf_v = d.pdf_from_prev.v; w.pdf_disc = d.pdf_from_prev.disc ? 1u : 0u; w.throughput = d.throughput; w.rr = d.rr;
    w.n = d.n; w.rng_d = rng_d; w.n_ap = d.n_ap; w.pad0 = 0u;
}
WT_D void bd_header_to_init(const BdHeader& h, BdSampleInit& si) {
    si.pixel = h.pixel; si.sample = h.sample; si.ex = h.ex; si.ey = h.ey; si.k = h.k; si.rspd = h.rspd; si.wpd_v = h.wpd_v; si.el.ex = h.elx; si.el.ey = h.ely; si.el.ox = h.ox; si.el.oy = h.oy;
}
// last act of a sample: the block splat of the summed t>1 strategies (plt_bdpt.cpp:143-146); frees the slot
WT_D uint32_t bd_finalize(const BdArgs& a, uint32_t slot, float L0) {
    BdHeader h; soa_load(h, a.headers, a.P, slot);
    Element el; el.ex = h.elx; el.ey = h.ely; el.ox = h.ox; el.oy = h.oy;
    const uint32_t n = film_splat(a.r.sc, a.r.film_block, a.r.film_light, false, el, L0, h.k);
    a.r.alive[slot] = 0u;
    atomicSub(&a.r.ctr->live, 1);
    return n;
}

__global__ void __launch_bounds__(128) k_bd_generate(const BdArgs a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    bool gen = false;
    if (slot < a.P && a.r.alive[slot] == 0u) {
        const unsigned long long id = atomicAdd(&a.r.ctr->next_sample, 1ull) * a.r.n_parts + a.r.part;
        if (id < a.r.total) {
            gen = true;
            Counters ctr; counters_zero(ctr);
            BCtx c = mk_bctx(a.r.sc, a.arena, slot, a.lut, &ctr);
            uint32_t ex, ey, sample; bd_sample_coords(id, a.r.tile_x0, a.r.tile_y0, a.r.tile_w, a.r.tile_h, a.r.sample_begin, ex, ey, sample);
            BdSampleInit si; BWalk w0, w1;
            bd_init_sample(c, a.r.seed_lo, a.r.seed_hi, ex, ey, sample, si, w0, w1);
            BdHeader h; h.pixel = si.pixel; h.sample = si.sample; h.ex = ex; h.ey = ey; h.k = si.k; h.rspd = si.rspd; h.wpd_v = si.wpd_v; h.ox = si.el.ox; h.oy = si.el.oy; h.elx = si.el.ex; h.ely = si.el.ey; h.pad0 = 0u;
            soa_store(h, a.headers, a.P, slot);
            BdWalker w; bd_state_to_walker(w0, 0u, w); soa_store(w, a.walkers, 2u * a.P, 2u * slot);
            bd_state_to_walker(w1, 0u, w); soa_store(w, a.walkers, 2u * a.P, 2u * slot + 1u);
            a.r.alive[slot] = 1u; a.pending[slot] = 2; a.L0[slot] = 0.f;
        }
    }
    list_append(a.r.trav_list, &a.r.ctr->n_trav, gen, 2u * slot);
    list_append(a.r.trav_list, &a.r.ctr->n_trav, gen, 2u * slot + 1u);
    const unsigned m = __activemask();
    const unsigned n = __reduce_add_sync(m, gen ? 1u : 0u);
    if (n && (threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) { atomicAdd(&a.r.ctr->live, (int)n); atomicAdd(&a.r.ctr->samples, (unsigned long long)n); }
}

__global__ void __launch_bounds__(128) k_bd_traverse(const BdArgs a) {
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    Counters ctr; counters_zero(ctr);
    bool ovf = false;
    if (li < (uint32_t)a.r.ctr->n_trav) {
        const uint32_t wid = a.r.trav_list[li];
        const DScene& sc = a.r.sc;
        BdWalker w; soa_load(w, a.walkers, 2u * a.P, wid);
        TriWriter tw = tri_writer(sc, a.trav_tris, wid);
        TravOut tr; BHit bh; HitRec h;
        traverse(sc, w.beam.env, w.prev_geo, wavenum_to_wavelen(w.beam.k), sc.sensor.ray_trace_only != 0u, tw, tr, ctr);
        bd_resolve_hit(sc, w.beam, tr, tri_list(sc, a.trav_tris, wid, tr.cone.n_tris, tr.cone.overflow), a.r.hit_edges + (size_t)wid * sc.cap.edges, bh);
        if (bh.need_edges) need_max(&a.r.ctr->need_edges, bh.need_edges);
        h.flags = (bh.empty ? H_EMPTY : 0u) | (bh.ballistic ? H_BALLISTIC : 0u) | (bh.overflow ? H_OVERFLOW : 0u) | (bh.primary != WTGPU_INVALID_IDX ? H_PRIMARY : 0u);
        h.primary = bh.primary; h.pdist = bh.pdist; h.bx = bh.bx; h.by = bh.by; h.d2i = bh.dist; h.region_depth = bh.region_depth; h.origin = bh.origin; h.n_edges = bh.n_edges; h.flux = bh.flux;
        ovf = bh.overflow;
        hit_store(h, a.r.hit, a.r.pool, wid);
        a.r.keys[wid] = bd_hit_key(sc, bh, a.r.n_keys);
    }
    flush_counters(a.r.ctr, ctr);
    count1(&a.r.ctr->overflow, ovf);
    count1(&a.r.ctr->walker_steps, li < (uint32_t)a.r.ctr->n_trav);
}

// traverse() for every walker of the list, eight lanes per walker (gtrav.cuh)
#ifndef WT_GT_MINB
#define WT_GT_MINB 5
#endif
__global__ void __launch_bounds__(128, WT_GT_MINB) k_bd_gtraverse(const BdArgs a) {
    __shared__ GShared shm[128 / kGW];
    Counters ctr; counters_zero(ctr);
    const DScene& sc = a.r.sc;
    g_traverse_all(sc, a.r.ctr->n_trav, &a.r.ctr->trav_head, shm, sc.sensor.ray_trace_only != 0u, false, ctr, a.r.big_save, &a.r.ctr->n_big, a.r.tiers.big_tested,
        [&](int i, Cone& env, Geo& prev, float& lambda, TriWriter& tw) {
            const uint32_t wid = a.r.trav_list[i];
            BdWalker w; soa_load(w, a.walkers, 2u * a.P, wid);
            env = w.beam.env; prev = w.prev_geo; lambda = wavenum_to_wavelen(w.beam.k);
            tw = tri_writer(sc, a.trav_tris, wid);
        },
        [&](int i, const TravRec& r, const GLane& g) { if (g.gl == 0u) a.trav_rec[a.r.trav_list[i]] = r; });
    flush_counters(a.r.ctr, ctr);
}
// the walkers k_bd_gtraverse handed over in the middle of a large cone query: one warp team per walker (ctrav.cuh), then one block team for the largest
__global__ void __launch_bounds__(128, WT_WT_MINB) k_bd_wtraverse(const BdArgs a) {
    __shared__ TShared<32> shm[4];
    Counters ctr; counters_zero(ctr);
    t_traverse_all<32>(a.r.sc, a.r.ctr->n_big, a.r.big_save, &a.r.ctr->big_head, shm[threadIdx.x >> 5], ctr, a.r.huge_save, &a.r.ctr->n_huge, a.r.tiers, a.r.ctr->dbg,
        [&](int i, const TravRec& r, const GLane& g) { if (g.gl == 0u) a.trav_rec[a.r.trav_list[i]] = r; });
    flush_counters(a.r.ctr, ctr);
}
__global__ void __launch_bounds__(256, WT_CT_MINB) k_bd_ctraverse(const BdArgs a) {
    __shared__ TShared<256> shm;
    Counters ctr; counters_zero(ctr);
    t_traverse_all<256>(a.r.sc, a.r.ctr->n_huge, a.r.huge_save, &a.r.ctr->huge_head, shm, ctr, nullptr, nullptr, a.r.tiers, a.r.ctr->dbg + 8,
        [&](int i, const TravRec& r, const GLane& g) { if (g.gl == 0u) a.trav_rec[a.r.trav_list[i]] = r; });
    flush_counters(a.r.ctr, ctr);
}
// what the vertex step needs from a traversal result: primary triangle, Gaussian power over clipped triangles, edges, sort key
WT_D void bd_trav_out(const TravRec& r, TravOut& tr) {
    tr.empty = (r.flags & TR_EMPTY) != 0u; tr.ballistic = (r.flags & TR_BALLISTIC) != 0u;
    tr.ray.tuid = r.ray_tuid; tr.ray.dist = r.ray_dist; tr.ray.bx = r.bx; tr.ray.by = r.by; tr.ray.front = (r.flags & TR_RAY_FRONT) != 0u;
    tr.cone.dist = r.cone_dist; tr.cone.front = (r.flags & TR_CONE_FRONT) != 0u; tr.cone.n_tris = r.n_tris; tr.cone.overflow = (r.flags & TR_OVERFLOW) != 0u;
    tr.region_depth = r.region_depth; tr.origin = mk3(r.ox, r.oy, r.oz);
}
WT_D void bd_store_hit(const BdArgs& a, uint32_t wid, const BHit& bh) {
    HitRec h;
    h.flags = (bh.empty ? H_EMPTY : 0u) | (bh.ballistic ? H_BALLISTIC : 0u) | (bh.overflow ? H_OVERFLOW : 0u) | (bh.primary != WTGPU_INVALID_IDX ? H_PRIMARY : 0u);
    h.primary = bh.primary; h.pdist = bh.pdist; h.bx = bh.bx; h.by = bh.by; h.d2i = bh.dist; h.region_depth = bh.region_depth; h.origin = bh.origin; h.n_edges = bh.n_edges; h.flux = bh.flux;
    hit_store(h, a.r.hit, a.r.pool, wid);
    a.r.keys[wid] = bd_hit_key(a.r.sc, bh, a.r.n_keys);
    if (bh.need_edges) need_max(&a.r.ctr->need_edges, bh.need_edges);
}
// ---- find_closest_triangle (plt_bdpt_detail.hpp:362-390) of a long list as flat tasks: a warp scans kClosestChunk consecutive entries and folds its
// (distance, list index) minimum -- the first entry, in list order, with the smallest hit distance -- into the walker's 64-bit cell with one atomicMin.
// Key: the distance's bits mapped so that unsigned order = float order, then the index.
constexpr uint32_t kClosestChunk = 256u;
WT_D unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
    return v;
}
WT_D unsigned long long closest_key(float d, uint32_t idx) { d = d + 0.f;      // (-0 -> +0: they compare equal in the sequential loop)
     const uint32_t b = __float_as_uint(d); return ((unsigned long long)(b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u)) << 32) | idx; }
__global__ void __launch_bounds__(128) k_bd_closest_chunks(const BdArgs a) {
    const unsigned lane = threadIdx.x & 31u, FULL = 0xffffffffu;
    const DScene& sc = a.r.sc;
    for (;;) {
        int i = 0;
        if (lane == 0u) i = atomicAdd(&a.r.ctr->closest_task_head, 1);
        i = __shfl_sync(FULL, i, 0);
        if (i >= a.r.ctr->n_closest_tasks) break;
        const uint2 task = a.r.closest_tasks[i];
        const uint32_t wid = a.r.trav_list[task.x];
        const TravRec r = a.trav_rec[wid];
        const TriList tl = tri_list(sc, a.trav_tris, wid, r.n_tris, (r.flags & TR_OVERFLOW) != 0u);
        const V3 origin = mk3(r.ox, r.oy, r.oz);
        // the beam's mean direction: floats 3..5 of BdWalker (beam.env = {o, d, ...}), i.e. chunk 0 .w and chunk 1 .xy
        static_assert(offsetof(BdWalker, beam) == 0 && offsetof(Beam, env) == 0 && offsetof(Cone, d) == 12, "BdWalker layout");
        const float4 c0 = a.walkers[wid], c1 = a.walkers[(size_t)2u * a.P + wid];
        const V3 dir = mk3(c0.w, c1.x, c1.y);
        const Range zr = mkr(r.cone_dist, r.cone_dist + r.region_depth);
        unsigned long long best = ~0ull;
        const uint32_t e0 = task.y * kClosestChunk, e1 = min(tl.n, e0 + kClosestChunk);
        for (uint32_t e = e0 + lane; e < e1; e += 32u) {
            const Tri3 t = load_tri(sc, tri_at(tl, e));
            const float tol = cone_intersection_tolerance(origin, t.a, t.b, t.c);
            const RayTri rt = intersect_ray_tri(origin, dir, t.a, t.b, t.c, mkr(zr.mn - tol, zr.mx + tol));
            if (rt.hit && rt.dist < WT_INF) best = min(best, closest_key(rt.dist, e));
        }
        best = warp_min_u64(best);
        if (lane == 0u && best != ~0ull) atomicMin(a.r.closest_best + wid, best);
    }
}
__global__ void __launch_bounds__(128) k_bd_resolve(const BdArgs a) {
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    bool act = li < (uint32_t)a.r.ctr->n_trav;
    const DScene& sc = a.r.sc;
    bool ovf = false;
    uint32_t wid = 0u;
    BdWalker w; TravOut tr; BHit bh;
    TriList tl; tl.row = nullptr; tl.spill = nullptr; tl.ext = nullptr; tl.n = 0u;
    uint32_t* edges = nullptr;
    tr.empty = true; tr.ballistic = true; tr.cone.n_tris = 0u;
    if (act) {
        wid = a.r.trav_list[li];
        const TravRec r = a.trav_rec[wid];
        bd_trav_out(r, tr);
        tl = tri_list(sc, a.trav_tris, wid, r.n_tris, tr.cone.overflow);
        if (tl.n > kBigQuery && !tr.empty && !tr.ballistic) {      // a long list: its closest-triangle search as flat 256-entry tasks, the rest by a warp (k_bd_resolve_big)
            a.r.big_res_list[atomicAdd(&a.r.ctr->n_big_res, 1)] = li; act = false;
            a.r.closest_best[wid] = ~0ull;
            const uint32_t nt = (tl.n + kClosestChunk - 1u) / kClosestChunk;
            const uint32_t t0 = (uint32_t)atomicAdd(&a.r.ctr->n_closest_tasks, (int)nt);
            for (uint32_t c = 0u; c < nt; ++c) a.r.closest_tasks[t0 + c] = make_uint2(li, c);
        }
        else { soa_load(w, a.walkers, 2u * a.P, wid); edges = a.r.hit_edges + (size_t)wid * sc.cap.edges; }
    }
    bd_resolve_hit_warp(sc, act, w.beam, tr, tl, edges, bh);
    if (act) { ovf = bh.overflow; bd_store_hit(a, wid, bh); }
    count1(&a.r.ctr->overflow, ovf);
    count1(&a.r.ctr->walker_steps, act);
}
__global__ void __launch_bounds__(128) k_bd_resolve_big(const BdArgs a, uint32_t bit_words) {
    const unsigned lane = threadIdx.x & 31u, FULL = 0xffffffffu;
    const DScene& sc = a.r.sc;
    uint32_t* bits = a.r.edge_bits + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * bit_words;
    unsigned long long n_step = 0, n_ovf = 0;
    for (;;) {
        int i = 0;
        if (lane == 0u) i = atomicAdd(&a.r.ctr->big_res_head, 1);
        i = __shfl_sync(FULL, i, 0);
        if (i >= a.r.ctr->n_big_res) break;
        const uint32_t li = a.r.big_res_list[i], wid = a.r.trav_list[li];
        const TravRec r = a.trav_rec[wid];
        TravOut tr; bd_trav_out(r, tr);
        BdWalker w; soa_load(w, a.walkers, 2u * a.P, wid);
        const TriList tl = tri_list(sc, a.trav_tris, wid, r.n_tris, tr.cone.overflow);
        BHit bh;
        bool deferred = false; uint32_t sbase = 0u;
        if (bd_resolve_closest_big(sc, w.beam, tr, tl, __ldcg(a.r.closest_best + wid), bh)) {
            // the Gaussian power: the list is queued for the flat kernels -- scratch for the piece values (16 B per entry) is bump-allocated;
            // none left -> computed here
            if (a.r.flux_cap) {
                if (lane == 0u) sbase = atomicAdd(&a.r.ctr->flux_scratch_head, tl.n);
                sbase = __shfl_sync(FULL, sbase, 0);
                deferred = sbase <= a.r.flux_cap && tl.n <= a.r.flux_cap - sbase;
            }
            if (!deferred) bd_resolve_flux_big(sc, w.beam, tr, tl, a.r.hit_edges + (size_t)wid * sc.cap.edges, bits, bh);
        }
        if (deferred) {       // queue the list: one item, ceil(n / 32) chunk tasks
            const uint32_t nt = (tl.n + 31u) / 32u;
            uint32_t it = 0u, t0 = 0u;
            if (lane == 0u) { it = (uint32_t)atomicAdd(&a.r.ctr->n_flux_items, 1); t0 = (uint32_t)atomicAdd(&a.r.ctr->n_flux_tasks, (int)nt); a.r.flux_items[it] = make_uint2(li, sbase); }
            it = __shfl_sync(FULL, it, 0); t0 = __shfl_sync(FULL, t0, 0);
            for (uint32_t c = lane; c < nt; c += 32u) a.r.flux_tasks[t0 + c] = make_uint2(it, c);
        } else if (lane == 0u) { bd_store_hit(a, wid, bh); ++n_step; n_ovf += bh.overflow ? 1u : 0u; }
        __syncwarp();
    }
    if (lane == 0u) { if (n_step) atomicAdd(&a.r.ctr->walker_steps, n_step); if (n_ovf) atomicAdd(&a.r.ctr->overflow, n_ovf); }
}
// what a flux task / item needs of its walker
struct FluxCtx { Beam beam; TravOut tr; TriList tl; Range zr; Frame beam_frame; G2 wf; float csz; BHit h; uint32_t wid; };
WT_D void bd_flux_ctx(const BdArgs& a, uint32_t li, FluxCtx& c) {
    const DScene& sc = a.r.sc;
    c.wid = a.r.trav_list[li];
    const TravRec r = a.trav_rec[c.wid];
    bd_trav_out(r, c.tr);
    BdWalker w; soa_load(w, a.walkers, 2u * a.P, c.wid);
    c.beam = w.beam;
    c.tl = tri_list(sc, a.trav_tris, c.wid, r.n_tris, c.tr.cone.overflow);
    bd_hit_init(c.h, c.beam, c.tr, c.zr);
    c.beam_frame = cone_frame(c.beam.env);
    c.wf = wavefront_of(c.beam, c.h.dist);
    c.csz = (c.zr.mx + c.zr.mn) / 2.f;
}
// flat tasks: the piece values of 32 consecutive entries of a queued list -> scratch
__global__ void __launch_bounds__(128) k_bd_flux_chunks(const BdArgs a) {
    const unsigned lane = threadIdx.x & 31u, FULL = 0xffffffffu;
    const DScene& sc = a.r.sc;
    for (;;) {
        int i = 0;
        if (lane == 0u) i = atomicAdd(&a.r.ctr->flux_task_head, 1);
        i = __shfl_sync(FULL, i, 0);
        if (i >= a.r.ctr->n_flux_tasks) break;
        const uint2 task = a.r.flux_tasks[i], item = a.r.flux_items[task.x];
        FluxCtx c; bd_flux_ctx(a, item.x, c);
        const uint32_t base = task.y * 32u;
        float v0, v1, v2; int cnt;
        const QuadQ qq = { a.r.quad_tasks, &a.r.ctr->n_quad_tasks, (int)a.r.quad_cap };
        bd_flux_chunk(sc, c.beam, c.tr.cone.front, c.tl, c.zr, c.beam_frame, c.wf, c.csz, base, v0, v1, v2, cnt, &qq, ((size_t)item.y + base + lane) * 4u);
        if (base + lane < c.tl.n) a.r.flux_scratch[(size_t)item.y + base + lane] = make_float4(v0, v1, v2, __int_as_float(cnt));
        __syncwarp();
    }
}
// the long quadrature pieces k_bd_flux_chunks queued: one warp per piece, the value into its slot of the scratch
__global__ void __launch_bounds__(128) k_bd_quad_tasks(const BdArgs a) {
    const unsigned lane = threadIdx.x & 31u, FULL = 0xffffffffu;
    const int n = min(a.r.ctr->n_quad_tasks, (int)a.r.quad_cap);
    for (;;) {
        int i = 0;
        if (lane == 0u) i = atomicAdd(&a.r.ctr->quad_task_head, 1);
        i = __shfl_sync(FULL, i, 0);
        if (i >= n) break;
        const float4 t0 = a.r.quad_tasks[2 * (size_t)i], t1 = a.r.quad_tasks[2 * (size_t)i + 1];
        const float v = g2_quadrature_warp(mk2(t0.x, t0.y), mk2(t0.z, t0.w), mk2(t1.x, t1.y));
        const size_t dst = (size_t)__float_as_uint(t1.z) | ((size_t)__float_as_uint(t1.w) << 32);
        if (lane == 0u) reinterpret_cast<float*>(a.r.flux_scratch)[dst] = v;
        __syncwarp();
    }
}
// one warp per queued list: its piece values added in list order, the edge set, the hit record
__global__ void __launch_bounds__(128) k_bd_flux_finish(const BdArgs a, uint32_t bit_words) {
    const unsigned lane = threadIdx.x & 31u, FULL = 0xffffffffu;
    const DScene& sc = a.r.sc;
    uint32_t* bits = a.r.edge_bits + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * bit_words;
    unsigned long long n_step = 0, n_ovf = 0;
    for (;;) {
        int i = 0;
        if (lane == 0u) i = atomicAdd(&a.r.ctr->flux_item_head, 1);
        i = __shfl_sync(FULL, i, 0);
        if (i >= a.r.ctr->n_flux_items) break;
        const uint2 item = a.r.flux_items[i];
        FluxCtx c; bd_flux_ctx(a, item.x, c);
        float flux = 0.f;
        for (uint32_t base = 0u; base < c.tl.n; base += 32u) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (base + lane < c.tl.n) v = __ldcg(a.r.flux_scratch + (size_t)item.y + base + lane);
            flux = bd_flux_add(flux, v.x, v.y, v.z, __float_as_int(v.w));
        }
        c.h.flux = flux;
        if (sc.integrator.fsd) { bool eo = false; uint32_t need = 0u; c.h.n_edges = w_collect_edges(sc, c.tl, a.r.hit_edges + (size_t)c.wid * sc.cap.edges, sc.cap.edges, bits, eo, need); if (eo) { c.h.overflow = true; c.h.need_edges = need; } }
        if (lane == 0u) { bd_store_hit(a, c.wid, c.h); ++n_step; n_ovf += c.h.overflow ? 1u : 0u; }
        __syncwarp();
    }
    if (lane == 0u) { if (n_step) atomicAdd(&a.r.ctr->walker_steps, n_step); if (n_ovf) atomicAdd(&a.r.ctr->overflow, n_ovf); }
}

__global__ void k_bd_reset(const BdArgs a) { if (threadIdx.x == 0 && blockIdx.x == 0) { a.r.ctr->n_trav = 0; a.r.ctr->trav_head = 0; for (int c = 0; c < kPairClasses; ++c) a.r.ctr->n_pairs[c] = 0; a.r.ctr->n_fsd_list[a.fl_next] = 0; reset_iteration_lists(a.r.ctr); } }

// Strategy classes = the branches of connect_subpaths (plt_bdpt_detail.hpp:747-923): each class has its own task list and its own
// launch, so a warp runs one branch (emission hit / sensor hit / emitter-direct / sensor-direct / vertex-vertex).
WT_D int bd_pair_class(int s, int t) { return s == 0 ? 0 : t == 0 ? 1 : s == 1 ? 2 : t == 1 ? 3 : 4; }
// a walk has ended with n vertices: when it is the sample's second, expand the sample's strategies into the task lists
WT_D void bd_walker_done(const BdArgs& a, uint32_t wid, uint32_t n, uint32_t& n_splat) {
    const uint32_t slot = wid >> 1, which = wid & 1u;
    a.nverts[wid] = n;
    __threadfence();
    if (atomicSub(&a.pending[slot], 1) == 1) {
        __threadfence();
        const uint32_t other = __ldcg(a.nverts + (wid ^ 1u));
        const uint32_t nsv = which == 0u ? n : other, nev = which == 1u ? n : other;
        int np = 0;
        bd_for_each_pair(a.r.sc, nsv, nev, [&](int, int) { ++np; });
        if (np == 0) n_splat += bd_finalize(a, slot, 0.f);
        else {
            a.pending[slot] = np;
            int cnt[kPairClasses] = { 0, 0, 0, 0, 0 };
            bd_for_each_pair(a.r.sc, nsv, nev, [&](int s, int t) { ++cnt[bd_pair_class(s, t)]; });
            uint32_t at[kPairClasses];
#pragma unroll
            for (int c = 0; c < kPairClasses; ++c) at[c] = cnt[c] ? (uint32_t)atomicAdd(&a.r.ctr->n_pairs[c], cnt[c]) : 0u;
            bd_for_each_pair(a.r.sc, nsv, nev, [&](int s, int t) { const int c = bd_pair_class(s, t); a.pairs[a.pair_off[c] + at[c]++] = (unsigned long long)slot | ((unsigned long long)s << 32) | ((unsigned long long)t << 48); });
        }
    }
}
WT_D void bd_hit_to_bhit(const HitRec& h, BHit& bh) {
    bh.empty = (h.flags & H_EMPTY) != 0u; bh.ballistic = (h.flags & H_BALLISTIC) != 0u; bh.overflow = (h.flags & H_OVERFLOW) != 0u;
    bh.primary = h.primary; bh.pdist = h.pdist; bh.bx = h.bx; bh.by = h.by; bh.dist = h.d2i; bh.region_depth = h.region_depth; bh.flux = h.flux; bh.origin = h.origin; bh.n_edges = h.n_edges; bh.need_edges = 0u;
}

__global__ void __launch_bounds__(128) k_bd_shade(const BdArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    Counters ctr; counters_zero(ctr);
    uint32_t n_vert = 0, n_splat = 0, wid = 0; int res = BD_END; bool overflow = false, act = false;
    if (i < (uint32_t)a.r.ctr->n_sorted) {
        const DScene& sc = a.r.sc;
        act = true;
        wid = a.r.order[i];
        const uint32_t slot = wid >> 1, which = wid & 1u;
        BCtx c = mk_bctx(sc, a.arena, slot, a.lut, &ctr);
        BdWalker w; soa_load(w, a.walkers, 2u * a.P, wid);
        HitRec h; hit_load(h, a.r.hit, a.r.pool, wid);
        BdHeader hd; soa_load(hd, a.headers, a.P, slot);
        BWalk d; bd_walker_to_state(sc, w, which, d);
        BHit bh; bd_hit_to_bhit(h, bh);
        Sampler smp; smp.k0 = a.r.seed_lo; smp.k1 = a.r.seed_hi; smp.pixel = hd.pixel; smp.sample = hd.sample; smp.d = w.rng_d; smp.stream = 1u + which;
        res = bd_walk_step(c, d, smp, bh, a.r.hit_edges + (size_t)wid * sc.cap.edges, n_vert, true);
        overflow = c.overflow; bd_flush_need(a.r.ctr, c);
        if (res != BD_END) { bd_state_to_walker(d, smp.d, w); soa_store(w, a.walkers, 2u * a.P, wid); }
        else bd_walker_done(a, wid, d.n, n_splat);
    }
    list_append(a.r.trav_list, &a.r.ctr->n_trav, act && res == BD_CONTINUE, wid);
    if (act && res == BD_FSD_DEFERRED) a.fsd_out[2u * wid + 1u] = make_float4(0.f, 0.f, 0.f, 0.f);      // .w: 0 fresh, 1 carried over, >= 16 sampled (iteration tag)
    list_append(a.fsd_list + (size_t)a.fl_cur * 2u * a.P, &a.r.ctr->n_fsd_list[a.fl_cur], act && res == BD_FSD_DEFERRED, wid);
    flush_counters(a.r.ctr, ctr, true);
    bd_flush_stats(a.r.ctr, n_splat, n_vert, 0u, 0u, overflow);
}

// fsd_sampler_t::sample (src/interaction/fsd/fraunhofer/fsd_sampler.cpp:81-113) + free_space_diffraction_t::sample (free_space_diffraction.hpp:68-93)
// for every walker whose aperture was just built.  Rejection sampling takes a geometric number of tries (mean = the segment count, ~21 on
// double_slits; a few apertures: thousands), each try a sum over the aperture's <= 48 segments.
//
// ONE WARP samples one walker, EIGHT TRIES AT A TIME (speculatively: tries are independent given where their draws start):
//   1. the draw index of try t+1 depends on try t only through "did try t pick the P0 lobe" (3 draws) "or a segment" (5 draws), i.e. on one
//      comparison of try t's first draw: a window of Philox blocks is computed lane-parallel and the eight start indices follow from a
//      short scan over it;
//   2. lane t generates the candidate xi of try t (selection by binary search of the cdf in shared memory, LUT lookup) from its own
//      position of the stream -- the expensive warp-uniform part of a try is done for eight tries at the cost of one;
//   3. the 8 x n (try, segment) terms Psi, |Psi|^2 are spread over the 32 lanes and written to shared memory;
//   4. lane t adds the n terms of try t IN SEGMENT ORDER (sums, and therefore accept/reject decisions, are bit-identical to the sequential
//      fraunhofer_sample), evaluates the acceptance test with try t's last draw; the first accepting try in sequence order wins.
// Tries evaluated past the accepted one are discarded (~3.5 of ~21 on average).  Runs on a second stream, overlapped with the next
// iteration's kernels; its results are picked up one iteration later.  A launch does not wait for stragglers: after `budget` tries a
// walker is carried over to the next iteration (the stream is counter-based: only the draw index and the try count are kept).
constexpr int kFsdK = 8;                    // speculative tries per batch
constexpr int kFsdRow = kFsdChunk + 1;      // padded row of the term arrays (bank-conflict free for the per-try sums)
// Shared-memory staging holds kFsdChunk segments.  An aperture with more (rare: a beam footprint over finely tessellated geometry) is walked
// chunk by chunk -- the per-try sums still run over the segments in order -- and its selection cdf is scanned from the aperture record in HBM.
struct FsdShared { float4 ea[kFsdChunk], eb[kFsdChunk]; float cdf[kFsdChunk + 1]; float re[kFsdK][kFsdRow], im[kFsdK][kFsdRow], dd[kFsdK][kFsdRow]; };
__global__ void __launch_bounds__(128, 6) k_bd_fsd_sample(const BdArgs a) {
    __shared__ FsdShared shm[4];
    FsdShared& sh = shm[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const int n_tasks = a.r.ctr->n_fsd_list[a.fl_cur];
    const uint32_t* list = a.fsd_list + (size_t)a.fl_cur * 2u * a.P;
    uint32_t* carry_list = a.fsd_list + (size_t)a.fl_next * 2u * a.P;
    int budget = n_tasks > 2048 ? 128 : 0x7fffffff;       // tries per WARP per launch; unbounded once only stragglers remain
    for (;;) {
        int t = 0;
        if (lane == 0u) t = atomicAdd(&a.r.ctr->fsd_head, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tasks) break;
        const uint32_t wid = list[t];
        if (budget <= 0) {      // out of budget: hand the rest of the list to the next iteration untouched
            if (lane == 0u) carry_list[atomicAdd(&a.r.ctr->n_fsd_list[a.fl_next], 1)] = wid;
            continue;
        }
        const uint32_t slot = wid >> 1, which = wid & 1u;
        const Arena A = mk_arena(a.r.sc, a.arena, slot);
        uint32_t w_rng_d, w_n_ap;
        { BdWalker w; soa_load(w, a.walkers, 2u * a.P, wid); w_rng_d = w.rng_d; w_n_ap = w.n_ap; }
        BdHeader h; soa_load(h, a.headers, a.P, slot);
        const int ai = (int)(which * a.r.sc.cap.ap_walk + w_n_ap - 1u);
        const FHead hd = ap_head(A, ai);
        const uint32_t n = hd.n;
        const float4 st = a.fsd_out[2u * wid + 1u];      // carried over: resume after the tries already spent
        const bool carried = st.w == 1.f;
        uint32_t d_base = carried ? __float_as_uint(st.y) : w_rng_d;
        uint32_t tries = carried ? __float_as_uint(st.z) : 0u;
        const bool rej = n > 1u; const uint32_t max_tries = n * 1024u; const float recp_M = 1.f / (float)n;
        // stage the aperture: segments, and the selection cdf of sampleN (fsd_sampler.cpp:37-79) summed in the sequential order
        // (entry 0 is the P0 lobe, entry i the segment i-1; non-decreasing, so "first i with p < cdf[i]" = #{i : !(p < cdf[i])})
        const bool small = n <= (uint32_t)kFsdChunk;
        auto stage = [&](uint32_t j0) {         // segments [j0, j0 + kFsdChunk) into shared memory
            __syncwarp();
            for (uint32_t j = j0 + lane; j < min(n, j0 + (uint32_t)kFsdChunk); j += 32u) {
                const FEdge e = ap_edge(A, ai, j);
                sh.ea[j - j0] = make_float4(e.e.x, e.e.y, e.v.x, e.v.y); sh.eb[j - j0] = make_float4(e.a_b.re, e.a_b.im, e.iab_2.re, e.iab_2.im);
            }
            __syncwarp();
        };
        stage(0u);
        if (small && lane == 0u) { float cdf = 0.f; for (uint32_t i = 0; i < n; ++i) { cdf += i == 0u ? hd.P0_pdf : ap_edge_pdf(A, ai, i - 1u); sh.cdf[i] = cdf; } }
        __syncwarp();
        const float cdf0 = hd.P0_pdf;       // == sh.cdf[0] (0 + P0_pdf)
        V3 wo = mk3(0.f, 0.f, 1.f); float dpd = 0.f, wgt = 0.f;
        bool finished = n == 0u;            // an empty aperture makes no try (max_tries == 0)
        uint32_t d_final = d_base;
        while (!finished && budget > 0) {
            const uint32_t kv = min((uint32_t)kFsdK, max_tries - tries);      // tries of this batch
            // 1. start index of every try: window of 13 Philox blocks (8 tries x <= 6 draws), then the scan
            const uint32_t blk0 = d_base >> 2;
            uint32_t c0, c1, c2, c3;
            {
                uint32_t x0 = blk0 + lane, x1 = h.sample, x2 = h.pixel, x3 = 1u + which, k0 = a.r.seed_lo, k1 = a.r.seed_hi;
#pragma unroll
                for (int r = 0; r < 10; ++r) {
                    const uint32_t h0 = __umulhi(0xD2511F53u, x0), l0 = 0xD2511F53u * x0;
                    const uint32_t h1 = __umulhi(0xCD9E8D57u, x2), l1 = 0xCD9E8D57u * x2;
                    const uint32_t n0 = h1 ^ x1 ^ k0, n2 = h0 ^ x3 ^ k1;
                    x0 = n0; x1 = l1; x2 = n2; x3 = l0;
                    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
                }
                c0 = x0; c1 = x1; c2 = x2; c3 = x3;
            }
            uint32_t my_d = d_base, dt = d_base;
#pragma unroll
            for (int q = 0; q < kFsdK; ++q) {
                if (lane == (unsigned)q) my_d = dt;
                const uint32_t word = dt & 3u;
                const uint32_t mine = word == 0u ? c0 : word == 1u ? c1 : word == 2u ? c2 : c3;
                const uint32_t u = __shfl_sync(0xffffffffu, mine, (int)((dt >> 2) - blk0));
                const float p = (float)(u >> 8) * (1.0f / 16777216.0f) * 1.f;
                dt += (p < cdf0 && n > 0u ? 3u : 5u) + (rej ? 1u : 0u);
            }
            // 2. lane t: candidate of try t
            Sampler sm0; sm0.k0 = a.r.seed_lo; sm0.k1 = a.r.seed_hi; sm0.pixel = h.pixel; sm0.sample = h.sample; sm0.d = my_d; sm0.stream = 1u + which;
            SamplerC smp = samplerc(sm0);
            V2 xi = mk2(0.f, 0.f);
            const bool mine_valid = lane < kv;
            if (mine_valid) {
                const float p = rnd(smp) * 1.f;
                uint32_t sel;                           // sel = #{i < n : !(p < cdf[i])}, cdf summed in sequence (fsd_sampler.cpp:37-79)
                if (small) { uint32_t lo = 0u, hi = n; while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (!(p < sh.cdf[mid])) lo = mid + 1u; else hi = mid; } sel = lo; }
                else { float cdf = 0.f; sel = n; for (uint32_t i = 0; i < n; ++i) { cdf += i == 0u ? hd.P0_pdf : ap_edge_pdf(A, ai, i - 1u); if (p < cdf) { sel = i; break; } } }
                if (sel == 0u) { const float u0 = rnd(smp); const float u1 = rnd(smp); xi = kP0s * normal2d(mk2(u0, u1)); }
                else {
                    float4 ea, eb;
                    if (small) { ea = sh.ea[sel - 1u]; eb = sh.eb[sel - 1u]; }
                    else { const FEdge e = ap_edge(A, ai, sel - 1u); ea = make_float4(e.e.x, e.e.y, e.v.x, e.v.y); eb = make_float4(e.a_b.re, e.a_b.im, e.iab_2.re, e.iab_2.im); }
                    const V2 ee = mk2(ea.x, ea.y);
                    const V2 m = mk2(ee.y, -ee.x);
                    const float od = 1.f / (ee.x * m.y - m.x * ee.y);
                    const float i00 = m.y * od, i01 = -ee.y * od, i10 = -m.x * od, i11 = ee.x * od;
                    const float Aa = cnorm(mkc(eb.x, eb.y)), Bb = cnorm(mkc(eb.z, eb.w));
                    const float pp = rnd(smp) * (Aa + Bb);
                    const float r0 = rnd(smp); const float r1 = rnd(smp); const float r2 = rnd(smp);
                    const V3 r3 = mk3(r0, r1, r2);
                    const V2 z = pp < Aa ? flut_sample(a.lut, r3, a.lut.th1, a.lut.c1) : flut_sample(a.lut, r3, a.lut.th2, a.lut.c2);
                    xi = mk2(z.x * i00 + z.y * i01, z.x * i10 + z.y * i11);
                }
            }
            // 3. + 4.  per chunk of staged segments: the kv x nc (try, segment) terms spread over the warp, then lane t adds try t's terms IN
            // SEGMENT ORDER to its running sums; after the last chunk, the acceptance test
            C2 acc = mkc(0.f, 0.f); float dens = 0.f;
            for (uint32_t j0 = 0u; j0 < n; j0 += (uint32_t)kFsdChunk) {
                const uint32_t nc = min(n - j0, (uint32_t)kFsdChunk);
                if (!small) stage(j0);
                {
                    const uint32_t items = kv * nc;
                    uint32_t tt = 0u, jj = lane;
                    for (uint32_t base = 0u; base < items; base += 32u) {
                        while (jj >= nc && tt < kv) { jj -= nc; ++tt; }
                        const bool act = tt < kv;
                        const float xx = __shfl_sync(0xffffffffu, xi.x, (int)(act ? tt : 0u)), xy = __shfl_sync(0xffffffffu, xi.y, (int)(act ? tt : 0u));
                        if (act) {
                            const float4 ea = sh.ea[jj], eb = sh.eb[jj];
                            FEdge e; e.e = mk2(ea.x, ea.y); e.v = mk2(ea.z, ea.w); e.a_b = mkc(eb.x, eb.y); e.iab_2 = mkc(eb.z, eb.w);
                            const V2 x2 = mk2(xx, xy);
                            const V2 z = fzeta(e, x2);
                            const C2 sx = e.a_b * falpha1(z.x, z.y) + e.iab_2 * falpha2(z.x, z.y);
                            const float rho = length2(e.e);
                            float sn, cs; pm::sincosf(-dot(e.v, x2), &sn, &cs);
                            const C2 tm = mkc(rho * cs, rho * sn) * sx;
                            sh.re[tt][jj] = tm.re; sh.im[tt][jj] = tm.im; sh.dd[tt][jj] = sqrf(rho) * cnorm(sx);
                        }
                        jj += 32u;
                    }
                }
                __syncwarp();
                if (mine_valid) for (uint32_t j = 0; j < nc; ++j) { acc = acc + mkc(sh.re[lane][j], sh.im[lane][j]); dens += sh.dd[lane][j]; }
                __syncwarp();
            }
            bool done = false; float f = 0.f;
            if (mine_valid) {
                const float g = dens * fchi_e(xi) + hd.P0 * kInvTwoPi / sqrf(kP0s) * fchi_0(xi);
                f = cnorm(acc) * fchi_e(xi) + hd.psi02 * fchi_0(xi);
                done = rej ? rnd(smp) * g < f * recp_M : true;
            }
            __syncwarp();
            const unsigned dm = __ballot_sync(0xffffffffu, done);
            if (dm) {
                const int win = __ffs(dm) - 1;
                if ((int)lane == win) {
                    const float pdf = f * hd.recp_I;
                    if (pdf > 0.f) {
                        const V2 zeta = xi / hd.k;
                        const V2 wl = mk2(zeta.x / sqrtf(1.f + sqrf(zeta.x)), zeta.y / sqrtf(1.f + sqrf(zeta.y)));
                        const float wo2 = length2(wl);
                        if (wo2 < .85f) { wo = mk3(wl.x, wl.y, sqrtf(1.f - wo2)); dpd = pdf; wgt = 1.f; }
                    }
                }
                wo.x = __shfl_sync(0xffffffffu, wo.x, win); wo.y = __shfl_sync(0xffffffffu, wo.y, win); wo.z = __shfl_sync(0xffffffffu, wo.z, win);
                dpd = __shfl_sync(0xffffffffu, dpd, win); wgt = __shfl_sync(0xffffffffu, wgt, win);
                d_final = __shfl_sync(0xffffffffu, smp.s.d, win);
                finished = true; budget -= win + 1;
            } else {
                d_final = __shfl_sync(0xffffffffu, smp.s.d, (int)kv - 1);
                d_base = d_final; tries += kv; budget -= (int)kv;
                if (tries == max_tries) finished = true;
            }
        }
        if (lane == 0u) {
            if (finished) {
                a.fsd_out[2u * wid] = make_float4(wo.x, wo.y, wo.z, dpd);
                a.fsd_out[2u * wid + 1u] = make_float4(wgt, __uint_as_float(d_final), 0.f, a.tag);
            } else {
                a.fsd_out[2u * wid + 1u] = make_float4(0.f, __uint_as_float(d_final), __uint_as_float(tries), 1.f);
                carry_list[atomicAdd(&a.r.ctr->n_fsd_list[a.fl_next], 1)] = wid;
            }
        }
    }
}

__global__ void __launch_bounds__(128) k_bd_fsd_finish(const BdArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    Counters ctr; counters_zero(ctr);
    uint32_t n_vert = 0, n_splat = 0, wid = 0; bool survive = false, overflow = false;
    if (i < (uint32_t)a.r.ctr->n_fsd_list[a.fl_fin] && a.fsd_out[2u * a.fsd_list[(size_t)a.fl_fin * 2u * a.P + i] + 1u].w == a.tag_fin) {
        const DScene& sc = a.r.sc;
        wid = a.fsd_list[(size_t)a.fl_fin * 2u * a.P + i];
        const uint32_t slot = wid >> 1, which = wid & 1u;
        BCtx c = mk_bctx(sc, a.arena, slot, a.lut, &ctr);
        BdWalker w; soa_load(w, a.walkers, 2u * a.P, wid);
        HitRec h; hit_load(h, a.r.hit, a.r.pool, wid);
        BdHeader hd; soa_load(hd, a.headers, a.P, slot);
        BWalk d; bd_walker_to_state(sc, w, which, d);
        const float4 o0 = a.fsd_out[2u * wid], o1 = a.fsd_out[2u * wid + 1u];
        Sampler smp; smp.k0 = a.r.seed_lo; smp.k1 = a.r.seed_hi; smp.pixel = hd.pixel; smp.sample = hd.sample; smp.d = __float_as_uint(o1.y); smp.stream = 1u + which;
        const V3 interaction_wp = h.origin + h.d2i * d.beam.env.d;
        survive = bd_walk_fsd_finish(c, d, smp, (int)(d.ap0 + d.n_ap - 1u), interaction_wp, h.d2i, mk3(o0.x, o0.y, o0.z), o0.w, o1.x, n_vert);
        overflow = c.overflow; bd_flush_need(a.r.ctr, c);
        if (survive) { bd_state_to_walker(d, smp.d, w); soa_store(w, a.walkers, 2u * a.P, wid); }
        else bd_walker_done(a, wid, d.n, n_splat);
    }
    list_append(a.r.trav_list, &a.r.ctr->n_trav, survive, wid);
    flush_counters(a.r.ctr, ctr, true);
    bd_flush_stats(a.r.ctr, n_splat, n_vert, 0u, 0u, overflow);
}

template <int CLS> __global__ void __launch_bounds__(128) k_bd_connect(const BdArgs a) {
    Counters ctr; counters_zero(ctr);
    uint32_t n_conn = 0, n_splat = 0; bool overflow = false;
    const uint32_t np = (uint32_t)a.r.ctr->n_pairs[CLS];
    const unsigned long long* pairs = a.pairs + a.pair_off[CLS];
    const DScene& sc = a.r.sc;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) {
        const unsigned long long task = pairs[i];
        const uint32_t slot = (uint32_t)task; const int s = (int)((task >> 32) & 0xffffull), t = (int)(task >> 48);
        BCtx c = mk_bctx(sc, a.arena, slot, a.lut, &ctr);
        BdHeader hd; soa_load(hd, a.headers, a.P, slot);
        BdSampleInit si; bd_header_to_init(hd, si);
        const uint32_t nsv = a.nverts[2u * slot], nev = a.nverts[2u * slot + 1u];
        BConn cr;
        const float flux = bd_eval_pair<CLS>(c, si, a.r.seed_lo, a.r.seed_hi, nsv, nev, s, t, cr);
        ++n_conn; overflow |= c.overflow;
        if (cr.L.s[0] > 0.f) {
            if (t > 1) atomicAdd(&a.L0[slot], flux);
            else n_splat += film_splat(sc, a.r.film_block, a.r.film_light, true, cr.el, flux, si.k);
        }
        __threadfence();
        if (atomicSub(&a.pending[slot], 1) == 1) { __threadfence(); n_splat += bd_finalize(a, slot, atomicAdd(&a.L0[slot], 0.f)); }
    }
    flush_counters(a.r.ctr, ctr, true);
    bd_flush_stats(a.r.ctr, n_splat, 0u, n_conn, 0u, overflow);
    {
        const unsigned m = __activemask();
        const unsigned nc = __reduce_add_sync(m, n_conn);
        if (nc && (threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) atomicAdd(&a.r.ctr->strategies[CLS], (unsigned long long)nc);
    }
}

} // namespace wt
