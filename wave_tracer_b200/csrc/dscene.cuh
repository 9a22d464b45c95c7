// dscene.cuh -- device-side scene functions of the B200 wave_tracer hot path: counter-based sampler, spectra,
// polarimetric algebra, beams, surface records, BSDFs, emitters, sensors, film splats.
// Each block cites the reference code whose behaviour it reproduces (paths relative to /root/reference).
#pragma once
#include "dtrav.cuh"
#include "dsobol.cuh"

namespace wt {

// ================================================================================================ sampler
// Counter-based stream (include/wtgpu.h "RNG contract"): draw d of (seed, pixel, sample) is lane d&3 of
// Philox4x32-10(key=seed, ctr=(d>>2, sample, pixel, stream)).  Replaces sampler::uniform_t (include/wt/sampler/uniform.hpp:36-50).
struct Sampler { uint32_t k0, k1, pixel, sample, d, stream; };      // stream: 0 (plt_path); plt_bdpt sub-streams, see include/wtgpu.h
WT_D uint32_t philox_lane(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t lane) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return lane == 0 ? c0 : lane == 1 ? c1 : lane == 2 ? c2 : c3;
}
// the same stream with the four lanes of a Philox block kept between draws (for code that draws many numbers in a row)
struct SamplerC { Sampler s; uint32_t blk; uint32_t c[4]; };
WT_D SamplerC samplerc(const Sampler& s) { SamplerC r; r.s = s; r.blk = 0xffffffffu; r.c[0] = r.c[1] = r.c[2] = r.c[3] = 0u; return r; }
WT_D float rnd(SamplerC& q) {
    const uint32_t blk = q.s.d >> 2;
    if (blk != q.blk) {
        uint32_t c0 = blk, c1 = q.s.sample, c2 = q.s.pixel, c3 = q.s.stream, k0 = q.s.k0, k1 = q.s.k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
            const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
            c0 = n0; c1 = l1; c2 = n2; c3 = l0;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        q.c[0] = c0; q.c[1] = c1; q.c[2] = c2; q.c[3] = c3; q.blk = blk;
    }
    const uint32_t lane = q.s.d & 3u;
    const uint32_t u = lane == 0u ? q.c[0] : lane == 1u ? q.c[1] : lane == 2u ? q.c[2] : q.c[3];
    ++q.s.d;
    return (float)(u >> 8) * (1.0f / 16777216.0f);
}
WT_D float rnd(Sampler& s) {
    // scene sampler = sobolld (include/wtgpu.h "sobolld contract"): stream carries the flag and the sensor's spp
    if (s.stream & kSobolStreamFlag) return sobol_draw(s.k0, s.k1, (uint64_t)s.pixel * (uint64_t)(s.stream & ~kSobolStreamFlag) + s.sample, s.d++);
    const uint32_t u = philox_lane(s.k0, s.k1, s.d >> 2, s.sample, s.pixel, s.stream, s.d & 3u);
    ++s.d;
    return (float)(u >> 8) * (1.0f / 16777216.0f);
}
WT_D V2 rnd2(Sampler& s) { const float a = rnd(s); const float b = rnd(s); return mk2(a, b); }
WT_D V3 rnd3(Sampler& s) { const float a = rnd(s); const float b = rnd(s); const float c = rnd(s); return mk3(a, b, c); }
WT_D int uniform_int_interval(Sampler& s, int start, int end) { return min(end - 1, int(rnd(s) * (end - start)) + start); }

// warps (include/wt/sampler/sampler.hpp:139-286)
WT_D V3 uniform_sphere(V2 u) { const float z = 1.f - 2.f * u.x; const float rr = sqrtf(fmaxf(0.f, 1.f - sqrf(z))); const float phi = kTwoPi * u.y; return mk3(rr * pm::cosf(phi), rr * pm::sinf(phi), z); }
WT_D V2 concentric_disk(V2 u) {
    const V2 o = 2.f * u - mk2(1.f, 1.f);
    float rr, th;
    if (o.x == 0.f && o.y == 0.f) { rr = 0.f; th = 0.f; }
    else if (fabsf(o.x) > fabsf(o.y)) { rr = o.x; th = kPi4 * (o.y / o.x); }
    else { rr = o.y; th = kPi2 - kPi4 * (o.x / o.y); }
    return rr * mk2(pm::cosf(th), pm::sinf(th));
}
WT_D V3 cosine_hemisphere(V2 u) { const V2 d = concentric_disk(u); return mk3(d.x, d.y, sqrtf(fmaxf(0.f, 1.f - sqrf(d.x) - sqrf(d.y)))); }
WT_D float cosine_hemisphere_pdf(float c) { return kInvPi * c; }
WT_D V3 uniform_cone(float sa, V2 u) {
    const float ctm = 1.f - kInvTwoPi * sa;
    const float ct = 1.f + u.x * (ctm - 1.f);
    const float st = sqrtf(fmaxf(0.f, 1.f - sqrf(ct)));
    const float phi = kTwoPi * u.y;
    return mk3(pm::cosf(phi) * st, pm::sinf(phi) * st, ct);
}
WT_D V2 normal2d(V2 u) { const float r = sqrtf(-2.f * pm::logf(1.f - u.x)); const float th = kTwoPi * u.y; return mk2(r * pm::cosf(th), r * pm::sinf(th)); }
WT_D V2 uniform_triangle(V2 u) { if (u.x + u.y > 1.f) u = mk2(1.f, 1.f) - u; return u; }

// sampling density with discrete flag (include/wt/sampler/density.hpp)
struct Pd { float v; bool disc; };
WT_D Pd pd_disc(float m) { Pd p; p.v = m; p.disc = true; return p; }
WT_D Pd pd_dens(float d) { Pd p; p.v = d; p.disc = false; return p; }

// ================================================================================================ spectra
WT_D C2 spectrum_value(const DScene& sc, int32_t id, float k) {
    if (id < 0) return mkc(1.f, 0.f);
    const wtgpu_spectrum s = sc.spectra[id];
    if (s.type == WTGPU_SPECTRUM_CONSTANT) return mkc(s.re, s.im);
    const float x = (k - s.k0) * s.inv_dk;
    if (!(x >= 0.f) || s.n == 0u) return mkc(0.f, 0.f);
    const uint32_t i0 = (uint32_t)x;
    if (i0 + 1u >= s.n) { if (x > (float)(s.n - 1u)) return mkc(0.f, 0.f); const float* p = sc.spectrum_data + 2u * (s.offset + s.n - 1u); return mkc(p[0], p[1]); }
    const float f = x - (float)i0;
    const float* p0 = sc.spectrum_data + 2u * (s.offset + i0);
    return mkc(mixf(p0[0], p0[2], f), mixf(p0[1], p0[3], f));
}
WT_D float spectrum_f(const DScene& sc, int32_t id, float k) { return spectrum_value(sc, id, k).re; }

// ================================================================================================ polarimetric
// Stokes (include/wt/interaction/polarimetric/stokes.hpp) / Mueller (mueller.hpp), column-major m[c*4+r] as glm.
struct Stokes { float s[4]; };
WT_D Stokes stokes_zero() { Stokes r; r.s[0] = r.s[1] = r.s[2] = r.s[3] = 0.f; return r; }
WT_D Stokes stokes_reorient(const Stokes& S, const Frame& cur, const Frame& nw) {      // stokes.hpp:152-175
    const V3 tl = to_local(cur, nw.t), bl = to_local(cur, nw.b);
    const M2 R = rotation2(mk2(1.f, 0.f), mk2(tl.x, tl.y));
    const V2 s12 = m2mul(R, m2mul(R, mk2(S.s[1], S.s[2])));
    Stokes r; r.s[0] = S.s[0]; r.s[1] = s12.x; r.s[2] = s12.y; r.s[3] = S.s[3];
    const V2 v = m2mul(R, mk2(0.f, 1.f));
    if (dot(v, mk2(bl.x, bl.y)) < 0.f) { r.s[2] = -r.s[2]; r.s[3] = -r.s[3]; }
    return r;
}
struct Mueller { float m[16]; };
WT_D Mueller mu_zero() { Mueller M; _Pragma("unroll") for (int i = 0; i < 16; ++i) M.m[i] = 0.f; return M; }
WT_D Mueller mu_identity() { Mueller M = mu_zero(); M.m[0] = M.m[5] = M.m[10] = M.m[15] = 1.f; return M; }
WT_D Mueller mu_flip() { Mueller M = mu_zero(); M.m[0] = M.m[5] = 1.f; M.m[10] = M.m[15] = -1.f; return M; }
WT_D Mueller mu_depol(float s) { Mueller M = mu_zero(); M.m[0] = s; return M; }
WT_D Mueller mu_scale(const Mueller& A, float s) { Mueller R; _Pragma("unroll") for (int i = 0; i < 16; ++i) R.m[i] = A.m[i] * s; return R; }
WT_D Mueller mu_div(const Mueller& A, float s) { Mueller R; _Pragma("unroll") for (int i = 0; i < 16; ++i) R.m[i] = A.m[i] / s; return R; }
WT_D Mueller mu_add(const Mueller& A, const Mueller& B) { Mueller R; _Pragma("unroll") for (int i = 0; i < 16; ++i) R.m[i] = A.m[i] + B.m[i]; return R; }
WT_D Mueller mu_mul(const Mueller& A, const Mueller& B) {     // glm mat4*mat4
    Mueller R;
    _Pragma("unroll") for (int c = 0; c < 4; ++c) _Pragma("unroll") for (int r = 0; r < 4; ++r)
        R.m[c * 4 + r] = A.m[r] * B.m[c * 4] + A.m[4 + r] * B.m[c * 4 + 1] + A.m[8 + r] * B.m[c * 4 + 2] + A.m[12 + r] * B.m[c * 4 + 3];
    return R;
}
WT_D Stokes mu_apply(const Mueller& M, const Stokes& S) {        // mueller.hpp:130-146
    Stokes r;
    _Pragma("unroll") for (int i = 0; i < 4; ++i) r.s[i] = fmaf(M.m[12 + i], S.s[3], fmaf(M.m[8 + i], S.s[2], fmaf(M.m[4 + i], S.s[1], M.m[i] * S.s[0])));
    return r;
}
WT_D Mueller mu_rotation(V2 t1, V2 t2) {                          // mueller.hpp:244-258 (incl. the final transpose)
    M2 R = rotation2(t1, t2);
    R = m2mm(R, R);
    Mueller T = mu_zero();
    T.m[0] = T.m[15] = 1.f;
    // T[1][1]=R[0][0]; T[2][1]=R[1][0]; T[1][2]=R[0][1]; T[2][2]=R[1][1]; then transpose
    T.m[1 * 4 + 1] = R.c0x; T.m[1 * 4 + 2] = R.c1x; T.m[2 * 4 + 1] = R.c0y; T.m[2 * 4 + 2] = R.c1y;
    return T;
}
WT_D Mueller mu_fresnel(C2 fs, C2 fp) {                           // mueller.hpp:294-309
    const float Rs = cnorm(fs), Rp = cnorm(fp);
    const float m00 = (Rs + Rp) / 2.f, m01 = (Rs - Rp) / 2.f;
    const C2 pc = fp * cconj(fs);
    Mueller M = mu_zero();
    // math layout [[m00,m01,0,0],[m01,m00,0,0],[0,0,m22,m23],[0,0,-m23,m22]] -> column-major storage
    M.m[0] = m00; M.m[1] = m01; M.m[4] = m01; M.m[5] = m00;
    M.m[10] = pc.re; M.m[2 * 4 + 3] = -pc.im; M.m[3 * 4 + 2] = pc.im; M.m[15] = pc.re;
    return M;
}
WT_D Stokes mu_apply_frames3(const Mueller& M, const Stokes& S, const Frame& Sin, const Frame& Min) {   // mueller.hpp:155-165
    if (S.s[1] == 0.f && S.s[2] == 0.f && S.s[3] == 0.f) return mu_apply(M, S);
    return mu_apply(M, stokes_reorient(S, Sin, Min));
}
WT_D Stokes mu_apply_frames5(const Mueller& M, const Stokes& S, const Frame& Sin, const Frame& Min, const Frame& Sout, const Frame& Mout) {  // :175-185
    return stokes_reorient(mu_apply(M, stokes_reorient(S, Sin, Min)), Mout, Sout);
}
WT_D Mueller mu_change_incident_frame(const Mueller& M, const Frame& oldf, const Frame& newf) {          // :191-203
    const V3 tl = to_local(oldf, newf.t);
    Mueller R = mu_rotation(mk2(tl.x, tl.y), mk2(1.f, 0.f));
    if (handness(oldf) != handness(newf)) R = mu_mul(R, mu_flip());
    return mu_mul(M, R);
}
WT_D Mueller mu_compose(const Mueller& M1, const Mueller& M2, const Frame& M1in, const Frame& M2out) {   // :402-415
    const V3 tl = to_local(M1in, M2out.t);
    Mueller R = mu_rotation(mk2(tl.x, tl.y), mk2(1.f, 0.f));
    if (handness(M1in) != handness(M2out)) R = mu_mul(mu_flip(), R);
    return mu_mul(mu_mul(M1, R), M2);
}

// ---- Fresnel (include/wt/interaction/fresnel.hpp)
WT_D V3 reflect_z(V3 w) { return 2.f * (dot(w, mk3(0.f, 0.f, 1.f)) * mk3(0.f, 0.f, 1.f)) - w; }
struct Refr { V3 t; float cost, eta; bool tir; };
WT_D Refr refract(float eta_12, V3 w, V3 n) {
    Refr r;
    const float wn = dot(w, n);
    eta_12 = wn > 0.f ? eta_12 : 1.f / eta_12;
    r.eta = eta_12;
    const float c2 = 1.f - sqrf(eta_12) * (1.f - sqrf(wn));
    if (c2 >= 0.f) { r.cost = sqrtf(c2); r.t = normalize(eta_12 * (wn * n - w) - r.cost * (wn >= 0.f ? n : -n)); r.tir = false; return r; }
    r.t = mk3(0.f, 0.f, 1.f); r.cost = 0.f; r.tir = true; return r;
}
struct Fresnel { V3 t; C2 eta; float Z; C2 rs, rp, ts, tp; float Ts, Tp; };
WT_D Fresnel fresnel(C2 eta_12, V3 w, V3 n) {                    // fresnel.hpp:74-117
    Fresnel f;
    if (eta_12.re == 1.f && eta_12.im == 0.f) { f.t = -w; f.eta = eta_12; f.Z = 1.f; f.rs = f.rp = mkc(0.f, 0.f); f.ts = f.tp = mkc(1.f, 0.f); f.Ts = f.Tp = 1.f; return f; }
    const float ac = fabsf(dot(w, n));
    const Refr r = refract(eta_12.re, w, n);
    if (ac == 0.f || r.tir) { f.t = mk3(0.f, 0.f, 1.f); f.eta = mkc(r.eta, 0.f); f.Z = 1.f; f.rs = f.rp = mkc(1.f, 0.f); f.ts = f.tp = mkc(0.f, 0.f); f.Ts = f.Tp = 0.f; return f; }
    const C2 e = mkc(r.eta, 0.f);
    f.rs = (e * ac - mkc(r.cost, 0.f)) / (e * ac + mkc(r.cost, 0.f));
    f.rp = (mkc(ac, 0.f) - e * r.cost) / (mkc(ac, 0.f) + e * r.cost);
    f.ts = f.rs + mkc(1.f, 0.f);
    f.tp = (f.rp + mkc(1.f, 0.f)) * e;
    f.Z = cabsf_(mkc(r.cost, 0.f) / (e * ac));
    f.t = r.t; f.eta = e;
    f.Ts = fminf(1.f, f.Z * cnorm(f.ts)); f.Tp = fminf(1.f, f.Z * cnorm(f.tp));
    return f;
}
WT_D void fresnel_reflection(C2 eta_12, V3 w, V3 n, C2& rs, C2& rp) {   // fresnel.hpp:128-144
    const float wn = dot(w, n);
    if ((eta_12.re == 1.f && eta_12.im == 0.f) || wn < 0.f) { rs = rp = mkc(0.f, 0.f); return; }
    const C2 t2 = mkc(1.f, 0.f) - (1.f - sqrf(wn)) * (eta_12 * eta_12);
    const C2 t = csqrt_(t2);
    const C2 i = mkc(wn, 0.f);
    rs = (eta_12 * i - t) / (eta_12 * i + t);
    rp = (i - eta_12 * t) / (i + eta_12 * t);
}
WT_D Mueller mu_fresnel_dir(C2 eta_12, bool reflection, V3 w, V3 n) {   // mueller.hpp:318-344
    if (reflection) { C2 rs, rp; fresnel_reflection(eta_12, w, n, rs, rp); return mu_fresnel(rs, rp); }
    const Fresnel f = fresnel(eta_12, w, n);
    return mu_scale(mu_fresnel(f.ts, f.tp), f.Z);
}

// ================================================================================================ beams
constexpr float kEnvelope = 3.f;             // gaussian_wavefront.hpp:26
constexpr float kMajorToZ = 2.f;             // beam_generic.hpp:50
WT_D float k_times_len(float k, float len) { return k * len * 1000.f; }
WT_D float wavenum_to_wavelen(float k) { return (kTwoPi / k) * 0.001f; }

struct Footprint { V2 x; float la, lb; };
struct Surface {            // intersection_surface_t (interaction/intersection.hpp:34-160)
    V3 wp; V2 uv; Footprint fp; uint32_t tuid; bool has_shape; Frame geo, shading;
};
WT_D V3 s_direction(const Surface& s, V3 w) {
    const V3 crs = cross(w, s.shading.n);
    const float l2 = length2(crs);
    const V3 ret = l2 < 1e-14f ? s.shading.t : crs / sqrtf(l2);
    return dot(w, s.shading.n) < 0.f ? -ret : ret;
}
WT_D Frame sp_frame(const Surface& s, V3 w) {
    const V3 sd = s_direction(s, w);
    const V3 p = cross(sd, w);
    Frame f; f.t = sd; f.b = dot(w, s.shading.n) < 0.f ? -p : p; f.n = w; return f;
}

struct Beam {               // beam_t (beam/beam.hpp:255-518); forward: Stokes in M.m[0..3]
    Cone env; float sid; float k; bool fwd;
    Mueller M; float scale; Frame frame;
};
WT_D float beam_intensity(const Beam& b) { return b.fwd ? b.M.m[0] : b.M.m[0] * b.scale; }
WT_D void beam_mul(Beam& b, float f) { if (b.fwd) { b.M.m[0] *= f; b.M.m[1] *= f; b.M.m[2] *= f; b.M.m[3] *= f; } else b.scale *= f; }
WT_D void beam_div(Beam& b, float f) { if (b.fwd) { b.M.m[0] /= f; b.M.m[1] /= f; b.M.m[2] /= f; b.M.m[3] /= f; } else b.scale /= f; }
WT_D Stokes beam_stokes(const Beam& b) { Stokes s; s.s[0] = b.M.m[0]; s.s[1] = b.M.m[1]; s.s[2] = b.M.m[2]; s.s[3] = b.M.m[3]; return s; }
WT_D void beam_set_stokes(Beam& b, const Stokes& s) { b.M.m[0] = s.s[0]; b.M.m[1] = s.s[1]; b.M.m[2] = s.s[2]; b.M.m[3] = s.s[3]; }
WT_D V3 beam_footprint(const Beam& b, float dist) { const V2 a = cone_axes(b.env, dist); return mk3(a.x, a.y, kMajorToZ * a.x); }

// sourcing (beam/beam_geometry.hpp:32-342), surface-less
struct Sourcing { float l; float ta; };      // isotropic initial spatial length, tan(alpha)
WT_D float mub_tan_alpha(float l, float k) { return l > 0.f ? sqrtf(0.25f) * sqrf(kEnvelope) / k_times_len(k, l) : 0.f; }
WT_D Sourcing source_extent(float spatial_extent, float ta) { Sourcing s; s.l = sqrtf(spatial_extent); s.ta = ta; return s; }
WT_D void enlarge(float& spatial_extent, float& ta, float scale) { if (scale == 1.f) return; spatial_extent = spatial_extent * sqrf(scale); ta = ta * scale; }
WT_D Beam beam_make(bool fwd, V3 o, V3 d, float s, float k, Sourcing sg) {
    Beam b; b.fwd = fwd; b.k = k; b.sid = 0.f;
    b.env = cone_iso(o, d, sg.ta, sg.l);
    b.frame = cone_frame(b.env);
    if (fwd) { b.M = mu_zero(); b.M.m[0] = s; b.scale = 0.f; } else { b.M = mu_identity(); b.scale = s; }
    return b;
}
WT_D void beam_add(Beam& b, const Beam& o) {        // operator+= (beam.hpp:95-98, 200-203)
    if (b.fwd) { const Stokes r = stokes_reorient(beam_stokes(o), o.frame, b.frame); _Pragma("unroll") for (int i = 0; i < 4; ++i) b.M.m[i] += r.s[i]; }
    else b.M = mu_add(b.M, mu_change_incident_frame(o.M, o.frame, b.frame));
}
WT_D Footprint surface_footprint_static(const Beam& b, const Surface& s, float z) {    // beam_generic.hpp:171-193
    const V3 ls = beam_footprint(b, z);
    const V3 x = to_local(s.geo, b.env.x);
    Footprint f;
    if (x.x != 0.f || x.y != 0.f) { f.x = normalize(mk2(x.x, x.y)); f.la = ls.x; f.lb = ls.y; }
    else { const float avg = (ls.x + ls.y) / 2.f; f.x = mk2(1.f, 0.f); f.la = avg; f.lb = avg; }
    return f;
}
WT_DN void beam_transform_surface(Beam& b, const Surface& s, V3 wo, const Mueller& bsdfM, float weight) {   // beam.hpp:379-397
    const V2 fa = s.fp.x * s.fp.la, fb = mk2(-s.fp.x.y, s.fp.x.x) * s.fp.lb;
    const V3 wa = to_world(s.geo, fa), wb = to_world(s.geo, fb);
    float nsid;
    b.env = cone_through_ellipse(wa, wb, s.geo.n, s.wp, wo, b.env.ta, &nsid);
    const Frame after = cone_frame(b.env);
    const Mueller op = mu_scale(bsdfM, weight);
    if (b.fwd) {    // beam.hpp:53-66
        const Frame SPin = sp_frame(s, b.frame.n), SPout = sp_frame(s, wo);
        beam_set_stokes(b, mu_apply_frames5(op, beam_stokes(b), b.frame, SPin, after, SPout));
        b.frame = after;
    } else {        // beam.hpp:163-173
        const Frame SPin = sp_frame(s, wo), SPout = sp_frame(s, b.frame.n);
        b.M = mu_compose(b.M, op, b.frame, SPout);
        b.frame = SPin;
    }
    b.sid = nsid;
}
WT_DN void beam_transform_region(Beam& b, V3 wp, float dist, V3 wo, float weight) {     // beam.hpp:407-425
    const V3 axes = beam_footprint(b, dist);
    b.env = cone_through_ellipsoid(axes, cone_frame(b.env), wp, wo, b.env.ta);
    beam_mul(b, weight);
    b.frame = cone_frame(b.env);
    b.sid = 0.f;
}
WT_D void beam_transform_restart(Beam& b, V3 wp, float dist) {   // beam.hpp:464-471
    b.env.o = wp; b.env.x0 = b.env.x0 + dist * b.env.ta; b.sid = 0.f;
}
WT_D Stokes integrate_beams(const Beam& S, const Beam& I) {       // beam.hpp:562-603
    if (beam_intensity(S) == 0.f || beam_intensity(I) == 0.f) return stokes_zero();
    Stokes r = mu_apply_frames3(S.M, beam_stokes(I), I.frame, S.frame);
    _Pragma("unroll") for (int i = 0; i < 4; ++i) r.s[i] *= S.scale;
    return r;
}

// ================================================================================================ surfaces
WT_DN Surface make_surface(const DScene& sc, uint32_t tuid, V2 bary, V3 centre) {       // src/interaction/intersection.cpp:33-70
    const wtgpu_tri_shading sh = sc.tri_shading[tuid];
    const float w2 = 1.f - bary.x - bary.y;
    Surface s;
    s.wp = centre; s.tuid = tuid; s.has_shape = true;
    s.uv = sh.has_uv ? mk2(bary.x * sh.uv0[0] + bary.y * sh.uv1[0] + w2 * sh.uv2[0], bary.x * sh.uv0[1] + bary.y * sh.uv1[1] + w2 * sh.uv2[1]) : mk2(0.f, 0.f);
    const V3 n = normalize(mk3(bary.x * sh.n0[0] + bary.y * sh.n1[0] + w2 * sh.n2[0], bary.x * sh.n0[1] + bary.y * sh.n1[1] + w2 * sh.n2[1], bary.x * sh.n0[2] + bary.y * sh.n1[2] + w2 * sh.n2[2]));
    const V3 dpdu = mk3(sh.dpdu);
    const Tri3 tr = load_tri(sc, tuid);
    s.geo = shading_frame(tr.n, dpdu);
    s.shading = shading_frame(n, dpdu);
    s.fp.x = mk2(1.f, 0.f); s.fp.la = s.fp.lb = 0.f;
    return s;
}
WT_D Surface make_surface_at_bary(const DScene& sc, uint32_t tuid, V2 bary) {
    const float w2 = 1.f - bary.x - bary.y;
    const Tri3 t = load_tri(sc, tuid);
    const V3 p = mk3(bary.x * t.a.x + bary.y * t.b.x + w2 * t.c.x, bary.x * t.a.y + bary.y * t.b.y + w2 * t.c.y, bary.x * t.a.z + bary.y * t.b.z + w2 * t.c.z);
    return make_surface(sc, tuid, bary, p);
}
WT_D Surface make_dummy_surface(V3 n, V3 p) {
    Surface s; s.wp = p; s.geo = orthogonal_frame(n); s.shading = s.geo; s.has_shape = false; s.tuid = WTGPU_INVALID_IDX; s.uv = mk2(0.f, 0.f);
    s.fp.x = mk2(1.f, 0.f); s.fp.la = s.fp.lb = 0.f; return s;
}
WT_D V3 tri_fp_errors(V3 a, V3 b, V3 c, V3 ro) {          // intersection.cpp:149-170
    const float c0 = 3e-6f, c1 = 5e-6f, c2 = 3e-6f;
    const V3 v0 = vabs(a), e1 = vabs(b - a), e2 = vabs(c - a);
    const float extent = vmaxel(e1 + e2 + vabs(e1 - e2));
    return ((c0 + c2) * v0 + mk3(c1 * extent, c1 * extent, c1 * extent)) + (c1 + c2) * vabs(ro);
}
// vertex geometry variant (integrator/traversal.hpp:251-268): 0 none, 1 point, 2 surface (tuid), 3 edge (edge id)
struct Geo { uint32_t kind; V3 p; uint32_t id; };
WT_D Geo geo_point(V3 p) { Geo g; g.kind = 1u; g.p = p; g.id = WTGPU_INVALID_IDX; return g; }
WT_D Geo geo_surface(V3 p, uint32_t tuid, bool has_shape) { Geo g; g.kind = has_shape ? 2u : 1u; g.p = p; g.id = tuid; return g; }
WT_D Geo geo_edge(V3 p, uint32_t e) { Geo g; g.kind = 3u; g.p = p; g.id = e; return g; }
WT_D V3 offseted_ray_origin(const DScene& sc, const Geo& g, V3 ro, V3 rd) {      // intersection.cpp:172-211
    if (g.kind == 2u) {
        const Tri3 t = load_tri(sc, g.id);
        const V3 err = tri_fp_errors(t.a, t.b, t.c, ro);
        const float od = dot(err, vabs(t.n));
        const V3 off = od * t.n;
        return ro + (dot(rd, off) >= 0.f ? off : -off);
    }
    if (g.kind == 3u) {
        const wtgpu_edge e = sc.edges[g.id];
        const V3 t1 = mk3(e.t1), t2 = mk3(e.t2);
        const bool has2 = e.tri2 != WTGPU_INVALID_IDX;
        V3 dir;
        if (!has2) dir = -t1; else { const V3 v = t1 + t2; dir = length2(v) > 1e-14f ? -normalize(v) : -t2; }
        const Tri3 ta = load_tri(sc, e.tri1);
        float dd = dot(tri_fp_errors(ta.a, ta.b, ta.c, ro), vabs(t1));
        if (has2) { const Tri3 tb = load_tri(sc, e.tri2); dd = fmaxf(dd, dot(tri_fp_errors(tb.a, tb.b, tb.c, ro), vabs(t2))); }
        return ro + dd * dir;
    }
    return ro;
}
// integrator::shadow (traversal.hpp:319-333)
#ifndef WT_SHADOW_INLINE
#define WT_SHADOW_INLINE __device__ __forceinline__
#endif
WT_SHADOW_INLINE bool shadow_between(const DScene& sc, const Geo& a, const Geo& b, Counters& ctr) {
    const V3 d0 = normalize(b.p - a.p);
    const V3 o = offseted_ray_origin(sc, a, a.p, d0);
    const V3 t = offseted_ray_origin(sc, b, b.p, -d0);
    const float dist = length(t - o);
    const V3 d = (t - o) / dist;
    return shadow_ray(sc, o, d, mkr(0.f, dist), ctr);
}

// ================================================================================================ BSDFs
struct BsdfSample { bool valid; V3 wo; Pd dpd; C2 eta; Mueller M; };
struct BsdfQuery { float k; bool fwd; uint32_t lobes; };

// resolve wrappers down to a leaf bsdf for wavenumber k; accumulates two_sided flips and scale factors
struct ResolvedBsdf { int32_t id; bool two_sided; float scale; };
WT_D ResolvedBsdf resolve_bsdf(const DScene& sc, int32_t id, float k) {
    ResolvedBsdf r; r.id = id; r.two_sided = false; r.scale = 1.f;
    for (int it = 0; it < 8 && r.id >= 0; ++it) {
        const wtgpu_bsdf b = sc.bsdfs[r.id];
        if (b.type == WTGPU_BSDF_TWO_SIDED) { r.two_sided = true; r.id = b.child; }
        else if (b.type == WTGPU_BSDF_SCALE) { r.scale *= spectrum_f(sc, b.spec[0], k); r.id = b.child; }
        else if (b.type == WTGPU_BSDF_COMPOSITE) {
            int32_t c = -1;
            for (uint32_t i = 0; i < b.n_bins; ++i) { const wtgpu_bsdf_bin bin = sc.bsdf_bins[b.bin_first + i]; if (bin.kmin <= k && k < bin.kmax) { c = bin.child; break; } }
            r.id = c;
        } else break;
    }
    return r;
}
// NOTE on wrapper order: two_sided flips depend only on sign(wi.z) and scale is linear, so hoisting them out of the
// recursion (two_sided.cpp:24-60, scale.hpp) is exact as long as a two_sided wrapper is outermost of the flips it applies to,
// which holds for every nesting the loader can produce (twosided(scale(x)), twosided(composite(x)), ...).

struct FractalP { float T, s2n, alpha; };
WT_D FractalP fractal_params(const DScene& sc, const wtgpu_bsdf& b, float k) {     // surface_profile/fractal.hpp:67-110
    FractalP p;
    const float gamma = b.gamma;
    if (b.profile_type == WTGPU_PROFILE_FRACTAL_ROUGHNESS) {
        const float rough = spectrum_f(sc, b.prof_spec[0], k);
        const float meank = kTwoPi / 550e-6f;
        const float a2 = sqrf(clampf_(rough, 0.f, .75f));
        p.T = fminf(70.f * 70.f, (1.f - a2) / (4.f * sqrf(meank) * a2));
        p.alpha = sqrf(rough / 9.f);
    } else {
        p.T = spectrum_f(sc, b.prof_spec[0], k);
        p.alpha = sqrf(spectrum_f(sc, b.prof_spec[1], k));
    }
    const float x = 1.f + k * k * p.T;
    const float pw = gamma == 3.f ? x : pm::powf(x, (gamma - 1.f) / 2.f);
    p.s2n = 1.f / (1.f - 1.f / pw);
    return p;
}
WT_D float fractal_psd(const wtgpu_bsdf& b, const FractalP& p, V2 z, float k) {
    const float gamma = b.gamma;
    const float x = 1.f + p.T * dot(z, z);
    const float pw = gamma == 3.f ? (x * x) : pm::powf(x, (gamma + 1.f) / 2.f);
    return p.s2n * (kInvTwoPi * k * k * (gamma - 1.f) * p.T * (1.f / pw));
}
// ---- gaussian profile (include/wt/interaction/surface_profile/gaussian.hpp:28-255)
WT_D bool is_gaussian_profile(const wtgpu_bsdf& b) { return b.profile_type == WTGPU_PROFILE_GAUSSIAN || b.profile_type == WTGPU_PROFILE_GAUSSIAN_SIGMA; }
struct GaussP { float sigma2, s2n, alpha; };
WT_D GaussP gaussian_params(const DScene& sc, const wtgpu_bsdf& b, float k) {      // gaussian.hpp:95-119
    GaussP p;
    if (b.profile_type == WTGPU_PROFILE_GAUSSIAN) {
        const float rough = spectrum_f(sc, b.prof_spec[0], k);
        const float meank = kTwoPi / 550e-6f;
        const float a2 = sqrf(clampf_(rough, 0.f, .75f));
        p.sigma2 = 1.f / fminf(70.f * 70.f, (1.f - a2) / (4.f * sqrf(meank) * a2));
        p.alpha = sqrf(rough / 9.f);
    } else {
        p.sigma2 = sqrf(spectrum_f(sc, b.prof_spec[0], k));
        p.alpha = p.sigma2;
    }
    p.s2n = 1.f / (1.f - pm::expf(-(k * k / 2.f / p.sigma2)));
    return p;
}
WT_D float gaussian_psd(const GaussP& p, V2 z, float k) {                          // gaussian.hpp:121-130
    const float z2 = dot(z, z);
    const float e = pm::expf(-(z2 / 2.f / p.sigma2));
    return e <= 1.1920929e-7f ? 0.f : p.s2n * (kInvTwoPi / p.sigma2 * k * k * e);
}
WT_D float boxmueller_max_phi(float r, float l) {                                   // gaussian.hpp:43-49, 70-76
    const float eps = 1.1920929e-7f;
    return (r < eps || l < eps) ? kPi : fmaxf(1e-2f, pm::acosf(clampf_((sqrf(r) + sqrf(l) - 1.f) / (2.f * r * l), -1.f, 1.f)));
}
WT_D float boxmueller_truncated_pdf(V2 wo, V2 mean, float sigma2) {               // gaussian.hpp:58-79
    const float l = sqrtf(fminf(1.f, dot(mean, mean)));
    const float coso = sqrtf(fmaxf(0.f, 1.f - dot(mean, mean)));
    wo = wo - mean;
    const float r2 = dot(wo, wo);
    const float x = pm::expf(-.5f * r2 / sigma2);
    const float r = sqrtf(r2);
    return .5f * x / (boxmueller_max_phi(r, l) * sigma2) * coso;
}
WT_D bool profile_delta_only(const DScene& sc, const wtgpu_bsdf& b, float k) {
    if (b.profile_type == WTGPU_PROFILE_DIRAC) return true;
    return spectrum_f(sc, b.prof_spec[0], k) == 0.f;
}
WT_D float profile_alpha(const DScene& sc, const wtgpu_bsdf& b, V3 wi, V3 wo, float k) {
    if (b.profile_type == WTGPU_PROFILE_DIRAC) return 1.f;
    const float palpha = is_gaussian_profile(b) ? gaussian_params(sc, b, k).alpha : fractal_params(sc, b, k).alpha;
    return pm::expf(-(sqrf((fabsf(wi.z) + fabsf(wo.z)) * k) * palpha));
}
WT_D float profile_psd(const DScene& sc, const wtgpu_bsdf& b, V3 wi, V3 wo, float k) {
    if (b.profile_type == WTGPU_PROFILE_DIRAC) return 0.f;
    if (is_gaussian_profile(b)) return gaussian_psd(gaussian_params(sc, b, k), k * (mk2(wi.x, wi.y) + mk2(wo.x, wo.y)), k);
    const FractalP p = fractal_params(sc, b, k);
    return fractal_psd(b, p, k * (mk2(wi.x, wi.y) + mk2(wo.x, wo.y)), k);
}
WT_D float profile_pdf(const DScene& sc, const wtgpu_bsdf& b, V3 wi, V3 wo, float k) {   // fractal.hpp:205-226
    if (b.profile_type == WTGPU_PROFILE_DIRAC) return 0.f;
    if (is_gaussian_profile(b)) return boxmueller_truncated_pdf(mk2(wo.x, wo.y), mk2(-wi.x, -wi.y), gaussian_params(sc, b, k).sigma2 / (k * k));   // gaussian.hpp:240-250
    const FractalP p = fractal_params(sc, b, k);
    const V2 zk = mk2(wi.x, wi.y) + mk2(wo.x, wo.y);
    const float fk = length(zk);
    const float s = sqrtf(fmaxf(0.f, 1.f - sqrf(wi.z)));
    const float phi_max = (fk == 0.f || s == 0.f) ? kPi : pm::acosf(clampf_((sqrf(fk) + sqrf(s) - 1.f) / (2.f * fk * s), -1.f, 1.f));
    const float psd = fractal_psd(b, p, zk * k, k);
    const float w = kInvPi * phi_max;
    return w > 1e-2f ? 1.f / w * fabsf(wo.z) * psd : 0.f;
}
struct ProfSample { V3 wo; float pdf, psd; };
WT_D ProfSample profile_sample(const DScene& sc, const wtgpu_bsdf& b, V3 wi, float k, Sampler& smp) {   // surface_profile/fractal.cpp:27-69
    ProfSample r;
    if (b.profile_type == WTGPU_PROFILE_DIRAC) { r.wo = mk3(0.f, 0.f, 1.f); r.pdf = r.psd = 0.f; return r; }
    if (is_gaussian_profile(b)) {       // gaussian.hpp:210-235 with the truncated Box-Mueller transform of 28-56
        const GaussP p = gaussian_params(sc, b, k);
        const float s2 = p.sigma2 / (k * k), eps = 1.1920929e-7f;
        const V2 mean = mk2(-wi.x, -wi.y);
        const V2 u2 = rnd2(smp);
        const float l = sqrtf(fminf(1.f, dot(mean, mean)));
        const float coso = sqrtf(fmaxf(0.f, 1.f - dot(mean, mean)));
        const float phi_i = (mean.x != 0.f || mean.y != 0.f) ? pm::atan2f(mean.y, mean.x) : 0.f;
        const float s = pm::expf(-.5f * sqrf(1.f + l) / s2);
        const float x = (1.f - s) * fmaxf(eps, u2.x) + s;
        const float rr = sqrtf(-2.f * s2 * pm::logf(x));
        const float max_phi = boxmueller_max_phi(rr, l);
        const float phi = phi_i + kPi + max_phi * (2.f * u2.y - 1.f);
        const V2 pt = rr * mk2(pm::cosf(phi), pm::sinf(phi));
        const V2 wo2 = pt + mean;
        r.pdf = .5f * x / (max_phi * s2) * coso;
        r.psd = gaussian_psd(p, k * (wo2 - mean), k);
        const float z = sqrtf(fmaxf(0.f, 1.f - dot(wo2, wo2)));
        r.wo = mk3(wo2.x, wo2.y, wi.z >= 0.f ? z : -z);
        return r;
    }
    const float gamma = b.gamma;
    const FractalP p = fractal_params(sc, b, k);
    const float s = sqrtf(fmaxf(0.f, 1.f - sqrf(wi.z)));
    const float phi_i = s > 0.f ? pm::atan2f(wi.y, wi.x) : 0.f;
    const float sqrtT = sqrtf(p.T);
    const V2 u2 = rnd2(smp);
    const float k2T = sqrf(k) * p.T;
    const float Mv = 1.f - pm::powf(1.f + k2T * sqrf(1.f + s), -(gamma - 1.f) / 2.f);
    const float f = sqrtf(pm::powf(1.f - Mv * u2.x, -2.f / (gamma - 1.f)) - 1.f) / sqrtT;
    const float fk = f / k;
    const float phi_max = (f == 0.f || s == 0.f) ? kPi : pm::acosf(clampf_((sqrf(fk) + sqrf(s) - 1.f) / (2.f * fk * s), -1.f, 1.f));
    const float phi_f = phi_i + (2.f * u2.y - 1.f) * phi_max;
    const V2 zeta = f * mk2(pm::cosf(phi_f), pm::sinf(phi_f));
    const V2 wo = zeta / k - mk2(wi.x, wi.y);
    const float z = sqrtf(fmaxf(0.f, 1.f - dot(wo, wo)));
    r.psd = fractal_psd(b, p, zeta, k);
    const float w = kInvPi * phi_max;
    r.pdf = w > 1e-2f ? z * r.psd / w : 0.f;
    r.wo = mk3(wo.x, wo.y, wi.z >= 0.f ? z : -z);
    return r;
}

WT_D V3 flipz(V3 w, float z) { return z >= 0.f ? w : mk3(w.x, w.y, -w.z); }
WT_D V3 flip_wo(V3 wo, float eta) {                               // surface_spm.cpp:27-34
    const float sc = wo.z > 0.f ? eta : 1.f / eta;
    const V2 xy = mk2(wo.x, wo.y) * sc;
    const float l2 = dot(xy, xy);
    return l2 > 1.f ? mk3(1.f, 0.f, 0.f) : mk3(xy.x, xy.y, (wo.z > 0.f ? -1.f : 1.f) * sqrtf(fmaxf(0.f, 1.f - l2)));
}
WT_D bool ior_has_transmission(C2 ior) { return sqrf(fabsf(ior.im)) / cnorm(ior) <= 1e-2f; }
WT_D float opt_scale(const DScene& sc, int32_t id, float k) { return id >= 0 ? spectrum_f(sc, id, k) : 1.f; }

WT_D bool bsdf_is_delta_only(const DScene& sc, int32_t id, float k) {
    const ResolvedBsdf r = resolve_bsdf(sc, id, k);
    if (r.id < 0) return true;
    const wtgpu_bsdf b = sc.bsdfs[r.id];
    if (b.type == WTGPU_BSDF_DIFFUSE) return false;
    if (b.type == WTGPU_BSDF_DIELECTRIC) return true;
    return profile_delta_only(sc, b, k);
}

WT_DN Mueller bsdf_f(const DScene& sc, int32_t id, V3 wi, V3 wo, const BsdfQuery& q) {
    const ResolvedBsdf r = resolve_bsdf(sc, id, q.k);
    if (r.id < 0) return mu_zero();
    if (r.two_sided) { const float z = wi.z; wi = flipz(wi, z); wo = flipz(wo, z); }
    const wtgpu_bsdf b = sc.bsdfs[r.id];
    Mueller M = mu_zero();
    if (b.type == WTGPU_BSDF_DIFFUSE) {                 // diffuse.cpp:23-36
        const float refl = clampf_(spectrum_f(sc, b.spec[0], q.k), 0.f, 1.f);
        M = mu_depol(((q.lobes & 1u) && wi.z > 0.f && wo.z > 0.f) ? wo.z * kInvPi * refl : 0.f);
    } else if (b.type == WTGPU_BSDF_SURFACE_SPM) {      // surface_spm.cpp:40-77
        const bool is_scatter = (q.lobes & 2u) && !profile_delta_only(sc, b, q.k);
        const bool is_refl = wi.z * wo.z >= 0.f;
        const C2 eta = spectrum_value(sc, b.spec[0], q.k) / spectrum_value(sc, b.spec[1], q.k);
        const bool has_tr = ior_has_transmission(eta);
        if (wi.z == 0.f || wo.z == 0.f || !is_scatter || (!is_refl && !has_tr)) return mu_zero();
        const V3 awo = is_refl ? wo : flip_wo(wo, eta.re);
        const float alpha = profile_alpha(sc, b, wi, awo, q.k);
        float J = 1.f;
        if (!is_refl && !q.fwd) J = sqrf(wi.z < 0.f ? 1.f / eta.re : eta.re);
        const float scl = is_refl ? opt_scale(sc, b.spec[2], q.k) : opt_scale(sc, b.spec[3], q.k);
        const V3 h = wi + awo;
        const V3 m = normalize(wi.z < 0.f ? -h : h);
        const Mueller F = mu_fresnel_dir(mkc(eta.re, 0.f), is_refl, wi, m);
        const float psd = profile_psd(sc, b, wi, awo, q.k);
        M = mu_scale(F, (1.f - alpha) * J * fabsf(wo.z) * psd * scl);
    }
    if (r.scale != 1.f) M = mu_scale(M, r.scale);
    return M;
}

WT_DN float bsdf_pdf(const DScene& sc, int32_t id, V3 wi, V3 wo, const BsdfQuery& q) {
    const ResolvedBsdf r = resolve_bsdf(sc, id, q.k);
    if (r.id < 0) return 0.f;
    if (r.two_sided) { const float z = wi.z; wi = flipz(wi, z); wo = flipz(wo, z); }
    const wtgpu_bsdf b = sc.bsdfs[r.id];
    if (b.type == WTGPU_BSDF_DIFFUSE) return ((q.lobes & 1u) && wi.z > 0.f && wo.z > 0.f) ? cosine_hemisphere_pdf(wo.z) : 0.f;
    if (b.type == WTGPU_BSDF_SURFACE_SPM) {             // surface_spm.cpp:172-200
        const bool is_refl = wi.z * wo.z >= 0.f;
        const C2 eta = spectrum_value(sc, b.spec[0], q.k) / spectrum_value(sc, b.spec[1], q.k);
        const bool has_tr = ior_has_transmission(eta);
        if (wi.z == 0.f || wo.z == 0.f || !(q.lobes & 2u) || (!is_refl && !has_tr)) return 0.f;
        const V3 awo = is_refl ? wo : flip_wo(wo, eta.re);
        const float alpha = profile_alpha(sc, b, wi, wi, q.k);
        const float pspec = (q.lobes & 1u) ? alpha : 0.f;
        const Fresnel fr = fresnel(mkc(eta.re, 0.f), wi, mk3(0.f, 0.f, 1.f));
        const float ptr = (fr.Ts + fr.Tp) / 2.f;
        return (1.f - pspec) * profile_pdf(sc, b, wi, awo, q.k) * (is_refl ? 1.f - ptr : ptr);
    }
    return 0.f;
}

WT_DN BsdfSample bsdf_sample(const DScene& sc, int32_t id, V3 wi, const BsdfQuery& q, Sampler& smp) {
    BsdfSample out; out.valid = false; out.wo = mk3(0.f, 0.f, 1.f); out.dpd = pd_disc(0.f); out.eta = mkc(1.f, 0.f); out.M = mu_zero();
    const ResolvedBsdf r = resolve_bsdf(sc, id, q.k);
    if (r.id < 0) return out;
    const float wiz0 = wi.z;
    if (r.two_sided) wi = flipz(wi, wiz0);
    const wtgpu_bsdf b = sc.bsdfs[r.id];
    const V3 nz = mk3(0.f, 0.f, 1.f);
    if (b.type == WTGPU_BSDF_DIFFUSE) {                 // diffuse.cpp:38-61
        if (wi.z <= 0.f) return out;
        const float refl = clampf_(spectrum_f(sc, b.spec[0], q.k), 0.f, 1.f);
        out.wo = cosine_hemisphere(rnd2(smp));
        out.dpd = pd_dens(cosine_hemisphere_pdf(out.wo.z));
        out.M = mu_depol(refl); out.valid = true;
    } else if (b.type == WTGPU_BSDF_DIELECTRIC) {       // dielectric.cpp:26-72
        const C2 er = spectrum_value(sc, b.spec[0], q.k) / spectrum_value(sc, b.spec[1], q.k);
        const Fresnel fr = fresnel(mkc(er.re, 0.f), wi, nz);
        const float T = (fr.Ts + fr.Tp) / 2.f;
        const bool is_refl = rnd(smp) >= T;
        out.wo = is_refl ? reflect_z(wi) : fr.t;
        const float pdf = is_refl ? 1.f - T : T;
        const float scl = is_refl ? opt_scale(sc, b.spec[2], q.k) : opt_scale(sc, b.spec[3], q.k);
        if (scl == 0.f) return out;
        Mueller M;
        if (is_refl) M = mu_scale(mu_fresnel(fr.rs, fr.rp), scl);
        else { M = mu_scale(mu_fresnel(fr.ts, fr.tp), fr.Z * scl); if (!q.fwd) M = mu_scale(M, (fr.eta * fr.eta).re); }
        out.dpd = pd_disc(1.f); out.eta = fr.eta; out.M = mu_div(M, pdf); out.valid = true;
    } else if (b.type == WTGPU_BSDF_SURFACE_SPM) {      // surface_spm.cpp:79-170
        const float alpha = profile_alpha(sc, b, wi, wi, q.k);
        const bool has_spec = (q.lobes & 1u) && alpha > 0.f;
        const bool has_scat = (q.lobes & 2u) && alpha < 1.f;
        const C2 eta = spectrum_value(sc, b.spec[0], q.k) / spectrum_value(sc, b.spec[1], q.k);
        const bool has_tr = ior_has_transmission(eta);
        if (wi.z == 0.f || (!has_spec && !has_scat)) return out;
        float pdf = 1.f;
        bool is_spec = has_spec;
        if (has_spec && has_scat) { const float ps = alpha; is_spec = ps == 1.f || rnd(smp) < ps; pdf = is_spec ? ps : 1.f - ps; }
        float J = 1.f;
        const Fresnel fr = fresnel(eta, wi, nz);
        const float ptr = (fr.Ts + fr.Tp) / 2.f;
        bool is_refl = true;
        if (has_tr) { is_refl = rnd(smp) >= ptr; pdf *= is_refl ? 1.f - ptr : ptr; }
        if (!is_refl && !q.fwd) J = sqrf(fr.eta.re);
        const float scl = is_refl ? opt_scale(sc, b.spec[2], q.k) : opt_scale(sc, b.spec[3], q.k);
        if (scl == 0.f || (!is_refl && !has_tr)) return out;
        if (is_spec) {
            out.wo = is_refl ? reflect_z(wi) : fr.t;
            const Mueller F = mu_fresnel_dir(eta, is_refl, wi, nz);
            out.M = mu_div(mu_scale(F, alpha * J * scl), pdf);
            out.dpd = pd_disc(pdf);
        } else {
            const ProfSample ps = profile_sample(sc, b, wi, q.k, smp);
            const V3 h = wi + ps.wo;
            const V3 m = normalize(wi.z < 0.f ? -h : h);
            const Mueller F = mu_fresnel_dir(eta, is_refl, wi, m);
            out.wo = is_refl ? ps.wo : flip_wo(ps.wo, eta.re);
            pdf *= ps.pdf;
            out.M = mu_div(mu_scale(F, (1.f - alpha) * J * fabsf(out.wo.z) * ps.psd * scl), pdf);
            out.dpd = pd_dens(pdf);
        }
        out.eta = is_refl ? mkc(1.f, 0.f) : fr.eta;
        out.valid = true;
    }
    if (out.valid) {
        if (r.two_sided) out.wo = flipz(out.wo, wiz0);
        if (r.scale != 1.f) out.M = mu_scale(out.M, r.scale);
    }
    return out;
}

// ================================================================================================ emitters
WT_D V3 m3mul(const float* M, V3 v) { return mk3(M[0] * v.x + M[1] * v.y + M[2] * v.z, M[3] * v.x + M[4] * v.y + M[5] * v.z, M[6] * v.x + M[7] * v.y + M[8] * v.z); }

WT_D Sourcing emitter_sourcing(const wtgpu_emitter& e, float k) {    // point.hpp:74-88, spot.hpp:115-130, area.hpp:151-165, directional.hpp:118-128
    if (e.type == WTGPU_EMITTER_DIRECTIONAL) {
        const float l0 = e.tan_alpha > 0.f ? (sqrtf(0.25f) * sqrf(kEnvelope) / (k * e.tan_alpha)) * 0.001f : 0.f;
        const float l = sqrtf(sqrf(l0));
        float se = l * l, ta = e.tan_alpha;
        enlarge(se, ta, e.pse_scale);
        return source_extent(se, ta);
    }
    const float extent = (e.type != WTGPU_EMITTER_AREA && e.extent > 0.f) ? e.extent : 10.f * wavenum_to_wavelen(k);
    float se = extent * extent, ta = mub_tan_alpha(extent, k);
    enlarge(se, ta, e.pse_scale);
    if (e.type == WTGPU_EMITTER_SPOT) ta = fminf(ta, pm::tanf(e.falloff));
    return source_extent(se, ta);
}
WT_D float spot_falloff(const wtgpu_emitter& e, V3 ld) {           // spot.hpp:76-81
    const float ct = ld.z;
    if (ct <= pm::cosf(e.cutoff)) return 0.f;
    if (ct >= pm::cosf(e.falloff)) return 1.f;
    return (e.cutoff - pm::acosf(ct)) * (1.f / (e.cutoff - e.falloff));
}
WT_D Beam area_Le(const DScene& sc, const wtgpu_emitter& e, V3 o, V3 d, float k, const Surface& s) {  // area.hpp:104-117,170-180
    const float rad = e.scale * spectrum_f(sc, e.spectrum, k);
    return beam_make(true, o, d, rad * fmaxf(0.f, dot(d, s.geo.n)), k, emitter_sourcing(e, k));
}
WT_D uint32_t icdf_index(const float* cdf, uint32_t n, float v) {   // discrete_distribution_t::icdf (discrete_distribution.hpp:102-108)
    // lower_bound over n+1 entries
    uint32_t lo = 0, hi = n + 1;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (cdf[mid] < v) lo = mid + 1; else hi = mid; }
    int64_t idx = (int64_t)lo - 1;
    if (idx < 0) idx = 0; if (idx > (int64_t)n - 1) idx = (int64_t)n - 1;
    for (; idx < (int64_t)n - 1 && cdf[idx + 1] - cdf[idx] == 0.f; ++idx) {}
    return (uint32_t)idx;
}
struct PosSample { V3 p; float ppd; Surface s; };
WT_D PosSample sample_shape_position(const DScene& sc, int32_t shape, Sampler& smp) {   // src/scene/shape.cpp:70-89
    const wtgpu_shape sh = sc.shapes[shape];
    const V3 r = rnd3(smp);
    const uint32_t idx = icdf_index(sc.shape_tri_cdf + sh.cdf_first, sh.n_tris, r.z);
    const V2 bary = uniform_triangle(mk2(r.x, r.y));
    const uint32_t tuid = sc.shape_tri_tuid[sh.tri_first + idx];
    PosSample ps; ps.s = make_surface_at_bary(sc, tuid, bary); ps.p = ps.s.wp; ps.ppd = 1.f / sh.surface_area;
    return ps;
}
struct EmitterSample { Beam beam; Pd ppd, dpd; bool has_surface; Surface s; };
WT_DN EmitterSample emitter_sample(const DScene& sc, int32_t i, Sampler& smp, float k) {
    const wtgpu_emitter e = sc.emitters[i];
    EmitterSample r; r.has_surface = false;
    const V3 pos = mk3(e.pos);
    if (e.type == WTGPU_EMITTER_POINT) {            // point.cpp:28-41
        const V3 d = uniform_sphere(rnd2(smp));
        r.beam = beam_make(true, pos, d, spectrum_f(sc, e.spectrum, k), k, emitter_sourcing(e, k));
        beam_mul(r.beam, kFourPi);
        r.ppd = pd_disc(1.f); r.dpd = pd_dens(kInvFourPi);
    } else if (e.type == WTGPU_EMITTER_SPOT) {      // spot.cpp:29-46
        const float csa = kTwoPi * (1.f - pm::cosf(e.cutoff));
        const V3 lwo = uniform_cone(csa, rnd2(smp));
        const V3 wo = normalize(m3mul(e.rot, lwo));
        const float w = spot_falloff(e, lwo);
        const float dpd = 1.f / csa;
        r.beam = beam_make(true, pos, wo, spectrum_f(sc, e.spectrum, k), k, emitter_sourcing(e, k));
        beam_mul(r.beam, w); beam_div(r.beam, dpd);
        r.ppd = pd_disc(1.f); r.dpd = pd_dens(dpd);
    } else if (e.type == WTGPU_EMITTER_DIRECTIONAL) {   // directional.cpp:28-47
        const V3 dir = mk3(e.dir);
        const Frame fr = orthogonal_frame(dir);
        const V2 p = concentric_disk(rnd2(smp)) * e.world_radius;
        const V3 wp = mk3(e.world_centre) + to_world(fr, p);
        const float area = kPi * sqrf(e.world_radius);
        r.beam = beam_make(true, wp + e.far_dist * dir, -dir, spectrum_f(sc, e.spectrum, k), k, emitter_sourcing(e, k));
        beam_mul(r.beam, area);
        r.ppd = pd_dens(1.f / area); r.dpd = pd_disc(1.f);
    } else {                                        // area.cpp:52-80
        const PosSample ps = sample_shape_position(sc, e.shape, smp);
        V3 d = cosine_hemisphere(rnd2(smp));
        const float dn = d.z;
        d = to_world(ps.s.geo, d);
        const float dpd = cosine_hemisphere_pdf(dn), ppd = ps.ppd;
        float rp = 1.f / (dpd * ppd);
        if (dpd * ppd == 0.f) rp = 0.f;
        r.beam = area_Le(sc, e, ps.p, d, k, ps.s);
        beam_mul(r.beam, rp);
        r.ppd = pd_dens(ppd); r.dpd = pd_dens(dpd); r.has_surface = true; r.s = ps.s;
    }
    return r;
}
struct EmitterDirect { int32_t emitter; float emitter_pdf; Pd dpd; Beam beam; bool has_surface; V3 sp; uint32_t stuid; };
WT_DN EmitterDirect emitter_sample_direct(const DScene& sc, int32_t i, Sampler& smp, V3 wp, float k) {
    const wtgpu_emitter e = sc.emitters[i];
    EmitterDirect r; r.emitter = i; r.emitter_pdf = 0.f; r.has_surface = false; r.stuid = WTGPU_INVALID_IDX; r.sp = mk3(0.f, 0.f, 0.f);
    const V3 pos = mk3(e.pos);
    if (e.type == WTGPU_EMITTER_POINT || e.type == WTGPU_EMITTER_SPOT) {    // point.cpp:43-60, spot.cpp:48-67
        const V3 dl = wp - pos;
        const float rd2 = 1.f / length2(dl);
        const V3 d = dl * sqrtf(rd2);
        r.beam = beam_make(true, pos, d, spectrum_f(sc, e.spectrum, k), k, emitter_sourcing(e, k));
        if (e.type == WTGPU_EMITTER_SPOT) beam_mul(r.beam, spot_falloff(e, normalize(m3mul(e.inv_rot, d))));
        beam_mul(r.beam, rd2);
        r.dpd = pd_disc(1.f);
    } else if (e.type == WTGPU_EMITTER_DIRECTIONAL) {   // directional.cpp:49-72
        const V3 dir = mk3(e.dir);
        const Frame fr = orthogonal_frame(dir);
        const V3 wc = mk3(e.world_centre);
        const V3 pl = to_local(fr, wp - wc);
        const V2 p = mk2(pl.x, pl.y);
        const float scale = length2(p) <= sqrf(e.world_radius) ? 1.f : 0.f;
        r.beam = beam_make(true, (wc + to_world(fr, p)) + e.far_dist * dir, -dir, spectrum_f(sc, e.spectrum, k), k, emitter_sourcing(e, k));
        beam_mul(r.beam, scale);
        r.dpd = pd_disc(1.f);
    } else {                                            // area.cpp:82-105, 130-141
        const PosSample ps = sample_shape_position(sc, e.shape, smp);
        const V3 d = normalize(wp - ps.p);
        const float l2 = length2(wp - ps.p);
        const float dn = fmaxf(0.f, dot(d, ps.s.geo.n));
        const float dpd = ps.ppd * l2 * (dn > 0.f ? 1.f / dn : 0.f);
        r.beam = area_Le(sc, e, ps.p, d, k, ps.s);
        beam_mul(r.beam, dpd > 0.f ? 1.f / dpd : 0.f);
        r.dpd = pd_dens(dpd); r.has_surface = true; r.sp = ps.p; r.stuid = ps.s.tuid;
    }
    return r;
}
WT_D int32_t sample_emitter(const DScene& sc, Sampler& smp) { return (int32_t)icdf_index(sc.emitter_cdf, sc.n_emitters, rnd(smp)); }
WT_D float pdf_emitter(const DScene& sc, int32_t i) { return sc.emitter_cdf[i + 1] - sc.emitter_cdf[i]; }
// scene_t::sample_emitter_direct (scene/scene.hpp:128-141)
WT_D EmitterDirect scene_sample_emitter_direct(const DScene& sc, Sampler& smp, V3 wp, float k) {
    const int32_t e = sample_emitter(sc, smp);
    const float pd = pdf_emitter(sc, e);
    EmitterDirect r = emitter_sample_direct(sc, e, smp, wp, k);
    r.emitter_pdf = pd; beam_div(r.beam, pd);
    return r;
}
// area_t::Li (area.cpp:35-50)
WT_D Stokes emitter_Li(const DScene& sc, int32_t i, const Beam& Sbeam, const Surface& s) {
    const wtgpu_emitter e = sc.emitters[i];
    if (e.type != WTGPU_EMITTER_AREA) return stokes_zero();
    const float dn = dot(-Sbeam.env.d, s.geo.n);
    if (dn <= 0.f) return stokes_zero();
    Beam I = area_Le(sc, e, s.wp, -Sbeam.env.d, Sbeam.k, s);
    beam_div(I, dn);
    return integrate_beams(Sbeam, I);
}
// product-spectrum wavenumber sampling (discrete_distribution.hpp:258-272, binned_piecewise_linear_distribution.hpp:251-292)
struct KSample { float k; Pd wpd; };
WT_D KSample sample_wavenumber(const DScene& sc, int32_t i, Sampler& smp) {
    const wtgpu_kdist kd = sc.emitter_kdist[i];
    const float* data = sc.kdist_data + kd.first;
    const float v = rnd(smp);
    const uint32_t n = kd.n;
    KSample r;
    if (kd.type == WTGPU_KDIST_DISCRETE) {
        const uint32_t idx = icdf_index(data + 2 * n, n, v);
        r.k = data[idx]; r.wpd = pd_disc(data[n + idx] * kd.norm);
        return r;
    }
    const float* ys = data; const float* dcdf = data + n;
    uint32_t lo = 0, hi = n;    // upper_bound
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (!(v < dcdf[mid])) lo = mid + 1; else hi = mid; }
    int64_t ii = (int64_t)lo - 1; if (ii < 0) ii = 0; if (ii > (int64_t)n - 2) ii = (int64_t)n - 2;
    uint32_t idx = (uint32_t)ii;
    while (idx + 1 < n - 1 && v > dcdf[idx + 1]) ++idx;
    const float f = (v - dcdf[idx]) / (dcdf[idx + 1] - dcdf[idx]);
    const float a = ys[idx], b = ys[idx + 1];
    if (a == b) { r.k = ((float)idx + f) * kd.dk; r.wpd = pd_dens(a * kd.norm); return r; }      // sic (reference line 270)
    const float dd = sqrtf(mixf(sqrf(a), sqrf(b), f));
    const float t = clampf_((a - dd) / (a - b), 0.f, 1.f);
    r.k = mixf(kd.k0 + (float)idx * kd.dk, kd.k0 + (float)(idx + 1) * kd.dk, t);
    r.wpd = pd_dens(mixf(a, b, t) * kd.norm);
    return r;
}
WT_D float pdf_wavenumber(const DScene& sc, int32_t i, float k) {     // scene_sensor.hpp:63-70
    const wtgpu_kdist kd = sc.emitter_kdist[i];
    const float* data = sc.kdist_data + kd.first;
    const uint32_t n = kd.n;
    if (kd.type == WTGPU_KDIST_DISCRETE) {
        uint32_t lo = 0, hi = n;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (data[mid] < k) lo = mid + 1; else hi = mid; }
        if (lo == n || data[lo] != k) return 0.f;
        return data[2 * n + lo + 1] - data[2 * n + lo];
    }
    const float bin = (k - kd.k0) * (1.f / kd.dk);
    if (bin < 0.f || bin > (float)(n - 1)) return 0.f;
    const uint32_t ii = (uint32_t)bin;
    return mixf(data[ii], data[min(n - 1, ii + 1)], bin - floorf(bin)) * kd.norm;
}
WT_D float sum_spectral_pdf(const DScene& sc, float k) {             // scene_sensor.hpp:115-123
    float s = 0.f;
    for (uint32_t i = 0; i < sc.n_emitters; ++i) s += pdf_emitter(sc, (int32_t)i) * pdf_wavenumber(sc, (int32_t)i, k);
    return s;
}

// ================================================================================================ sensors
struct Element { uint32_t ex, ey; float ox, oy; };
WT_D void m4mul(const float* M, const float v[4], float o[4]) { _Pragma("unroll") for (int r = 0; r < 4; ++r) o[r] = M[4 * r] * v[0] + M[4 * r + 1] * v[1] + M[4 * r + 2] * v[2] + M[4 * r + 3] * v[3]; }
WT_D V3 persp_point_on_sensor(const wtgpu_sensor& s, V2 fp) { const float v[4] = { fp.x, fp.y, 1.f, 1.f }; float p[4]; m4mul(s.s2c, v, p); return mk3(p[0], p[1], p[2]) / p[3]; }
WT_D V2 persp_point_on_film(const wtgpu_sensor& s, V3 dir) { const V3 p = dir / fabsf(dir.z); const float v[4] = { p.x, p.y, 1.f, 1.f }; float q[4]; m4mul(s.c2s, v, q); return mk2(q[0], q[1]) / q[3]; }
WT_D V2 persp_extent(const wtgpu_sensor& s) {
    const V3 f0 = persp_point_on_sensor(s, mk2(0.f, 0.f));
    return mk2(length(persp_point_on_sensor(s, mk2((float)s.width, 0.f)) - f0), length(persp_point_on_sensor(s, mk2(0.f, (float)s.height)) - f0));
}
WT_D float persp_recp_sa(const wtgpu_sensor& s, V3 d) { const V2 e = persp_extent(s); return (e.x * e.y) / sqrf(0.01f) * (d.z * d.z * d.z); }
WT_D Sourcing persp_sourcing(const wtgpu_sensor& s, float k) {      // perspective.hpp:190-206
    const float ise = (persp_extent(s).x / (float)s.width) * .25f * kEnvelope;
    float se = ise * ise, ta = s.sourcing_tan_alpha;
    enlarge(se, ta, s.pse_scale);
    return source_extent(se, ta);
}
WT_D V2 vp_elem_extent(const wtgpu_sensor& s) { return mk2(s.extent[0] / (float)s.width, s.extent[1] / (float)s.height); }
WT_D Sourcing vp_sourcing(const wtgpu_sensor& s, float k) {         // virtual_plane_sensor.hpp:137-153
    const V2 ee = vp_elem_extent(s);
    Sourcing g; g.l = (ee.x + ee.y) / 2.f * .25f * kEnvelope;
    g.ta = s.requested_tan_alpha >= 0.f ? s.requested_tan_alpha : mub_tan_alpha(g.l, k);
    return g;
}
WT_D Beam vp_Se(const wtgpu_sensor& s, V3 o, V3 d, float k) {       // virtual_plane_sensor.hpp:165-183
    const float W = 1.f / kPi * (1.f / (s.extent[0] * s.extent[1]));
    return beam_make(false, o, d, W * fmaxf(0.f, dot(d, mk3(s.frame_n))), k, vp_sourcing(s, k));
}
struct SensorSample { Beam beam; Pd ppd, dpd; Element el; bool has_surface; };
WT_DN SensorSample sensor_sample(const DScene& sc, Sampler& smp, uint32_t ex, uint32_t ey, float k) {
    const wtgpu_sensor& s = sc.sensor;
    SensorSample r;
    if (s.type == WTGPU_SENSOR_PERSPECTIVE) {       // perspective.hpp:229-269
        const V3 centre = persp_point_on_sensor(s, mk2((float)ex, (float)ey) + mk2(.5f, .5f));
        const V2 off = rnd2(smp) - mk2(.5f, .5f);
        const V3 p00 = persp_point_on_sensor(s, mk2(0.f, 0.f));
        const V3 ddx = persp_point_on_sensor(s, mk2(1.f, 0.f)) - p00, ddy = persp_point_on_sensor(s, mk2(0.f, 1.f)) - p00;
        const V3 dl = normalize(centre + off.x * ddx + off.y * ddy);
        const V3 dir = normalize(m3mul(s.rot, dl));
        const float rdpd = persp_recp_sa(s, dl);
        r.beam = beam_make(false, mk3(s.pos), dir, 1.f / rdpd, k, persp_sourcing(s, k));
        beam_mul(r.beam, rdpd);
        r.ppd = pd_disc(1.f); r.dpd = pd_dens(1.f / rdpd); r.el.ex = ex; r.el.ey = ey; r.el.ox = off.x; r.el.oy = off.y; r.has_surface = false;
        return r;
    }
    // virtual_plane_sensor.cpp:101-132
    const V2 off = rnd2(smp) - mk2(.5f, .5f);
    const V2 ee = vp_elem_extent(s);
    const V2 local = mk2((float)((double)((float)ex + off.x) + .5), (float)((double)((float)ey + off.y) + .5)) * ee;
    Frame f; f.t = mk3(s.frame_t); f.b = mk3(s.frame_b); f.n = mk3(s.frame_n);
    const V3 p = mk3(s.origin) + local.x * f.t + local.y * f.b;
    const float rppd = s.extent[0] * s.extent[1];
    const V3 wo = cosine_hemisphere(rnd2(smp));
    const float dpd = cosine_hemisphere_pdf(wo.z);
    r.beam = vp_Se(s, p, to_world(f, wo), k);
    beam_mul(r.beam, rppd); beam_mul(r.beam, dpd > 0.f ? 1.f / dpd : 0.f);
    r.ppd = pd_dens(1.f / rppd); r.dpd = pd_dens(dpd); r.el.ex = ex; r.el.ey = ey; r.el.ox = off.x; r.el.oy = off.y; r.has_surface = true;
    return r;
}
struct SensorDirect { Beam beam; Pd dpd; Element el; };
WT_DN SensorDirect sensor_sample_direct(const DScene& sc, Sampler& smp, V3 wp, float k) {
    const wtgpu_sensor& s = sc.sensor;
    SensorDirect r;
    if (s.type == WTGPU_SENSOR_PERSPECTIVE) {       // perspective.hpp:274-314
        const V3 pos = mk3(s.pos);
        const V3 wdl = wp - pos;
        const float rd2 = 1.f / length2(wdl);
        const V3 wd = wdl * sqrtf(rd2);
        const V3 dl = normalize(m3mul(s.inv_rot, wd));
        const V2 fp = persp_point_on_film(s, dl);
        const float rsa = persp_recp_sa(s, dl);
        const bool inside = dl.z > 1.1920929e-7f && fp.x >= 0.f && fp.y >= 0.f && fp.x < (float)s.width && fp.y < (float)s.height;
        r.el.ex = inside ? (uint32_t)fp.x : 0u; r.el.ey = inside ? (uint32_t)fp.y : 0u;
        r.el.ox = (fp.x - floorf(fp.x)) - .5f; r.el.oy = (fp.y - floorf(fp.y)) - .5f;
        r.beam = beam_make(false, pos, wd, 1.f / rsa, k, persp_sourcing(s, k));
        beam_mul(r.beam, rd2); beam_mul(r.beam, inside ? 1.f : 0.f);
        r.dpd = pd_disc(1.f);
        return r;
    }
    // virtual_plane_sensor.cpp:134-176
    Frame f; f.t = mk3(s.frame_t); f.b = mk3(s.frame_b); f.n = mk3(s.frame_n);
    const V2 spl = rnd2(smp) * mk2(s.extent[0], s.extent[1]);
    const V3 sp = mk3(s.origin) + spl.x * f.t + spl.y * f.b;
    const V2 ee = vp_elem_extent(s);
    const V2 efp = mk2(spl.x / ee.x, spl.y / ee.y);
    r.el.ex = (uint32_t)efp.x; r.el.ey = (uint32_t)efp.y;
    r.el.ox = efp.x - (float)r.el.ex - .5f; r.el.oy = efp.y - (float)r.el.ey - .5f;
    const V3 wdl = wp - sp;
    const float d2 = length2(wdl);
    const V3 wd = wdl / sqrtf(d2);
    const float lz = dot(wd, f.n);
    const float rdn = lz > 0.f ? 1.f / lz : 0.f;
    const float dpd = (1.f / (s.extent[0] * s.extent[1])) * d2 * rdn;
    r.beam = vp_Se(s, sp, wd, k);
    beam_mul(r.beam, dpd > 0.f ? 1.f / dpd : 0.f); beam_mul(r.beam, rdn);
    r.dpd = pd_dens(dpd);
    return r;
}
// virtual_plane_sensor_t::Si (virtual_plane_sensor.cpp:65-99)
WT_DN bool sensor_Si(const DScene& sc, const Beam& beam, Range range, Beam& se, Element& el) {
    const wtgpu_sensor& s = sc.sensor;
    if (s.type != WTGPU_SENSOR_VIRTUAL_PLANE) return false;
    Frame f; f.t = mk3(s.frame_t); f.b = mk3(s.frame_b); f.n = mk3(s.frame_n);
    const float dn = dot(-beam.env.d, f.n);
    if (dn <= 0.f) return false;
    const V3 o = mk3(s.origin);
    const V3 a = o, b = o + s.extent[0] * f.t, c = o + s.extent[1] * f.b, d = (o + s.extent[0] * f.t) + s.extent[1] * f.b;
    const RayTri i1 = intersect_ray_tri(beam.env.o, beam.env.d, a, b, c, range);
    const RayTri i2 = intersect_ray_tri(beam.env.o, beam.env.d, c, b, d, range);
    if (!i1.hit && !i2.hit) return false;
    const V3 p = beam.env.o + beam.env.d * (i1.hit ? i1.dist : i2.dist);
    se = vp_Se(s, p, -beam.env.d, beam.k);
    beam_div(se, dn);
    const V3 sp = p - o;
    const V2 ee = vp_elem_extent(s);
    const V2 efp = mk2(dot(sp, f.t) * (1.f / ee.x), dot(sp, f.b) * (1.f / ee.y));
    el.ex = (uint32_t)efp.x; el.ey = (uint32_t)efp.y; el.ox = efp.x - (float)el.ex - .5f; el.oy = efp.y - (float)el.ey - .5f;
    return true;
}

// ================================================================================================ film
WT_D float erf_lut(const DScene& sc, float x) {                    // math/erf_lut.hpp:20-55
    const float sg = signf_(x);
    x = fabsf(x) * (1023.f / 3.5f);
    const float fr = x - floorf(x);
    const uint32_t i0 = (uint32_t)x, i1 = i0 + 1u;
    return (i1 >= 1024u || i1 == 0u) ? sg : sg * mixf(__ldg(sc.erf_lut + i0), __ldg(sc.erf_lut + i1), fr);
}
WT_D float rf_integrate(const DScene& sc, float mn, float mx) {    // gaussian1d.hpp:100-106
    const float sigma = sc.sensor.rfilter_stddev;
    if (sigma == 0.f) return (mn <= 0.f && 0.f <= mx) ? 1.f : 0.f;
    const float n = kInvSqrtTwo * (1.f / sigma);
    return (erf_lut(sc, mx * n) - erf_lut(sc, mn * n)) / 2.f;
}
// film_t::splat / splat_direct (sensor/film/film.hpp:214-288, 308-340) -> atomic accumulation in f32.
// Returns the number of taps written.
WT_DN uint32_t film_splat(const DScene& sc, float* film_block, float* film_light, bool direct, const Element& e, float I, float k) {
    const wtgpu_sensor& s = sc.sensor;
    const int r = (int)s.rf_radius;
    const int Wd = 2 * r + 1;
    float tx[9], ty[9];
    for (int x = -r; x <= r; ++x) { tx[x + r] = rf_integrate(sc, x + e.ox - .5f, x + e.ox + .5f); ty[x + r] = rf_integrate(sc, x + e.oy - .5f, x + e.oy + .5f); }
    float tw = 0.f;
    for (int x = 0; x < Wd; ++x) for (int y = 0; y < Wd; ++y) tw += fmaxf(0.f, 1.f * tx[x] * ty[y]);
    const float rtw = tw > 0.f ? 1.f / tw : 0.f;
    uint32_t taps = 0;
    for (uint32_t c = 0; c < s.channels; ++c) {
        float val = I * spectrum_f(sc, s.response[c], k);
        if (direct) { if (val <= 0.f || !isfinite(val)) continue; }
        else val = (val >= 0.f && isfinite(val)) ? val : 0.f;
        for (int dx = -r; dx <= r; ++dx) for (int dy = -r; dy <= r; ++dy) {
            const float w = fmaxf(0.f, 1.f * tx[dx + r] * ty[dy + r]) * rtw;
            const long px = (long)e.ex + dx, py = (long)e.ey + dy;
            if (px < 0 || py < 0 || px >= (long)s.width || py >= (long)s.height) continue;
            const size_t pi = ((size_t)py * s.width + (size_t)px) * s.channels + c;
            if (direct) atomicAdd(film_light + pi, w * val);
            else { atomicAdd(film_block + 2 * pi, w * val); atomicAdd(film_block + 2 * pi + 1, w); }
            ++taps;
        }
    }
    return taps;
}

} // namespace wt
