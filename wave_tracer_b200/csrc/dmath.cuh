// dmath.cuh -- device arithmetic substrate of the B200 wave_tracer hot path (product code, sm_100a).
//
// Restates, for the GPU, the f32 arithmetic of the reference's math headers so that hit/miss decisions agree with
// the CPU path (SURVEY.md 7 "hard part 8"): compensated products where the reference uses eft:: (include/wt/math/eft/eft.hpp),
// fma chains where it uses m::fma (include/wt/math/vecmath.hpp:21-66), and nothing contracted implicitly
// (this translation unit is compiled with --fmad=false).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include "pmath.h"

#define WT_D __device__ __forceinline__
// "large" device functions.  Force-inlined by default: out-of-line calls pass Beam/Surface/Mueller structs through local memory
// (ncu r01: 41% of k_shade's stall samples sat on STL); inlining lets the compiler scalarise them.  -DWT_DN=... overrides for A/B runs.
#ifndef WT_DN
#define WT_DN __device__ __forceinline__
#endif

namespace wt {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kFourPi = 12.5663706143591729539f;
constexpr float kPi2 = 1.57079632679489661923f;
constexpr float kPi4 = 0.78539816339744830962f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kInvTwoPi = 0.15915494309189533577f;
constexpr float kInvFourPi = 0.07957747154594766788f;
constexpr float kInvSqrtTwo = 0.70710678118654752440f;
constexpr float kInvSqrtTwoPi = 0.39894228040143267794f;
constexpr float kSqrtPi2 = 1.25331413731550025121f;
#define WT_INF CUDART_INF_F

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct C2 { float re, im; };     // complex

WT_D V2 mk2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
WT_D V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
WT_D V3 mk3(const float* p) { return mk3(p[0], p[1], p[2]); }
WT_D V2 operator+(V2 a, V2 b) { return mk2(a.x + b.x, a.y + b.y); }
WT_D V2 operator-(V2 a, V2 b) { return mk2(a.x - b.x, a.y - b.y); }
WT_D V2 operator*(V2 a, float s) { return mk2(a.x * s, a.y * s); }
WT_D V2 operator*(float s, V2 a) { return mk2(a.x * s, a.y * s); }
WT_D V2 operator*(V2 a, V2 b) { return mk2(a.x * b.x, a.y * b.y); }
WT_D V2 operator/(V2 a, float s) { return mk2(a.x / s, a.y / s); }
WT_D V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
WT_D V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
WT_D V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
WT_D V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
WT_D V3 operator*(float s, V3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
WT_D V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
WT_D V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
WT_D bool veq(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
WT_D V3 vabs(V3 a) { return mk3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
WT_D float max3f(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }
WT_D float min3f(float a, float b, float c) { return fminf(a, fminf(b, c)); }
WT_D float vmaxel(V3 a) { return max3f(a.x, a.y, a.z); }
WT_D bool vfinite(V3 a) { return isfinite(a.x) && isfinite(a.y) && isfinite(a.z); }
WT_D float sqrf(float x) { return x * x; }
WT_D float signf_(float t) { return (t > 0.f ? 1.f : 0.f) - (t < 0.f ? 1.f : 0.f); }
WT_D float mixf(float a, float b, float x) { if (x == 0.f) return a; if (x == 1.f) return b; return a * (1.f - x) + b * x; }
WT_D float clampf_(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
// AVX max/min semantics of the reference's 8-wide slab test (second operand wins on NaN)
WT_D float vmaxps(float a, float b) { return a > b ? a : b; }
WT_D float vminps(float a, float b) { return a < b ? a : b; }

// ---- error-free transformations (eft.hpp:36-53, 117-125, 156-162)
WT_D float diff_prod(float a, float b, float c, float d) { const float cd = c * d; const float r = fmaf(a, b, -cd); return r + fmaf(-c, d, cd); }
WT_D float sum_prod(float a, float b, float c, float d) { return diff_prod(a, b, -c, d); }
WT_D float two_prod(float& err, float a, float b) { const float p = a * b; err = fmaf(a, b, -p); return p; }
WT_D float two_sum(float& err, float a, float b) { const float s = a + b; const float e1 = s - a; const float e2 = s - e1; err = (b - e1) + (a - e2); return s; }

WT_D float dot(V2 a, V2 b) { return fmaf(a.y, b.y, a.x * b.x); }
WT_D float dot(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
WT_D V3 cross(V3 x, V3 y) { return mk3(diff_prod(x.y, y.z, x.z, y.y), diff_prod(x.z, y.x, x.x, y.z), diff_prod(x.x, y.y, x.y, y.x)); }
WT_D float length2(V2 a) { return dot(a, a); }
WT_D float length2(V3 a) { return dot(a, a); }
WT_D float length(V2 a) { return sqrtf(dot(a, a)); }
WT_D float length(V3 a) { return sqrtf(dot(a, a)); }
WT_D V2 normalize(V2 a) { return a / length(a); }
WT_D V3 normalize(V3 a) { return a / length(a); }
WT_D float eft_dot3(V3 a, V3 b) {        // eft.hpp:170-183
    float d = 0.f, err = 0.f, e1, e2;
    float t = two_prod(e1, a.x, b.x); d = two_sum(e2, d, t); err = err + e1 + e2;
    t = two_prod(e1, a.y, b.y); d = two_sum(e2, d, t); err = err + e1 + e2;
    t = two_prod(e1, a.z, b.z); d = two_sum(e2, d, t); err = err + e1 + e2;
    return d + err;
}

// ---- complex helpers (std::complex<float> semantics as compiled by g++ 13 / libgcc: naive product, binary64 quotient)
WT_D C2 mkc(float re, float im) { C2 c; c.re = re; c.im = im; return c; }
WT_D C2 operator+(C2 a, C2 b) { return mkc(a.re + b.re, a.im + b.im); }
WT_D C2 operator-(C2 a, C2 b) { return mkc(a.re - b.re, a.im - b.im); }
WT_D C2 operator-(C2 a) { return mkc(-a.re, -a.im); }
WT_D C2 operator*(C2 a, C2 b) { return mkc(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
WT_D C2 operator*(C2 a, float s) { return mkc(a.re * s, a.im * s); }
WT_D C2 operator*(float s, C2 a) { return mkc(a.re * s, a.im * s); }
WT_D C2 cconj(C2 a) { return mkc(a.re, -a.im); }
WT_D float cnorm(C2 a) { return a.re * a.re + a.im * a.im; }
WT_D float cabsf_(C2 a) { return pm::hypotf(a.re, a.im); }
WT_D C2 operator/(C2 a, C2 b) {
    // std::complex<float> division = libgcc's __divsc3, which since GCC 12 evaluates in binary64 (libgcc2.c, L_divsc3: "float is handled by
    // using double arithmetic"): products exact, one rounding per sum and quotient, then to float.  Bit-identical to the oracle's c_t / c_t.
    const double c = (double)b.re, d = (double)b.im;
    const double den = c * c + d * d;
    return mkc((float)(((double)a.re * c + (double)a.im * d) / den), (float)(((double)a.im * c - (double)a.re * d) / den));
}
WT_D C2 cexpi(float phase) { float s, c; pm::sincosf(phase, &s, &c); return mkc(c, s); }
WT_D C2 csqrt_(C2 z) {
    if (z.re == 0.f && z.im == 0.f) return mkc(0.f, 0.f);
    const float r = pm::hypotf(z.re, z.im);
    const float t = sqrtf(0.5f * (r + fabsf(z.re)));
    if (z.re >= 0.f) return mkc(t, z.im / (2.f * t));
    return mkc(fabsf(z.im) / (2.f * t), z.im >= 0.f ? t : -t);
}

// ---- range (include/wt/math/range.hpp)
struct Range { float mn, mx; };
WT_D Range mkr(float a, float b) { Range r; r.mn = a; r.mx = b; return r; }
WT_D bool rcontains(Range r, float p) { return (p < r.mx && r.mn < p) || p == r.mn || p == r.mx; }
WT_D bool rempty(Range r) { if (r.mn == r.mx && !isfinite(r.mn)) return true; return r.mn > r.mx; }
WT_D Range rand_(Range a, Range b) { return mkr(fmaxf(a.mn, b.mn), fminf(a.mx, b.mx)); }

// ---- frame (include/wt/math/frame.hpp)
struct Frame { V3 t, b, n; };
WT_D V3 to_local(const Frame& f, V3 v) { return mk3(dot(v, f.t), dot(v, f.b), dot(v, f.n)); }
WT_D V2 to_local2(const Frame& f, V2 v) { return mk2(dot(v, mk2(f.t.x, f.t.y)), dot(v, mk2(f.b.x, f.b.y))); }
WT_D V3 to_world(const Frame& f, V3 v) { return f.t * v.x + f.b * v.y + f.n * v.z; }
WT_D V3 to_world(const Frame& f, V2 v) { return f.t * v.x + f.b * v.y; }
WT_D float handness(const Frame& f) { return dot(cross(f.n, f.t), f.b) > 0.f ? 1.f : -1.f; }
WT_D Frame orthogonal_frame(V3 n) {
    V3 b;
    if (fabsf(n.x) > fabsf(n.y)) { const float x = 1.f / sqrtf(sqrf(n.x) + sqrf(n.z)); b = mk3(x * n.z, 0.f, -x * n.x); }
    else { const float x = 1.f / sqrtf(sqrf(n.y) + sqrf(n.z)); b = mk3(0.f, x * n.z, -x * n.y); }
    Frame f; f.t = cross(b, n); f.b = b; f.n = n; return f;
}
WT_D Frame shading_frame(V3 n, V3 dpdu) {
    if (dpdu.x == 0.f && dpdu.y == 0.f && dpdu.z == 0.f) return orthogonal_frame(n);
    const V3 t = normalize(dpdu - n * dot(n, dpdu));
    const V3 b = normalize(cross(n, t));
    Frame f; f.t = cross(b, n); f.b = b; f.n = n; return f;
}

// ---- 2x2 column-major helpers, rotation, QR/SVD (rotation.hpp:66-77, linalg.hpp:24-135)
struct M2 { float c0x, c0y, c1x, c1y; };      // columns
WT_D V2 m2mul(const M2& A, V2 v) { return mk2(A.c0x * v.x + A.c1x * v.y, A.c0y * v.x + A.c1y * v.y); }
WT_D M2 m2mm(const M2& A, const M2& B) {
    M2 R;
    R.c0x = A.c0x * B.c0x + A.c1x * B.c0y; R.c0y = A.c0y * B.c0x + A.c1y * B.c0y;
    R.c1x = A.c0x * B.c1x + A.c1x * B.c1y; R.c1y = A.c0y * B.c1x + A.c1y * B.c1y;
    return R;
}
WT_D M2 rotation2(V2 from, V2 to) {
    const float X = sum_prod(from.x, to.x, from.y, to.y);
    M2 R; R.c0x = X; R.c0y = diff_prod(from.x, to.y, to.x, from.y); R.c1x = diff_prod(to.x, from.y, from.x, to.y); R.c1y = X; return R;
}
struct SVD2 { float Ucos, Usin, s1, s2; };
WT_D SVD2 svd2(const M2& A) {
    float a = A.c0x, b = A.c1x, c = A.c0y, d = A.c1y;
    float x, y, z;
    if (c == 0.f) { x = a; y = b; z = d; }
    else {
        const float mm = fmaxf(fabsf(c), fabsf(d));
        const float rm = 1.f / mm;
        c *= rm; d *= rm;
        const float r = sqrtf(c * c + d * d);
        const float l = 1.f / r;
        x = diff_prod(a, d, b, c) * l;
        y = sum_prod(a, c, b, d) * l;
        z = mm * r;
    }
    SVD2 o;
    const float n = fmaxf(fabsf(x), fabsf(y));
    if (n == 0.f) { o.Ucos = 1.f; o.Usin = 0.f; o.s1 = A.c0x; o.s2 = A.c1y; return o; }
    const float numer = (z - x) * (z + x) + sqrf(y);
    const float tt = numer != 0.f ? numer / (n * x * y) : 0.f;
    const float t = 2.f * (tt >= 0.f ? 1.f : -1.f) / (fabsf(tt) + sqrtf(sqrf(tt) + 4.f));
    const float c1 = 1.f / sqrtf(1.f + sqrf(t));
    const float s1 = c1 * t;
    const float usa = diff_prod(c1, x, s1, y);
    const float usb = sum_prod(s1, x, c1, y);
    const float usc = -s1 * z;
    const float usd = c1 * z;
    float sigma1 = sqrtf(sqrf(usa) + sqrf(usc));
    float sigma2 = sqrtf(sqrf(usb) + sqrf(usd));
    float dmax = fmaxf(sigma1, sigma2);
    const float usmax1 = sigma2 > sigma1 ? usd : usa;
    const float usmax2 = sigma2 > sigma1 ? usb : -usc;
    const float sg = x * z > 0.f ? 1.f : -1.f;
    dmax *= sigma2 > sigma1 ? sg : 1.f;
    sigma2 *= sg;
    const float r = 1.f / dmax;
    o.Ucos = dmax != 0.f ? usmax1 * r : 1.f; o.Usin = dmax != 0.f ? usmax2 * r : 0.f; o.s1 = sigma1; o.s2 = sigma2;
    return o;
}

// ---- ray + elliptic cone (include/wt/math/shapes/{ray,elliptic_cone}.hpp)
struct Ray { V3 o, d, invd; };
WT_D Ray mkray(V3 o, V3 d) { Ray r; r.o = o; r.d = d; r.invd = mk3(1.f / d.x, 1.f / d.y, 1.f / d.z); return r; }
WT_D V3 propagate(const Ray& r, float t) { return r.o + r.d * t; }

struct Cone {
    V3 o, d, x;             // origin, mean direction, tangent (major axis)
    float x0, ta, e, ooe;   // initial major half-axis, tan(alpha), major/minor, minor/major
};
WT_D float cone_zapex(const Cone& c) { return (c.x0 != 0.f || c.ta != 0.f) ? -c.x0 / c.ta : -WT_INF; }
WT_D bool cone_is_ray(const Cone& c) { return c.ta == 0.f && c.x0 == 0.f; }
WT_D Frame cone_frame(const Cone& c) { Frame f; f.t = c.x; f.b = cross(c.d, c.x); f.n = c.d; return f; }
WT_D V2 cone_axes(const Cone& c, float z) { const float r = c.ta * z + c.x0; return mk2(r * 1.f, r * c.ooe); }
WT_D Cone mkcone(V3 o, V3 d, V3 x, float x0, float ta, float ooe, float e) { Cone c; c.o = o; c.d = d; c.x = x; c.x0 = x0; c.ta = ta; c.e = e; c.ooe = ooe; return c; }
WT_D Cone cone_iso(V3 o, V3 d, float ta, float x0) { return mkcone(o, d, orthogonal_frame(d).t, x0, ta, 1.f, 1.f); }
WT_D Cone cone_ecc(V3 o, V3 d, V3 x, float ta, float ecc, float x0) { const float ooe = sqrtf(fmaxf(0.f, 1.f - sqrf(ecc))); return mkcone(o, d, x, x0, ta, ooe, 1.f / ooe); }
WT_D bool cone_contains_local(const Cone& c, V3 p, Range r) {
    return rcontains(r, p.z) && cone_zapex(c) <= p.z && sqrf(p.x) + sqrf(c.e * p.y) <= sqrf(p.z * c.ta + c.x0);
}
WT_D bool cone_contains_local_w(const Cone& c, V3 p, Range r) {       // wide (fma) variant, elliptic_cone.hpp:170-185
    const float ztx = fmaf(p.z, c.ta, c.x0);
    return cone_zapex(c) <= p.z && (r.mn <= p.z && r.mx >= p.z) && (sqrf(p.x) + sqrf(p.y * c.e)) <= sqrf(ztx);
}
WT_D bool cone_contains(const Cone& c, V3 p) { return cone_contains_local(c, to_local(cone_frame(c), p - c.o), mkr(0.f, WT_INF)); }

// ---- primitive tests
struct RayTri { bool hit; float dist; float bx, by; };
// scalar Moeller-Trumbore (intersect/ray.hpp:147-179)
WT_D RayTri intersect_ray_tri(V3 ro, V3 rd, V3 a, V3 b, V3 c, Range range) {
    RayTri out; out.hit = false; out.dist = WT_INF; out.bx = out.by = -1.f;
    const V3 ray = ro - a, e1 = b - a, e2 = c - a;
    const V3 crs = cross(rd, e2);
    float det = dot(e1, crs);
    if (det == 0.f) return out;
    const float sdet = det >= 0.f ? 1.f : -1.f;
    det *= sdet;
    const V3 q = cross(ray, e1);
    const float qe2 = sdet * dot(q, e2);
    const float bx = sdet * dot(ray, crs), by = sdet * dot(rd, q);
    if (bx >= 0.f && by >= 0.f && bx + by <= det && rcontains(mkr(det * range.mn, det * range.mx), qe2)) {
        const float rdet = 1.f / det;
        out.hit = true; out.dist = qe2 * rdet;
        const float bux = bx * rdet, buy = by * rdet;
        out.bx = 1.f - (bux + buy); out.by = bux;
    }
    return out;
}
// one lane of the 8-wide variant (intersect/ray.hpp:192-236) -- what ray traversal uses
WT_D float intersect_ray_tri_w(V3 ro, V3 rd, V3 a, V3 b, V3 c, Range range, float& baryx, float& baryy) {
    const V3 ray = ro - a, e1 = b - a, e2 = c - a;
    const V3 crs = cross(rd, e2);
    const float det = dot(e1, crs);
    const float rdet = 1.f / det;
    const V3 q = cross(ray, e1);
    const float z = dot(q, e2) * rdet;
    baryy = dot(ray, crs) * rdet;
    const float baryz = dot(rd, q) * rdet;
    baryx = 1.f - (baryy + baryz);
    const bool valid = det != 0.f && baryx >= 0.f && baryy >= 0.f && baryz >= 0.f && (range.mn <= z && range.mx >= z);
    return valid ? z : -WT_INF;
}
WT_D bool test_ray_tri_w(V3 ro, V3 rd, V3 a, V3 b, V3 c, Range range) {  // intersect/ray.hpp:77-113
    const V3 ray = ro - a, e1 = b - a, e2 = c - a;
    const V3 crs = cross(rd, e2);
    const float det = dot(e1, crs);
    const float rdet = 1.f / det;
    const V3 q = cross(ray, e1);
    const float betax = dot(ray, crs), betay = dot(rd, q);
    const float z = dot(q, e2) * rdet;
    return det != 0.f && (betax * rdet) >= 0.f && (betay * rdet) >= 0.f && ((betax + betay) * rdet) <= 1.f && z >= range.mn && z <= range.mx;
}

WT_D bool intersect_edge_plane(V3 p0, V3 p1, V3 pp, V3 n, V3& out) {   // intersect/misc.hpp:163-179
    const float d0 = dot(pp - p0, n), d1 = dot(pp - p1, n);
    const V3 E = p1 - p0;
    const float EdN = dot(E, n);
    if (signf_(d0) == signf_(d1) || EdN == 0.f) return false;
    const float d = d0 / EdN;
    if (d >= 0.f && 1.f >= d) { out = p0 + d * E; return true; }
    return false;
}
WT_D bool intersect_line_plane_z(V3 p0, V3 p1, float z, float& t) {     // intersect/ray.hpp:28-41 with n=(0,0,1)
    const V3 n = mk3(0.f, 0.f, 1.f);
    const float dn = dot(p1 - p0, n);
    if (dn == 0.f) return false;
    t = dot(mk3(0.f, 0.f, z) - p0, n) / dn;
    return true;
}
WT_D int intersect_edge_ellipse_points(V2 point0, V2 point1, float rx, float ry) {  // intersect/misc.hpp:77-127 (points only)
    const V2 rs = mk2(1.f / rx, 1.f / ry);
    const V2 p0 = point0 * rs, p1 = point1 * rs;
    const V2 d = p1 - p0;
    const float a = dot(d, d), b = 2.f * dot(p0, d), c = dot(p0, p0) - 1.f;
    const float det2 = b * b - 4.f * a * c;
    if (det2 <= 0.f || a == 0.f) return 0;
    const float ra = 1.f / a, det = sqrtf(det2);
    float t1 = .5f * (-b - signf_(b) * det) * ra;
    float t2 = t1 == 0.f ? -b * ra : c * ra / t1;
    if (t1 > t2) { const float t = t1; t1 = t2; t2 = t; }
    const bool u1 = t1 >= 0.f && 1.f >= t1, u2 = t2 >= 0.f && 1.f >= t2;
    return (u1 ? 1 : 0) + (u2 ? 1 : 0);
}
WT_D void intersect_edge_ellipsoid(V3 p0w, V3 p1w, V3 centre, V3 x, V3 y, V3 axes, float& t1, float& t2) {  // misc.hpp:40-72
    t1 = t2 = 0.f;
    const V3 z = cross(x, y);
    p0w = p0w - centre; p1w = p1w - centre;
    const V3 q0 = mk3(dot(p0w, x) / axes.x, dot(p0w, y) / axes.y, dot(p0w, z) / axes.z);
    const V3 q1 = mk3(dot(p1w, x) / axes.x, dot(p1w, y) / axes.y, dot(p1w, z) / axes.z);
    const V3 d = q1 - q0;
    const float a = dot(d, d), b = dot(q0, d) * 2.f, c = dot(q0, q0) - 1.f;
    const float det2 = b * b - 4.f * a * c;
    if (det2 <= 0.f || a == 0.f) return;
    const float ra = 1.f / a, det = sqrtf(det2);
    t1 = .5f * (-b - signf_(b) * det) * ra;
    t2 = t1 == 0.f ? -b * ra : c * ra / t1;
    if (t1 > t2) { const float t = t1; t1 = t2; t2 = t; }
}
WT_D bool point_in_triangle(V3 p, V3 a, V3 b, V3 c) {                    // math/util.hpp:88-107
    const V3 v0 = b - a, v1 = c - a, u = p - a;
    const float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(u, v0), d21 = dot(u, v1);
    const float d = diff_prod(d00, d11, d01, d01);
    const float sg = d > 0.f ? 1.f : -1.f;
    const float alpha = diff_prod(d11, d20, d01, d21), beta = diff_prod(d00, d21, d01, d20);
    return sg * alpha >= 0.f && sg * beta >= 0.f && sg * (alpha + beta) <= sg * d;
}
WT_D float cone_intersection_tolerance(V3 origin, V3 a, V3 b, V3 c) {   // cone_intersection_tolerance.hpp:23-41
    const float c0 = 4e-7f, c1 = 1e-6f, c2 = 1e-6f;
    const V3 mn = mk3(min3f(a.x, b.x, c.x), min3f(a.y, b.y, c.y), min3f(a.z, b.z, c.z));
    const V3 mx = mk3(max3f(a.x, b.x, c.x), max3f(a.y, b.y, c.y), max3f(a.z, b.z, c.z));
    const float ext = 2.f * fmaxf(vmaxel(vabs(mn)), vmaxel(vabs(mx)));
    const V3 obj = (c0 + c2) * vabs(origin) + mk3(c1 * ext, c1 * ext, c1 * ext);
    const V3 wrd = (c1 + c2) * vabs(origin);
    return vmaxel(obj + wrd);
}

// cone-edge in the cone's local frame (intersect/cone.hpp:38-128, in_local, edge semantics, clip planes on).
// Returns the closest point (p0 of the reference's result) when an intersection exists.
WT_D bool intersect_cone_edge_local(const Cone& cone, V3 p0, V3 p1, Range range, V3& out_p0) {
    V3 lp0 = p0, lp1 = p1;
    const bool p0closer = lp1.z > lp0.z;
    if (!p0closer) { const V3 t = lp0; lp0 = lp1; lp1 = t; }
    const V3 p = lp0, l = lp1 - lp0;
    const float x0 = cone.x0, ta = cone.ta, e = cone.e;
    const float cs = p.z * ta + x0;
    const float epy = e * p.y, ely = e * l.y, lzta = l.z * ta;
    const float c = sqrf(p.x) + diff_prod(epy, epy, cs, cs);
    const float b = 2.f * eft_dot3(mk3(p.x, epy, -lzta), mk3(l.x, ely, cs));
    const float a = sqrf(l.x) + diff_prod(ely, ely, lzta, lzta);
    const float D = b * b - 4.f * a * c;
    if (D < 0.f) return false;
    const float sD = sqrtf(D);
    float t1 = b >= 0.f ? (-b - sD) / (2.f * a) : (-b + sD) / (2.f * a);
    float t2 = (-b / a) - t1;
    const float zapex = cone_zapex(cone);
    if (p.z + t1 * l.z <= zapex) t1 = WT_INF;
    if (p.z + t2 * l.z < zapex) t2 = WT_INF;
    if (t2 < t1) { const float t = t1; t1 = t2; t2 = t; }
    float z1 = t1 < WT_INF ? p.z + t1 * l.z : -WT_INF;
    float z2 = t2 < WT_INF ? p.z + t2 * l.z : WT_INF;
    if (z1 > range.mx || z2 < range.mn || (!isfinite(z1) && !isfinite(z2))) return false;
    if (range.mn > zapex && z1 < range.mn) { float tm; if (intersect_line_plane_z(p, p + l, range.mn, tm)) { t1 = tm; z1 = range.mn; } }
    if (z2 > range.mx) { float tm; if (intersect_line_plane_z(p, p + l, range.mx, tm)) { t2 = tm; z2 = range.mx; } }
    const V3 base = p0closer ? p0 : p1;
    const V3 dir = p0closer ? p1 - p0 : p0 - p1;
    const bool has1 = t1 >= 0.f && 1.f >= t1, has2 = t2 >= 0.f && 1.f >= t2;
    if (!has1 && !has2) return false;
    out_p0 = has1 ? base + t1 * dir : base + t2 * dir;
    return true;
}

struct ConePlane { Range range; V3 nearp, farp; };
// cone-plane (intersect/cone.hpp:170-258)
WT_D ConePlane intersect_cone_plane(const Cone& cone, V3 n, float d, Range range, bool in_local) {
    const Frame frame = cone_frame(cone);
    if (!in_local) { d -= dot(cone.o, n); n = to_local(frame, n); }
    const float x0 = cone.x0, e = cone.ooe;
    const float vd2 = sqrf(n.x) + sqrf(e * n.y);
    const V2 v = vd2 > 0.f ? mk2(n.x, e * n.y) / sqrtf(vd2) : mk2(0.f, 0.f);
    const V2 u = v * mk2(1.f, e);
    const float nu = dot(n, mk3(u.x, u.y, 0.f));
    const float zapex = cone_zapex(cone);
    float z01 = (d - x0 * nu) / (n.z + cone.ta * nu);
    float z02 = (d + x0 * nu) / (n.z - cone.ta * nu);
    const bool h1 = z01 >= zapex && !isnan(z01), h2 = z02 >= zapex && !isnan(z02);
    if (!h1) z01 = WT_INF;
    if (!h2) z02 = WT_INF;
    const float s1 = z01 * cone.ta + x0, s2 = z02 * cone.ta + x0;
    V3 p1 = h1 ? mk3(s1 * u.x, s1 * u.y, z01) : mk3(WT_INF, WT_INF, WT_INF);
    V3 p2 = h2 ? mk3(s2 * (-u.x), s2 * (-u.y), z02) : mk3(WT_INF, WT_INF, WT_INF);
    if (z01 > z02) { const float t = z01; z01 = z02; z02 = t; const V3 tp = p1; p1 = p2; p2 = tp; }
    ConePlane out;
    Range rng = mkr(z01, z02);
    if ((!h1 && !h2) || rempty(rand_(rng, range))) { out.range = mkr(WT_INF, -WT_INF); out.nearp = out.farp = mk3(0.f, 0.f, 0.f); return out; }
    if (isfinite(rng.mn)) {
        if (rng.mn < range.mn) {
            const float z = range.mn; float xx, yy;
            if (fabsf(n.y) > fabsf(n.x)) { yy = (d - n.z * z) / n.y; xx = n.x != 0.f ? (d - n.z * z - n.y * yy) / n.x : 0.f; }
            else { xx = (d - n.z * z) / n.x; yy = n.y != 0.f ? (d - n.z * z - n.x * xx) / n.y : 0.f; }
            const float s = xx * v.x + yy * v.y;
            p1 = mk3(s * v.x, s * v.y, z); rng.mn = range.mn;
        }
        if (!in_local) p1 = cone.o + to_world(frame, p1);
    }
    const bool has_inf = h1 != h2;
    if (isfinite(rng.mx) || has_inf) {
        if (rng.mx > range.mx) {
            const float z = range.mx; float xx, yy;
            if (fabsf(n.y) > fabsf(n.x)) { yy = (d - n.z * z) / n.y; xx = n.x != 0.f ? (d - n.z * z - n.y * yy) / n.x : 0.f; }
            else { xx = (d - n.z * z) / n.x; yy = n.y != 0.f ? (d - n.z * z - n.x * xx) / n.y : 0.f; }
            const float s = xx * v.x + yy * v.y;
            p2 = mk3(s * v.x, s * v.y, z); rng.mx = range.mx;
        }
        if (!in_local) p2 = cone.o + to_world(frame, p2);
    }
    out.range = rng; out.nearp = p1; out.farp = p2;
    return out;
}

// The rejections intersect_cone_tri (below) takes before its containment / plane / edge stages, as a predicate of their own: false = that
// function returns +inf for these arguments (same arithmetic, same comparisons).  The team traversal (ctrav.cuh) runs it over a whole batch of
// triangles and sends only the survivors through the full test, so the expensive stages run with full warps.
WT_D bool cone_tri_maybe(const Cone& cone, const Frame& frame, V3 a, V3 b, V3 c, Range range) {
    if (cone_is_ray(cone)) return true;
    const V3 o = cone.o;
    const V3 v0 = to_local(frame, a - o), v1 = to_local(frame, b - o), v2 = to_local(frame, c - o);
    const float closest_z = min3f(v0.z, v1.z, v2.z), farthest_z = max3f(v0.z, v1.z, v2.z);
    if (farthest_z < range.mn || closest_z > range.mx) return false;
#ifndef WT_NO_CONE_QUICK_REJECT
    if (cone.ta >= 0.f) {
        const float r = fmaf(fminf(farthest_z, range.mx), cone.ta, cone.x0);
        if (r > 0.f && r < WT_INF) {
            const float u0 = v0.x, u1 = v1.x, u2 = v2.x, w0 = v0.y * cone.e, w1 = v1.y * cone.e, w2 = v2.y * cone.e;
            const float ext = fmaxf(fmaxf(max3f(fabsf(u0), fabsf(u1), fabsf(u2)), max3f(fabsf(w0), fabsf(w1), fabsf(w2))), fmaxf(fabsf(closest_z), fabsf(farthest_z)));
            const float bb = r + (1e-3f * r + 1e-5f * ext), bd = 1.41421356f * r + (2e-3f * r + 2e-5f * ext);
            const float p0 = u0 + w0, p1 = u1 + w1, p2 = u2 + w2, q0 = u0 - w0, q1 = u1 - w1, q2 = u2 - w2;
            if (min3f(u0, u1, u2) > bb || max3f(u0, u1, u2) < -bb || min3f(w0, w1, w2) > bb || max3f(w0, w1, w2) < -bb ||
                min3f(p0, p1, p2) > bd || max3f(p0, p1, p2) < -bd || min3f(q0, q1, q2) > bd || max3f(q0, q1, q2) < -bd) return false;
        }
    }
#endif
    return true;
}

// cone-triangle closest distance (intersect/cone.hpp:550-626); returns +inf when there is no intersection
WT_DN float intersect_cone_tri(const Cone& cone, const Frame& frame, V3 a, V3 b, V3 c, V3 n, Range range) {
    if (cone_is_ray(cone)) { const RayTri r = intersect_ray_tri(cone.o, cone.d, a, b, c, range); return r.hit ? r.dist : WT_INF; }
    const V3 o = cone.o;
    V3 vs[3];
    vs[0] = to_local(frame, a - o); vs[1] = to_local(frame, b - o); vs[2] = to_local(frame, c - o);
    const V3 ln = to_local(frame, n);
    // (the z-range rejection is taken before the containment tests: same result, less work for the triangles a leaf step rejects)
    const float closest_z = min3f(vs[0].z, vs[1].z, vs[2].z), farthest_z = max3f(vs[0].z, vs[1].z, vs[2].z);
    if (farthest_z < range.mn || closest_z > range.mx) return WT_INF;
#ifndef WT_NO_CONE_QUICK_REJECT
    // Separating-axis rejection (ours; the reference goes straight to the plane and edge tests).  Inside the z slab the cone's cross-section is
    // contained in |x| <= r, |e y| <= r, |x +- e y| <= sqrt(2) r with r = tan(alpha) z1 + x0 at the far end z1 of the slab; a triangle whose three
    // vertices lie beyond one of these lines -- by a slack of 1e-3 r + 1e-5 x its own extent, orders of magnitude above the rounding of
    // the tests it skips -- cannot touch the cone, and every later stage would report "no intersection".
    if (cone.ta >= 0.f) {
        const float r = fmaf(fminf(farthest_z, range.mx), cone.ta, cone.x0);
        if (r > 0.f && r < WT_INF) {
            const float u0 = vs[0].x, u1 = vs[1].x, u2 = vs[2].x, v0 = vs[0].y * cone.e, v1 = vs[1].y * cone.e, v2 = vs[2].y * cone.e;
            const float ext = fmaxf(fmaxf(max3f(fabsf(u0), fabsf(u1), fabsf(u2)), max3f(fabsf(v0), fabsf(v1), fabsf(v2))), fmaxf(fabsf(closest_z), fabsf(farthest_z)));
            const float b = r + (1e-3f * r + 1e-5f * ext), bd = 1.41421356f * r + (2e-3f * r + 2e-5f * ext);
            const float p0 = u0 + v0, p1 = u1 + v1, p2 = u2 + v2, q0 = u0 - v0, q1 = u1 - v1, q2 = u2 - v2;
            if (min3f(u0, u1, u2) > b || max3f(u0, u1, u2) < -b || min3f(v0, v1, v2) > b || max3f(v0, v1, v2) < -b ||
                min3f(p0, p1, p2) > bd || max3f(p0, p1, p2) < -bd || min3f(q0, q1, q2) > bd || max3f(q0, q1, q2) < -bd) return WT_INF;
        }
    }
#endif
    bool in[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) in[i] = cone_contains_local_w(cone, vs[i], range);
#pragma unroll
    for (int i = 0; i < 3; ++i) if (in[i] && vs[i].z == closest_z) return closest_z;
    const ConePlane icp = intersect_cone_plane(cone, ln, dot(vs[0], ln), range, true);
    if (!rempty(icp.range) && point_in_triangle(icp.nearp, vs[0], vs[1], vs[2])) return icp.range.mn;
    bool hasp = false; float pz = 0.f;
    // one copy of the cone-edge test instead of three (the vertices rotate through registers; same edges, same order): the unrolled form
    // was ~700 SASS instructions of kernels whose dominant stall is instruction fetch (etoile-like k_gtraverse 111 -> 91 ms/step, profiles/r01s3_phases.txt session O)
    V3 ea = vs[0], eb = vs[1], ec = vs[2]; bool ia = in[0], ib = in[1], ic = in[2];
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
        if (!(ia && ib) && !(ea.z > range.mx && eb.z > range.mx) && !(ea.z < range.mn && eb.z < range.mn)) {
            V3 cp;
            if (intersect_cone_edge_local(cone, ea, eb, range, cp) && (!hasp || pz > cp.z)) { pz = cp.z; hasp = true; }
        }
        const V3 tv = ea; ea = eb; eb = ec; ec = tv;
        const bool tb = ia; ia = ib; ib = ic; ic = tb;
    }
    return hasp ? pz : WT_INF;
}

// elliptic_cone_t::cone_through_ellipse (src/math/elliptic_cone.cpp:19-88)
WT_DN Cone cone_through_ellipse(V3 x, V3 y, V3 n, V3 ro, V3 rd, float tan_alpha, float* sid) {
    const bool xz = x.x == 0.f && x.y == 0.f && x.z == 0.f, yz = y.x == 0.f && y.y == 0.f && y.z == 0.f;
    if (xz && yz) { if (sid) *sid = 0.f; return mkcone(ro, rd, orthogonal_frame(rd).t, 0.f, tan_alpha, 1.f, 1.f); }
    const Frame of = orthogonal_frame(rd);
    const V3 xl = to_local(of, x), yl = to_local(of, y);
    M2 A; A.c0x = xl.x; A.c0y = xl.y; A.c1x = yl.x; A.c1y = yl.y;
    const SVD2 s = svd2(A);
    V2 X = mk2(s.Ucos, -s.Usin);
    float lX = fabsf(s.s1), lY = fabsf(s.s2);
    if (lX < lY) { const float t = lX; lX = lY; lY = t; X = mk2(s.Usin, s.Ucos); }
    const float e = lY > 0.f ? sqrtf(lX / lY) : 1.f;
    const V3 wx = to_world(of, X);
    const Cone cone = mkcone(ro, rd, wx, lX, tan_alpha, 1.f / e, e);
    if (sid) {
        const ConePlane cp = intersect_cone_plane(cone, n, dot(n, ro), mkr(0.f, WT_INF), false);
        *sid = rempty(cp.range) ? 0.f : cp.range.mx;
    }
    return cone;
}
// elliptic_cone_t::cone_through_ellipsoid (src/math/elliptic_cone.cpp:90-145)
WT_DN Cone cone_through_ellipsoid(V3 axes, const Frame& axes_frame, V3 ro, V3 rd, float tan_alpha) {
    const V3 wol = to_local(axes_frame, rd);
    const Frame frame = orthogonal_frame(wol);
    const V3 nn = normalize(axes * wol);
    const Frame fc = orthogonal_frame(nn);
    const V3 t1 = axes * fc.t, t2 = axes * fc.b;
    const V2 c0 = to_local2(frame, mk2(t1.x, t1.y)), c1 = to_local2(frame, mk2(t2.x, t2.y));
    M2 A; A.c0x = c0.x; A.c0y = c0.y; A.c1x = c1.x; A.c1y = c1.y;
    if (A.c0x * A.c1y == A.c1x * A.c0y) return mkcone(ro, rd, orthogonal_frame(rd).t, 0.f, tan_alpha, 1.f, 1.f);
    const SVD2 s = svd2(A);
    V2 X = mk2(s.Ucos, -s.Usin);
    float lX = fabsf(s.s1), lY = fabsf(s.s2);
    if (lX < lY) { const float t = lX; lX = lY; lY = t; X = mk2(s.Usin, s.Ucos); }
    const float e = lY > 0.f ? sqrtf(lX / lY) : 1.f;
    const V3 X3 = normalize(to_world(frame, X));
    return mkcone(ro, rd, to_world(axes_frame, X3), lX, tan_alpha, 1.f / e, e);
}

} // namespace wt
