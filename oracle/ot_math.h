// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// ot_math.h: CPU restatement (plain C++17, f32, no mp-units/glm) of the arithmetic substrate of
// wave_tracer's hot path.  Every function cites the reference file:line it follows.
// PARITY UNPINNED: the reference has no tests/golden vectors and cannot be compiled here
// (SURVEY.md 8c); the oracle is pinned only by analytic KATs (tests/test_oracle_kats.py).
//
// Arithmetic contract shared with the CUDA path: IEEE f32, no implicit FMA contraction (oracle built with
// -ffp-contract=off, CUDA with --fmad=false); fma is used exactly where the reference calls m::fma / eft::*.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <algorithm>
#include <complex>
#include <optional>

namespace ot {

using f_t = float;
using c_t = std::complex<float>;
static constexpr f_t inf = std::numeric_limits<f_t>::infinity();
static constexpr f_t pi = 3.14159265358979323846f;
static constexpr f_t two_pi = 6.28318530717958647692f;
static constexpr f_t four_pi = 12.5663706143591729539f;
static constexpr f_t pi_2 = 1.57079632679489661923f;
static constexpr f_t pi_4 = 0.78539816339744830962f;
static constexpr f_t inv_pi = 0.31830988618379067154f;
static constexpr f_t inv_two_pi = 0.15915494309189533577f;
static constexpr f_t inv_four_pi = 0.07957747154594766788f;
static constexpr f_t sqrt_two = 1.41421356237309504880f;
static constexpr f_t inv_sqrt_two = 0.70710678118654752440f;
static constexpr f_t inv_sqrt_two_pi = 0.39894228040143267794f;
static constexpr f_t sqrt_pi_2 = 1.25331413731550025121f;   // sqrt(pi/2)

// ---- elementary functions.  OT_PORTABLE_LIBM=1 (liboracle.so, what the GPU is compared with): the portable binary64-internal functions of
// wave_tracer_b200/csrc/pmath.h, which return the same bits on the host and on the device -- so that device-vs-oracle differences are the
// code's, not the math library's (the path is ill-conditioned in libm ulps: a phase is k*L ~ 1e5..1e8 rad).  OT_PORTABLE_LIBM=0
// (liboracle_glibc.so): the host libm through std::, exactly as the reference calls it (m::sin ... -> std::sin, include/wt/math/common.hpp);
// this build is the one pinned bit for bit against the reference's own translation units (tests/test_oracle_kats.py) and is compared with
// the portable build in tests/test_pmath.py (films agree to the f32 conditioning of the path, functions to <= 1 ulp).
#ifndef OT_PORTABLE_LIBM
#define OT_PORTABLE_LIBM 1
#endif
} // namespace ot
#if OT_PORTABLE_LIBM
#include "../wave_tracer_b200/csrc/pmath.h"
#endif
namespace ot {
namespace lm {
#if OT_PORTABLE_LIBM
inline f_t sin(f_t x) { return pm::sinf(x); }
inline f_t cos(f_t x) { return pm::cosf(x); }
inline f_t tan(f_t x) { return pm::tanf(x); }
inline f_t exp(f_t x) { return pm::expf(x); }
inline f_t log(f_t x) { return pm::logf(x); }
inline f_t pow(f_t x, f_t y) { return pm::powf(x, y); }
inline f_t atan2(f_t y, f_t x) { return pm::atan2f(y, x); }
inline f_t acos(f_t x) { return pm::acosf(x); }
inline f_t hypot(f_t a, f_t b) { return pm::hypotf(a, b); }
inline c_t expi(f_t x) { f_t s, c; pm::sincosf(x, &s, &c); return { c, s }; }                       // std::exp(c_t{0, x})
inline c_t polar(f_t rho, f_t th) { f_t s, c; pm::sincosf(th, &s, &c); return { rho * c, rho * s }; }   // std::polar
inline f_t cabs(c_t z) { return pm::hypotf(z.real(), z.imag()); }
inline c_t csqrt(c_t z) {       // principal square root, the formula of the device's csqrt_ (glibc csqrtf without its over/underflow scaling)
    if (z.real() == 0 && z.imag() == 0) return { 0, 0 };
    const f_t r = pm::hypotf(z.real(), z.imag());
    const f_t t = std::sqrt(0.5f * (r + std::fabs(z.real())));
    if (z.real() >= 0) return { t, z.imag() / (2 * t) };
    return { std::fabs(z.imag()) / (2 * t), z.imag() >= 0 ? t : -t };
}
#else
inline f_t sin(f_t x) { return std::sin(x); }
inline f_t cos(f_t x) { return std::cos(x); }
inline f_t tan(f_t x) { return std::tan(x); }
inline f_t exp(f_t x) { return std::exp(x); }
inline f_t log(f_t x) { return std::log(x); }
inline f_t pow(f_t x, f_t y) { return std::pow(x, y); }
inline f_t atan2(f_t y, f_t x) { return std::atan2(y, x); }
inline f_t acos(f_t x) { return std::acos(x); }
inline f_t hypot(f_t a, f_t b) { return std::hypot(a, b); }
inline c_t expi(f_t x) { return std::exp(c_t{ 0, x }); }
inline c_t polar(f_t rho, f_t th) { return std::polar<f_t>(rho, th); }
inline f_t cabs(c_t z) { return std::abs(z); }
inline c_t csqrt(c_t z) { return std::sqrt(z); }
#endif
}

inline f_t sqr(f_t x) { return x * x; }
inline f_t sign(f_t t) { return f_t(t > 0) - f_t(t < 0); }             // glm::sign
inline f_t mix(f_t a, f_t b, f_t x) {                                   // math/common.hpp:258-264
    if (x == 0) return a;
    if (x == 1) return b;
    return a * (1 - x) + b * x;
}
inline f_t clampf(f_t v, f_t lo, f_t hi) { return std::min(std::max(v, lo), hi); }
inline f_t max3(f_t a, f_t b, f_t c) { return std::max(a, std::max(b, c)); }
inline f_t min3(f_t a, f_t b, f_t c) { return std::min(a, std::min(b, c)); }

// ---- error-free transformations: include/wt/math/eft/eft.hpp
inline f_t diff_prod(f_t a, f_t b, f_t c, f_t d) {                      // eft.hpp:117-125
    const f_t cd = c * d;
    const f_t ret = std::fma(a, b, -cd);
    return ret + std::fma(-c, d, cd);
}
inline f_t sum_prod(f_t a, f_t b, f_t c, f_t d) { return diff_prod(a, b, -c, d); }   // eft.hpp:156-162
inline f_t two_prod(f_t& err, f_t a, f_t b) { const f_t p = a * b; err = std::fma(a, b, -p); return p; }  // eft.hpp:36-42
inline f_t two_sum(f_t& err, f_t a, f_t b) {                            // eft.hpp:44-53
    const f_t s = a + b; const f_t e1 = s - a; const f_t e2 = s - e1; err = (b - e1) + (a - e2); return s;
}

struct v2 { f_t x, y; };
struct v3 { f_t x, y, z; };
inline v2 operator+(v2 a, v2 b) { return { a.x + b.x, a.y + b.y }; }
inline v2 operator-(v2 a, v2 b) { return { a.x - b.x, a.y - b.y }; }
inline v2 operator*(v2 a, f_t s) { return { a.x * s, a.y * s }; }
inline v2 operator*(f_t s, v2 a) { return { a.x * s, a.y * s }; }
inline v2 operator*(v2 a, v2 b) { return { a.x * b.x, a.y * b.y }; }
inline v2 operator/(v2 a, f_t s) { return { a.x / s, a.y / s }; }
inline v3 operator+(v3 a, v3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline v3 operator-(v3 a, v3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline v3 operator-(v3 a) { return { -a.x, -a.y, -a.z }; }
inline v3 operator*(v3 a, f_t s) { return { a.x * s, a.y * s, a.z * s }; }
inline v3 operator*(f_t s, v3 a) { return { a.x * s, a.y * s, a.z * s }; }
inline v3 operator*(v3 a, v3 b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
inline v3 operator/(v3 a, f_t s) { return { a.x / s, a.y / s, a.z / s }; }
inline bool operator==(v3 a, v3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(v3 a, v3 b) { return !(a == b); }
inline v3 absv(v3 a) { return { std::fabs(a.x), std::fabs(a.y), std::fabs(a.z) }; }
inline f_t max_element(v3 a) { return max3(a.x, a.y, a.z); }
inline bool isfinite3(v3 a) { return std::isfinite(a.x) && std::isfinite(a.y) && std::isfinite(a.z); }

// include/wt/math/vecmath.hpp:21-66: dot is an fma chain, cross uses compensated diff_prod
inline f_t dot(v2 a, v2 b) { return std::fma(a.y, b.y, a.x * b.x); }
inline f_t dot(v3 a, v3 b) { return std::fma(a.z, b.z, std::fma(a.y, b.y, a.x * b.x)); }
inline v3 cross(v3 x, v3 y) {
    return { diff_prod(x.y, y.z, x.z, y.y), diff_prod(x.z, y.x, x.x, y.z), diff_prod(x.x, y.y, x.y, y.x) };
}
inline f_t length2(v2 a) { return dot(a, a); }
inline f_t length2(v3 a) { return dot(a, a); }
inline f_t length(v2 a) { return std::sqrt(length2(a)); }
inline f_t length(v3 a) { return std::sqrt(length2(a)); }
inline v2 normalize(v2 a) { return a / length(a); }
inline v3 normalize(v3 a) { return a / length(a); }
// eft::dot (eft.hpp:170-183)
inline f_t eft_dot(v3 a, v3 b) {
    f_t d = 0, err = 0, e1, e2;
    const f_t av[3] = { a.x, a.y, a.z }, bv[3] = { b.x, b.y, b.z };
    for (int i = 0; i < 3; ++i) { const f_t t = two_prod(e1, av[i], bv[i]); d = two_sum(e2, d, t); err = err + e1 + e2; }
    return d + err;
}
inline f_t eft_dot(v2 a, v2 b) {
    f_t d = 0, err = 0, e1, e2;
    const f_t av[2] = { a.x, a.y }, bv[2] = { b.x, b.y };
    for (int i = 0; i < 2; ++i) { const f_t t = two_prod(e1, av[i], bv[i]); d = two_sum(e2, d, t); err = err + e1 + e2; }
    return d + err;
}

// ---- range: include/wt/math/range.hpp
struct range_t {
    f_t min, max;
    bool contains(f_t p) const { return (p < max && min < p) || p == min || p == max; }
    bool empty() const { if (min == max && !std::isfinite(min)) return true; return min > max; }
    f_t length() const { return max - min; }
    f_t centre() const { return (max + min) / 2.f; }
    range_t grow(f_t e) const { return { min - e, max + e }; }
    bool overlaps(const range_t& r) const { return min <= r.max && r.min <= max; }
    static range_t positive() { return { 0, inf }; }
    static range_t all() { return { -inf, inf }; }
    static range_t null() { return { inf, -inf }; }
};
inline range_t operator&(range_t a, range_t b) { return { std::max(a.min, b.min), std::min(a.max, b.max) }; }
inline range_t operator|(range_t a, range_t b) { return { std::min(a.min, b.min), std::max(a.max, b.max) }; }

// ---- frame: include/wt/math/frame.hpp:18-190
struct frame_t {
    v3 t, b, n;
    v3 to_local(v3 v) const { return { dot(v, t), dot(v, b), dot(v, n) }; }
    v2 to_local2(v2 v) const { return { dot(v, v2{ t.x, t.y }), dot(v, v2{ b.x, b.y }) }; }   // frame.hpp:22-27
    v3 to_world(v3 v) const { return t * v.x + b * v.y + n * v.z; }
    v3 to_world(v2 v) const { return t * v.x + b * v.y; }
    f_t handness() const { const f_t h = dot(cross(n, t), b); return h > 0 ? 1.f : -1.f; }
    static frame_t canonical() { return { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } }; }
    static frame_t build_orthogonal_frame(v3 n) {                       // frame.hpp:158-174
        v3 b;
        if (std::fabs(n.x) > std::fabs(n.y)) { const f_t x = 1 / std::sqrt(sqr(n.x) + sqr(n.z)); b = { x * n.z, 0, -x * n.x }; }
        else { const f_t x = 1 / std::sqrt(sqr(n.y) + sqr(n.z)); b = { 0, x * n.z, -x * n.y }; }
        return { cross(b, n), b, n };
    }
    static frame_t build_shading_frame(v3 n, v3 dpdu) {                 // frame.hpp:139-153
        if (dpdu.x == 0 && dpdu.y == 0 && dpdu.z == 0) return build_orthogonal_frame(n);
        const v3 t = normalize(dpdu - n * dot(n, dpdu));
        const v3 b = normalize(cross(n, t));
        return { cross(b, n), b, n };
    }
};

// ---- column-major 2x2 (glm::mat2 semantics: m[col][row])
struct mat2 {
    f_t m[2][2];
    mat2() : m{ { 0, 0 }, { 0, 0 } } {}
    mat2(f_t c0r0, f_t c0r1, f_t c1r0, f_t c1r1) : m{ { c0r0, c0r1 }, { c1r0, c1r1 } } {}
    mat2(v2 c0, v2 c1) : m{ { c0.x, c0.y }, { c1.x, c1.y } } {}
};
inline v2 operator*(const mat2& A, v2 v) { return { A.m[0][0] * v.x + A.m[1][0] * v.y, A.m[0][1] * v.x + A.m[1][1] * v.y }; }
inline mat2 operator*(const mat2& A, const mat2& B) {                   // glm: result[c] = A * B[c]
    mat2 R;
    for (int c = 0; c < 2; ++c) for (int r = 0; r < 2; ++r) R.m[c][r] = A.m[0][r] * B.m[c][0] + A.m[1][r] * B.m[c][1];
    return R;
}

// include/wt/math/rotation.hpp:66-77
inline mat2 rotation_matrix(v2 from, v2 to) {
    const f_t xa = from.x, xb = to.x, ya = from.y, yb = to.y;
    const f_t X = sum_prod(xa, xb, ya, yb);
    return mat2{ X, diff_prod(xa, yb, xb, ya), diff_prod(xb, ya, xa, yb), X };
}

// ---- 2x2 QR / SVD: include/wt/math/linalg.hpp:24-135
struct QR_ret_t { f_t Qcos, Qsin; f_t x, y, z; };   // R = [[x,y],[0,z]]
inline QR_ret_t QR(const mat2& A) {
    f_t a = A.m[0][0], b = A.m[1][0], c = A.m[0][1], d = A.m[1][1];
    f_t x, y, z, Qc, Qs;
    if (c == 0) { x = a; y = b; z = d; Qc = 1; Qs = 0; }
    else {
        const f_t mm = std::max(std::fabs(c), std::fabs(d));
        const f_t recp_m = 1 / mm;
        c *= recp_m; d *= recp_m;
        const f_t r = std::sqrt(c * c + d * d);
        const f_t l = 1.f / r;
        x = diff_prod(a, d, b, c) * l;
        y = sum_prod(a, c, b, d) * l;
        z = mm * r;
        Qs = -c * l; Qc = d * l;
    }
    return { Qc, Qs, x, y, z };
}
struct SVD_ret_t { f_t Ucos, Usin, Vcos, Vsin, sigma1, sigma2; };
inline SVD_ret_t SVD(const mat2& A) {
    const QR_ret_t qr = QR(A);
    const f_t x = qr.x, y = qr.y, z = qr.z;
    f_t c2 = qr.Qcos, s2 = qr.Qsin;
    const f_t n = std::max(std::fabs(x), std::fabs(y));
    if (n == 0) return { 1, 0, c2, s2, A.m[0][0], A.m[1][1] };
    const f_t numer = (z - x) * (z + x) + sqr(y);
    const f_t tt = numer != 0 ? numer / (n * x * y) : 0;
    f_t t = 2 * f_t(tt >= 0 ? 1 : -1) / (std::fabs(tt) + std::sqrt(sqr(tt) + 4));
    const f_t c1 = 1 / std::sqrt(1 + sqr(t));
    const f_t s1 = c1 * t;
    const f_t usa = diff_prod(c1, x, s1, y);
    const f_t usb = sum_prod(s1, x, c1, y);
    const f_t usc = -s1 * z;
    const f_t usd = c1 * z;
    t = sum_prod(c1, c2, s1, s2);
    s2 = diff_prod(c2, s1, c1, s2);
    c2 = t;
    f_t sigma1 = std::sqrt(sqr(usa) + sqr(usc));
    f_t sigma2 = std::sqrt(sqr(usb) + sqr(usd));
    f_t dmax = std::max(sigma1, sigma2);
    const f_t usmax1 = sigma2 > sigma1 ? usd : usa;
    const f_t usmax2 = sigma2 > sigma1 ? usb : -usc;
    const f_t signsigma1 = f_t(x * z > 0 ? 1 : -1);
    dmax *= sigma2 > sigma1 ? signsigma1 : 1;
    sigma2 *= signsigma1;
    const f_t r = 1 / dmax;
    return { dmax != 0 ? usmax1 * r : 1, dmax != 0 ? usmax2 * r : 0, c2, s2, sigma1, sigma2 };
}

// ---- ray: include/wt/math/shapes/ray.hpp
struct ray_t {
    v3 o, d, invd;
    ray_t() = default;
    ray_t(v3 o_, v3 d_) : o(o_), d(d_), invd{ 1.f / d_.x, 1.f / d_.y, 1.f / d_.z } {}
    v3 propagate(f_t t) const { return o + d * t; }
};

struct aabb_t {
    v3 min, max;
    static aabb_t from_points(v3 a, v3 b, v3 c) {
        return { { min3(a.x, b.x, c.x), min3(a.y, b.y, c.y), min3(a.z, b.z, c.z) },
                 { max3(a.x, b.x, c.x), max3(a.y, b.y, c.y), max3(a.z, b.z, c.z) } };
    }
};

// ---- primitive ray tests: include/wt/math/intersect/ray.hpp
struct intersect_ray_tri_ret_t { f_t dist = inf; v2 bary{ -1, -1 }; };

// scalar Moeller-Trumbore, ray.hpp:147-179
inline std::optional<intersect_ray_tri_ret_t> intersect_ray_tri(const ray_t& r, v3 a, v3 b, v3 c, range_t range = range_t::positive()) {
    const v3 ray = r.o - a;
    const v3 e1 = b - a, e2 = c - a;
    const v3 crs = cross(r.d, e2);
    f_t det = dot(e1, crs);
    if (det == 0) return std::nullopt;
    const f_t sdet = det >= 0 ? 1.f : -1.f;
    det *= sdet;
    const v3 q = cross(ray, e1);
    const f_t qe2 = sdet * dot(q, e2);
    const f_t bx = sdet * dot(ray, crs);
    const f_t by = sdet * dot(r.d, q);
    const range_t dr{ det * range.min, det * range.max };
    if (bx >= 0 && by >= 0 && bx + by <= det && dr.contains(qe2)) {
        const f_t recp_det = 1.f / det;
        const f_t dist = qe2 * recp_det;
        const f_t bux = bx * recp_det, buy = by * recp_det;
        return intersect_ray_tri_ret_t{ dist, { 1 - (bux + buy), bux } };
    }
    return std::nullopt;
}
// the 8-wide variant evaluated for one lane, ray.hpp:192-236 (this is what bvh8w ray traversal uses)
struct ray_tri_w_ret_t { f_t result; f_t baryx, baryy; };
inline ray_tri_w_ret_t intersect_ray_tri_w(v3 ro, v3 rd, v3 a, v3 b, v3 c, range_t range) {
    const v3 ray = ro - a;
    const v3 e1 = b - a, e2 = c - a;
    const v3 crs = cross(rd, e2);
    const f_t det = dot(e1, crs);
    const f_t recp_det = 1.f / det;
    bool valid = det != 0;
    const v3 q = cross(ray, e1);
    const f_t qe2 = dot(q, e2);
    const f_t betax = dot(ray, crs);
    const f_t betay = dot(rd, q);
    const f_t z = qe2 * recp_det;
    const f_t baryy = betax * recp_det;
    const f_t baryz = betay * recp_det;
    const f_t baryx = 1.f - (baryy + baryz);
    valid = valid && baryx >= 0 && baryy >= 0 && baryz >= 0 && (range.min <= z && range.max >= z);
    return { valid ? z : -inf, baryx, baryy };
}
// ray.hpp:77-113 (wide test_ray_tri, one lane)
inline bool test_ray_tri_w(v3 ro, v3 rd, v3 a, v3 b, v3 c, range_t range) {
    const v3 ray = ro - a;
    const v3 e1 = b - a, e2 = c - a;
    const v3 crs = cross(rd, e2);
    const f_t det = dot(e1, crs);
    const f_t recp_det = 1.f / det;
    const bool valid = det != 0;
    const v3 q = cross(ray, e1);
    const f_t qe2 = dot(q, e2);
    const f_t betax = dot(ray, crs), betay = dot(rd, q);
    const f_t bxy = betax + betay;
    const f_t z = qe2 * recp_det;
    return valid && (betax * recp_det) >= 0 && (betay * recp_det) >= 0 && (bxy * recp_det) <= 1.f && z >= range.min && z <= range.max;
}
// scalar test_ray_tri, ray.hpp:56-76
inline bool test_ray_tri(const ray_t& r, v3 a, v3 b, v3 c, range_t range = range_t::positive(), f_t tol = 0) {
    const v3 ray = r.o - a;
    const v3 e1 = b - a, e2 = c - a;
    const v3 crs = cross(r.d, e2);
    const f_t det = dot(e1, crs);
    if (det == 0) return false;
    const f_t recp_det = 1.f / det;
    const v3 q = cross(ray, e1);
    const f_t qe2 = dot(q, e2);
    const f_t bx = dot(ray, crs), by = dot(r.d, q);
    return bx * recp_det >= -tol && by * recp_det >= -tol && (bx + by) * recp_det <= 1 + tol && range.contains(qe2 * recp_det);
}

// include/wt/math/intersect/misc.hpp:163-179
inline std::optional<v3> intersect_edge_plane(v3 p0, v3 p1, v3 pp, v3 n) {
    const f_t d0 = dot(pp - p0, n);
    const f_t d1 = dot(pp - p1, n);
    const v3 E = p1 - p0;
    const f_t EdN = dot(E, n);
    if (sign(d0) == sign(d1) || EdN == 0) return std::nullopt;
    const f_t d = d0 / EdN;
    if (d >= 0 && 1 >= d) return p0 + d * E;
    return std::nullopt;
}
// include/wt/math/intersect/ray.hpp:28-41 (intersect_line_plane)
inline std::optional<f_t> intersect_line_plane(v3 p0, v3 p1, v3 pp, v3 n) {
    const f_t dn = dot(p1 - p0, n);
    if (dn == 0) return std::nullopt;
    return dot(pp - p0, n) / dn;
}

// include/wt/math/intersect/misc.hpp:77-127
struct intersect_edge_circle_ret_t { int points = 0; v2 u1{}, u2{}; f_t t1 = 0, t2 = 0; };
inline intersect_edge_circle_ret_t intersect_edge_ellipse(v2 point0, v2 point1, f_t rx, f_t ry, bool line = false) {
    const v2 scale{ rx, ry };
    const v2 recp_scale{ 1.f / rx, 1.f / ry };
    const v2 p0 = point0 * recp_scale, p1 = point1 * recp_scale;
    const v2 d = p1 - p0;
    const f_t a = dot(d, d), b = 2 * dot(p0, d), c = dot(p0, p0) - 1;
    const f_t det2 = b * b - 4 * a * c;
    if (det2 <= 0 || a == 0) return {};
    const f_t recp_a = 1 / a;
    const f_t det = std::sqrt(det2);
    f_t t1 = .5f * (-b - sign(b) * det) * recp_a;
    f_t t2 = t1 == 0 ? -b * recp_a : c * recp_a / t1;
    if (t1 > t2) std::swap(t1, t2);
    const bool u1valid = line || (t1 >= 0 && 1 >= t1);
    const bool u2valid = line || (t2 >= 0 && 1 >= t2);
    intersect_edge_circle_ret_t ret; ret.t1 = t1; ret.t2 = t2;
    if (!u1valid && !u2valid) { ret.points = 0; return ret; }
    if (u1valid && u2valid) { ret.points = 2; ret.u1 = (p0 + t1 * d) * scale; ret.u2 = (p0 + t2 * d) * scale; return ret; }
    ret.points = 1;
    ret.u1 = (u1valid ? p0 + t1 * d : p0 + t2 * d) * scale;
    ret.t1 = u1valid ? t1 : t2; ret.t2 = u1valid ? t2 : t1;
    return ret;
}
// misc.hpp:40-72
struct intersect_edge_sphere_ret_t { f_t t1 = 0, t2 = 0; };
inline intersect_edge_sphere_ret_t intersect_edge_ellipsoid(v3 point0, v3 point1, v3 centre, v3 x, v3 y, v3 axes) {
    const v3 z = cross(x, y);
    point0 = point0 - centre; point1 = point1 - centre;
    const v3 p0 = v3{ dot(point0, x), dot(point0, y), dot(point0, z) } * v3{ 1.f, 1.f, 1.f } ;
    const v3 q0{ p0.x / axes.x, p0.y / axes.y, p0.z / axes.z };
    const v3 p1r{ dot(point1, x), dot(point1, y), dot(point1, z) };
    const v3 q1{ p1r.x / axes.x, p1r.y / axes.y, p1r.z / axes.z };
    const v3 d = q1 - q0;
    const f_t a = dot(d, d), b = dot(q0, d) * 2, c = dot(q0, q0) - 1;
    const f_t det2 = b * b - 4 * a * c;
    if (det2 <= 0 || a == 0) return {};
    const f_t recp_a = 1 / a;
    const f_t det = std::sqrt(det2);
    f_t t1 = .5f * (-b - sign(b) * det) * recp_a;
    f_t t2 = t1 == 0 ? -b * recp_a : c * recp_a / t1;
    if (t1 > t2) std::swap(t1, t2);
    return { t1, t2 };
}

// include/wt/math/util.hpp:88-107
inline bool is_point_in_triangle(v3 p, v3 a, v3 b, v3 c) {
    const v3 v0 = b - a, v1 = c - a, u = p - a;
    const f_t d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(u, v0), d21 = dot(u, v1);
    const f_t d = diff_prod(d00, d11, d01, d01);
    const f_t sgn = d > 0 ? 1.f : -1.f;
    const f_t alpha = diff_prod(d11, d20, d01, d21);
    const f_t beta = diff_prod(d00, d21, d01, d20);
    return sgn * alpha >= 0 && sgn * beta >= 0 && sgn * (alpha + beta) <= sgn * d;
}

// include/wt/math/intersect/cone_intersection_tolerance.hpp:23-41
inline f_t cone_intersection_tolerance(v3 origin, const aabb_t& bb) {
    const f_t c0 = 4e-7f, c1 = 1e-6f, c2 = 1e-6f;
    const f_t obj_extent = 2 * std::max(max_element(absv(bb.min)), max_element(absv(bb.max)));
    const v3 ce{ c1 * obj_extent, c1 * obj_extent, c1 * obj_extent };
    const v3 obj_err = (c0 + c2) * absv(origin) + ce;
    const v3 wrd_err = (c1 + c2) * absv(origin);
    return max_element(obj_err + wrd_err);
}

// ---- elliptic cone: include/wt/math/shapes/elliptic_cone.hpp:30-333
struct elliptic_cone_t {
    ray_t r;
    v3 tangent{ 1, 0, 0 };
    f_t x0 = 0;             // initial_x_length
    f_t one_over_e = 1, e = 1;
    f_t tan_alpha = 0;
    f_t z_apex = -inf;

    elliptic_cone_t() = default;
    void update_apex() { z_apex = (x0 != 0 || tan_alpha != 0) ? -x0 / tan_alpha : -inf; }
    // private ctor (elliptic_cone.hpp:313-330)
    static elliptic_cone_t make(const ray_t& r, v3 x, f_t x0, f_t tan_alpha, f_t one_over_e, f_t e) {
        elliptic_cone_t c; c.r = r; c.tangent = x; c.x0 = x0; c.one_over_e = one_over_e; c.e = e; c.tan_alpha = tan_alpha; c.update_apex(); return c;
    }
    // elliptic_cone.hpp:64-78 (eccentricity ctor)
    static elliptic_cone_t make_ecc(const ray_t& r, v3 x, f_t tan_alpha, f_t eccentricity, f_t x0) {
        const f_t ooe = std::sqrt(std::max(0.f, 1 - sqr(eccentricity)));
        return make(r, x, x0, tan_alpha, ooe, 1.f / ooe);
    }
    // elliptic_cone.hpp:50-53 (isotropic)
    static elliptic_cone_t make_iso(const ray_t& r, f_t tan_alpha, f_t x0) {
        return make_ecc(r, frame_t::build_orthogonal_frame(r.d).t, tan_alpha, 0, x0);
    }
    bool is_ray() const { return tan_alpha == 0 && x0 == 0; }
    const v3& o() const { return r.o; }
    const v3& d() const { return r.d; }
    const v3& x() const { return tangent; }
    v3 y() const { return cross(r.d, tangent); }
    frame_t frame() const { return { x(), y(), d() }; }
    void set_o(v3 o) { r.o = o; }
    void set_x0(f_t v) { x0 = v; update_apex(); }
    v2 axes(f_t z) const { const f_t rr = tan_alpha * z + x0; return { rr * 1.f, rr * one_over_e }; }
    bool contains_local(v3 p, range_t range = { 0, inf }) const {
        return range.contains(p.z) && z_apex <= p.z && sqr(p.x) + sqr(e * p.y) <= sqr(p.z * tan_alpha + x0);
    }
    // wide variant (fma), elliptic_cone.hpp:170-200
    bool contains_local_w(v3 p, range_t range) const {
        const f_t x2 = sqr(p.x);
        const f_t ey = p.y * e;
        const f_t ztx = std::fma(p.z, tan_alpha, x0);
        return z_apex <= p.z && (range.min <= p.z && range.max >= p.z) && (x2 + sqr(ey)) <= sqr(ztx);
    }
    bool contains(v3 p, range_t range = { 0, inf }) const { return contains_local(frame().to_local(p - r.o), range); }
    v2 project_local(v3 p, f_t z) const {                               // elliptic_cone.hpp:205-214
        const v2 xy{ p.x, p.y };
        const f_t z0 = p.z;
        const f_t scale = (tan_alpha * z + x0) / std::fabs(tan_alpha * z0 + x0);
        return (x0 == 0 && tan_alpha == 0) ? xy : xy * scale;
    }
    static elliptic_cone_t cone_through_ellipse(v3 x, v3 y, v3 n, const ray_t& ray, f_t tan_alpha, f_t* self_intersection_distance);
    static elliptic_cone_t cone_through_ellipsoid(v3 axes, const frame_t& axes_frame, const ray_t& ray, f_t tan_alpha);
};

// ---- cone tests: include/wt/math/intersect/cone.hpp
struct intersect_cone_edge_ret_t { v3 p0{}, p1{}; range_t range{}; int pts = 0; };
// cone.hpp:38-128 ; in_local == true everywhere the hot path calls it with local points, false for world
inline std::optional<intersect_cone_edge_ret_t> intersect_cone_edge(
        const elliptic_cone_t& cone, v3 p0, v3 p1, range_t range, bool in_local, bool ray = false, bool line = false, bool test_clip_planes = true) {
    v3 lp0, lp1;
    if (!in_local) { const frame_t f = cone.frame(); lp0 = f.to_local(p0 - cone.o()); lp1 = f.to_local(p1 - cone.o()); }
    else { lp0 = p0; lp1 = p1; }
    const bool p0closer = lp1.z > lp0.z;
    if (!p0closer) std::swap(lp0, lp1);
    const v3 p = lp0, l = lp1 - lp0;
    const f_t x0 = cone.x0, ta = cone.tan_alpha, e = cone.e;
    const f_t cs = p.z * ta + x0;
    const f_t epy = e * p.y, ely = e * l.y, lzta = l.z * ta;
    const f_t c = sqr(p.x) + diff_prod(epy, epy, cs, cs);
    const f_t b = 2 * eft_dot(v3{ p.x, epy, -lzta }, v3{ l.x, ely, cs });
    const f_t a = sqr(l.x) + diff_prod(ely, ely, lzta, lzta);
    const f_t D = b * b - 4 * a * c;
    if (D < 0) return std::nullopt;
    const f_t sqrtD = std::sqrt(D);
    f_t t1 = b >= 0 ? (-b - sqrtD) / (2 * a) : (-b + sqrtD) / (2 * a);
    f_t t2 = (-b / a) - t1;
    const f_t zapex = cone.z_apex;
    if (p.z + t1 * l.z <= zapex) t1 = inf;
    if (p.z + t2 * l.z < zapex) t2 = inf;
    if (t2 < t1) std::swap(t1, t2);
    f_t z1 = t1 < inf ? p.z + t1 * l.z : -inf;
    f_t z2 = t2 < inf ? p.z + t2 * l.z : inf;
    if (z1 > range.max || z2 < range.min || (!std::isfinite(z1) && !std::isfinite(z2))) return std::nullopt;
    if (range.min > zapex && z1 < range.min) {
        if (test_clip_planes) if (const auto tmin = intersect_line_plane(p, p + l, v3{ 0, 0, range.min }, v3{ 0, 0, 1 }); tmin) { t1 = *tmin; z1 = range.min; }
    }
    if (z2 > range.max) {
        if (test_clip_planes) if (const auto tmax = intersect_line_plane(p, p + l, v3{ 0, 0, range.max }, v3{ 0, 0, 1 }); tmax) { t2 = *tmax; z2 = range.max; }
    }
    bool has1 = false, has2 = false; v3 v1{}, v2v{};
    const v3 base = p0closer ? p0 : p1;
    const v3 dir = p0closer ? p1 - p0 : p0 - p1;
    if (line || (t1 >= 0 && (ray || 1 >= t1))) { v1 = base + t1 * dir; has1 = true; } else z1 = z2;
    if (line || (t2 >= 0 && (ray || 1 >= t2))) { v2v = base + t2 * dir; has2 = true; } else z2 = z1;
    if (!has1 && !has2) return std::nullopt;
    intersect_cone_edge_ret_t ret;
    ret.range = { z1, z2 };
    ret.pts = has1 && has2 ? 2 : 1;
    ret.p0 = has1 ? v1 : v2v;
    if (has1 && has2) ret.p1 = v2v;
    return ret;
}

struct intersect_cone_plane_ret_t { range_t range; v3 near_{}, far_{}; };
// cone.hpp:170-258
inline intersect_cone_plane_ret_t intersect_cone_plane(const elliptic_cone_t& cone, v3 n, f_t d, range_t range, bool in_local) {
    const frame_t frame = cone.frame();
    if (!in_local) { d -= dot(cone.o(), n); n = frame.to_local(n); }
    const f_t x0 = cone.x0;
    const f_t e = cone.one_over_e;
    const f_t v_denom2 = sqr(n.x) + sqr(e * n.y);
    const v2 v = v_denom2 > 0 ? v2{ n.x, e * n.y } / std::sqrt(v_denom2) : v2{ 0, 0 };
    const v2 u = v * v2{ 1, e };
    const f_t nu = dot(n, v3{ u.x, u.y, 0 });
    const f_t zapex = cone.z_apex;
    f_t z01 = (d - x0 * nu) / (n.z + cone.tan_alpha * nu);
    f_t z02 = (d + x0 * nu) / (n.z - cone.tan_alpha * nu);
    const bool has_z01 = z01 >= zapex && !std::isnan(z01);
    const bool has_z02 = z02 >= zapex && !std::isnan(z02);
    if (!has_z01) z01 = inf;
    if (!has_z02) z02 = inf;
    const f_t s1 = z01 * cone.tan_alpha + x0, s2 = z02 * cone.tan_alpha + x0;
    v3 p1 = has_z01 ? v3{ s1 * u.x, s1 * u.y, z01 } : v3{ inf, inf, inf };
    v3 p2 = has_z02 ? v3{ s2 * (-u.x), s2 * (-u.y), z02 } : v3{ inf, inf, inf };
    if (z01 > z02) { std::swap(z01, z02); std::swap(p1, p2); }
    range_t rng{ z01, z02 };
    const bool empty = (!has_z01 && !has_z02) || (rng & range).empty();
    if (empty) return { range_t::null() };
    auto closest = [](f_t z, v2 u, v3 n, f_t d) {
        f_t x0, y0;
        if (std::fabs(n.y) > std::fabs(n.x)) { y0 = (d - n.z * z) / n.y; x0 = n.x != 0 ? (d - n.z * z - n.y * y0) / n.x : 0.f; }
        else { x0 = (d - n.z * z) / n.x; y0 = n.y != 0 ? (d - n.z * z - n.x * x0) / n.y : 0.f; }
        const f_t s = x0 * u.x + y0 * u.y;
        return v3{ s * u.x, s * u.y, z };
    };
    if (std::isfinite(rng.min)) {
        if (rng.min < range.min) { p1 = closest(range.min, v, n, d); rng.min = range.min; }
        if (!in_local) p1 = cone.o() + frame.to_world(p1);
    }
    const bool has_infinite = has_z01 != has_z02;
    if (std::isfinite(rng.max) || has_infinite) {
        if (rng.max > range.max) { p2 = closest(range.max, v, n, d); rng.max = range.max; }
        if (!in_local) p2 = cone.o() + frame.to_world(p2);
    }
    return { rng, p1, p2 };
}

struct intersect_cone_tri_ret_t { f_t dist = inf; v3 p{}; };
// cone.hpp:550-626
inline std::optional<intersect_cone_tri_ret_t> intersect_cone_tri(const elliptic_cone_t& cone, v3 a, v3 b, v3 c, v3 n, range_t range) {
    if (cone.is_ray()) {
        const auto cr = intersect_ray_tri(cone.r, a, b, c, range);
        if (cr) return intersect_cone_tri_ret_t{ cr->dist, cone.r.propagate(cr->dist) };
        return std::nullopt;
    }
    const frame_t frame = cone.frame();
    const v3 o = cone.o();
    const v3 vs[3] = { frame.to_local(a - o), frame.to_local(b - o), frame.to_local(c - o) };
    const v3 ln = frame.to_local(n);
    bool contains[3];
    for (int i = 0; i < 3; ++i) contains[i] = cone.contains_local_w(vs[i], range);
    const f_t closest_z = min3(vs[0].z, vs[1].z, vs[2].z);
    const f_t farthest_z = max3(vs[0].z, vs[1].z, vs[2].z);
    if (farthest_z < range.min || closest_z > range.max) return std::nullopt;
    for (int i = 0; i < 3; ++i)
        if (contains[i] && vs[i].z == closest_z)
            return intersect_cone_tri_ret_t{ closest_z, frame.to_world(vs[i]) + o };
    const auto icp = intersect_cone_plane(cone, ln, dot(vs[0], ln), range, true);
    if (!icp.range.empty()) {
        if (is_point_in_triangle(icp.near_, vs[0], vs[1], vs[2]))
            return intersect_cone_tri_ret_t{ icp.range.min, frame.to_world(icp.near_) + o };
    }
    bool hasp = false; v3 p{};
    for (int i = 0; i < 3; ++i) {
        const int j = (i + 1) % 3;
        const v3 ea = vs[i], eb = vs[j];
        if (contains[i] && contains[j]) continue;
        if (ea.z > range.max && eb.z > range.max) continue;
        if (ea.z < range.min && eb.z < range.min) continue;
        const auto cp = intersect_cone_edge(cone, ea, eb, range, true);
        if (cp && (!hasp || p.z > cp->p0.z)) { p = cp->p0; hasp = true; }
    }
    if (!hasp) return std::nullopt;
    return intersect_cone_tri_ret_t{ p.z, frame.to_world(p) + o };
}

// cone.hpp:479-539
inline bool test_cone_tri(const elliptic_cone_t& cone, v3 a, v3 b, v3 c, range_t range) {
    if (test_ray_tri(cone.r, a, b, c, range)) return true;
    const frame_t frame = cone.frame();
    const v3 o = cone.o();
    const v3 vs[3] = { frame.to_local(a - o), frame.to_local(b - o), frame.to_local(c - o) };
    if (max3(vs[0].z, vs[1].z, vs[2].z) < range.min || min3(vs[0].z, vs[1].z, vs[2].z) > range.max) return false;
    bool contains[3];
    for (int i = 0; i < 3; ++i) contains[i] = cone.contains_local_w(vs[i], range);
    if (contains[0] || contains[1] || contains[2] ||
        intersect_cone_edge(cone, vs[0], vs[1], range, true) ||
        intersect_cone_edge(cone, vs[0], vs[2], range, true) ||
        intersect_cone_edge(cone, vs[1], vs[2], range, true))
        return true;
    if (range.min <= 0) return false;
    v2 Ns[2], Fs[2]; int ns = 0, fs = 0;
    for (int i = 0; i < 3; ++i) {
        const int j = (i + 1) % 3;
        const auto np = range.min > 0 ? intersect_edge_plane(vs[i], vs[j], v3{ 0, 0, range.min }, v3{ 0, 0, 1 }) : std::nullopt;
        const auto fp = range.max < inf ? intersect_edge_plane(vs[i], vs[j], v3{ 0, 0, range.max }, v3{ 0, 0, 1 }) : std::nullopt;
        if (np && ns < 2) Ns[ns++] = { np->x, np->y };
        if (fp && fs < 2) Fs[fs++] = { fp->x, fp->y };
    }
    if (ns == 2) { const v2 ax = cone.axes(range.min); if (intersect_edge_ellipse(Ns[0], Ns[1], ax.x, ax.y).points > 0) return true; }
    if (fs == 2) { const v2 ax = cone.axes(range.max); if (intersect_edge_ellipse(Fs[0], Fs[1], ax.x, ax.y).points > 0) return true; }
    return false;
}

// src/math/elliptic_cone.cpp:19-88
inline elliptic_cone_t elliptic_cone_t::cone_through_ellipse(v3 x, v3 y, v3 n, const ray_t& ray, f_t tan_alpha, f_t* sid) {
    const bool xz = x.x == 0 && x.y == 0 && x.z == 0, yz = y.x == 0 && y.y == 0 && y.z == 0;
    if (xz && yz) {
        if (sid) *sid = 0;
        return make(ray, frame_t::build_orthogonal_frame(ray.d).t, 0, tan_alpha, 1, 1);
    }
    const frame_t of = frame_t::build_orthogonal_frame(ray.d);
    const v3 xl = of.to_local(x), yl = of.to_local(y);
    const v2 xhat{ xl.x, xl.y }, yhat{ yl.x, yl.y };
    const SVD_ret_t svd = SVD(mat2{ xhat, yhat });
    v2 X{ svd.Ucos, -svd.Usin };
    f_t lX = std::fabs(svd.sigma1), lY = std::fabs(svd.sigma2);
    if (lX < lY) { std::swap(lX, lY); X = { svd.Usin, svd.Ucos }; }
    const f_t e = lY > 0 ? std::sqrt(lX / lY) : 1.f;
    const v3 wx = of.to_world(X);
    const elliptic_cone_t cone = make(ray, wx, lX, tan_alpha, 1 / e, e);
    if (sid) {
        const auto cp = intersect_cone_plane(cone, n, dot(n, ray.o), { 0, inf }, false);
        *sid = cp.range.empty() ? 0.f : cp.range.max;
    }
    return cone;
}
// src/math/elliptic_cone.cpp:90-145
inline elliptic_cone_t elliptic_cone_t::cone_through_ellipsoid(v3 axes, const frame_t& axes_frame, const ray_t& ray, f_t tan_alpha) {
    const v3 wolocal = axes_frame.to_local(ray.d);
    const frame_t frame = frame_t::build_orthogonal_frame(wolocal);
    const v3 t = axes;
    const v3 nn = normalize(t * wolocal);
    const frame_t fc = frame_t::build_orthogonal_frame(nn);
    const v3 t1 = t * fc.t, t2 = t * fc.b;
    const mat2 A{ frame.to_local2(v2{ t1.x, t1.y }), frame.to_local2(v2{ t2.x, t2.y }) };
    if (A.m[0][0] * A.m[1][1] == A.m[1][0] * A.m[0][1])
        return make(ray, frame_t::build_orthogonal_frame(ray.d).t, 0, tan_alpha, 1, 1);
    const SVD_ret_t svd = SVD(A);
    v2 X{ svd.Ucos, -svd.Usin };
    f_t lX = std::fabs(svd.sigma1), lY = std::fabs(svd.sigma2);
    if (lX < lY) { std::swap(lX, lY); X = { svd.Usin, svd.Ucos }; }
    const f_t e = lY > 0 ? std::sqrt(lX / lY) : 1.f;
    const v3 X3 = normalize(frame.to_world(X));
    return make(ray, axes_frame.to_world(X3), lX, tan_alpha, 1 / e, e);
}

} // namespace ot
