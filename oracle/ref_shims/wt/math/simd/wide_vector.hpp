// ORACLE -- TEST INFRASTRUCTURE ONLY.
// Build shim for oracle/_ref: stands in for /root/reference/include/wt/math/simd/wide_vector.hpp (AVX 4- / 8- / 16-wide vectors over mp-units
// quantities).  The headers compiled into oracle/_ref only NAME the wide types -- in member templates and in result structs of the wide
// (AVX) entry points, none of which is instantiated here -- so declarations suffice.
#pragma once
#include <cstddef>
#include <wt/math/common.hpp>
#ifndef WT_SHIM_WIDE_LANES
namespace wt {
template <std::size_t W> struct f_w_t;
template <std::size_t W> struct b_w_t;
template <std::size_t W> struct length_w_t;
template <std::size_t W> struct vec3_w_t;
template <std::size_t W> struct pqvec3_w_t;
}
#else
// WT_SHIM_WIDE_LANES (oracle/ref_cone.cpp only): the wide types as plain arrays of lanes, with exactly the members and operators that
// frame_t::to_local(pqvec3_w_t) (frame.hpp:194-200), elliptic_cone_t::contains_local(pqvec3_w_t, range) (elliptic_cone.hpp:178-193) and the
// scalar-entry cone-triangle tests (cone.hpp:479-626) use.  Each lane operation is the single IEEE operation the AVX instruction performs
// (vsubps, vmulps, vaddps, vfmadd, vcmpps); the mixed-quantity dot product is the multiply + two fused multiply-adds of simd/math.hpp:536-541.
#include <bitset>
#include <cmath>
#include <limits>
namespace wt {
template <std::size_t W> struct b_w_t {
    bool v[W];
    std::bitset<W> to_bitmask() const noexcept { std::bitset<W> r; for (std::size_t i = 0; i < W; ++i) r[i] = v[i]; return r; }
};
template <std::size_t W> inline b_w_t<W> operator&&(const b_w_t<W>& a, const b_w_t<W>& b) noexcept { b_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = a.v[i] && b.v[i]; return r; }
template <std::size_t W> inline b_w_t<W>& operator&=(b_w_t<W>& a, const b_w_t<W>& b) noexcept { for (std::size_t i = 0; i < W; ++i) a.v[i] = a.v[i] && b.v[i]; return a; }
template <std::size_t W> struct f_w_t {
    f_t v[W];
    f_w_t() = default;
    explicit f_w_t(f_t s) noexcept { for (std::size_t i = 0; i < W; ++i) v[i] = s; }
    template <int I> f_t reads() const noexcept { return v[I]; }
    f_t read(int i) const noexcept { return v[i]; }
    static f_w_t one() noexcept { return f_w_t(f_t(1)); }
    static f_w_t zero() noexcept { return f_w_t(f_t(0)); }
    static f_w_t inf() noexcept { return f_w_t(std::numeric_limits<f_t>::infinity()); }
};
template <std::size_t W> struct length_w_t : f_w_t<W> {
    length_w_t() = default;
    explicit length_w_t(f_t s) noexcept : f_w_t<W>(s) {}
    length_w_t(const f_w_t<W>& b) noexcept : f_w_t<W>(b) {}
};
template <std::size_t W> inline f_w_t<W> operator+(const f_w_t<W>& a, const f_w_t<W>& b) noexcept { f_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = a.v[i] + b.v[i]; return r; }
template <std::size_t W> inline f_w_t<W> operator-(const f_w_t<W>& a, const f_w_t<W>& b) noexcept { f_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = a.v[i] - b.v[i]; return r; }
template <std::size_t W> inline f_w_t<W> operator*(const f_w_t<W>& a, const f_w_t<W>& b) noexcept { f_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = a.v[i] * b.v[i]; return r; }
template <std::size_t W> inline b_w_t<W> operator<=(const f_w_t<W>& a, const f_w_t<W>& b) noexcept { b_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = a.v[i] <= b.v[i]; return r; }
template <std::size_t W> inline b_w_t<W> operator>=(const f_w_t<W>& a, const f_w_t<W>& b) noexcept { b_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = a.v[i] >= b.v[i]; return r; }
template <std::size_t W> inline b_w_t<W> operator<(const f_w_t<W>& a, const f_w_t<W>& b) noexcept { b_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = a.v[i] < b.v[i]; return r; }
template <std::size_t W> inline b_w_t<W> operator>(const f_w_t<W>& a, const f_w_t<W>& b) noexcept { b_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = a.v[i] > b.v[i]; return r; }
template <std::size_t W> inline f_w_t<W> operator/(const f_w_t<W>& a, const f_w_t<W>& b) noexcept { f_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = a.v[i] / b.v[i]; return r; }
template <std::size_t W> inline f_w_t<W> operator-(const f_w_t<W>& a) noexcept { f_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = -a.v[i]; return r; }
template <std::size_t W> inline b_w_t<W> operator!=(const f_w_t<W>& a, const f_w_t<W>& b) noexcept { b_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = a.v[i] != b.v[i]; return r; }
template <std::size_t W> inline b_w_t<W> operator>=(const f_w_t<W>& a, zero_t) noexcept { b_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = a.v[i] >= f_t(0); return r; }
namespace simd { struct unaligned_data_t {}; inline constexpr unaligned_data_t unaligned_data{}; }
template <std::size_t W> struct vec3_w_t {
    f_w_t<W> c[3];
    vec3_w_t() = default;
    vec3_w_t(const f_t* x, const f_t* y, const f_t* z, simd::unaligned_data_t) noexcept { for (std::size_t i = 0; i < W; ++i) { c[0].v[i] = x[i]; c[1].v[i] = y[i]; c[2].v[i] = z[i]; } }      // (W consecutive values per component)
    explicit vec3_w_t(const vec3_t& s) noexcept : c{ f_w_t<W>(s.x), f_w_t<W>(s.y), f_w_t<W>(s.z) } {}
    const f_w_t<W>& x() const noexcept { return c[0]; }
    const f_w_t<W>& y() const noexcept { return c[1]; }
    const f_w_t<W>& z() const noexcept { return c[2]; }
};
template <std::size_t W> struct pqvec3_w_t {
    length_w_t<W> c[3];
    pqvec3_w_t() = default;
    explicit pqvec3_w_t(const pqvec3_t& s) noexcept : c{ length_w_t<W>(s.x), length_w_t<W>(s.y), length_w_t<W>(s.z) } {}
    explicit pqvec3_w_t(const f_w_t<W>& s) noexcept : c{ s, s, s } {}                 // (a wide scalar to all three components: wide_vector.hpp:715)
    pqvec3_w_t(const length_w_t<W>& x, const length_w_t<W>& y, const length_w_t<W>& z) noexcept : c{ x, y, z } {}
    pqvec3_w_t(const f_t* x, const f_t* y, const f_t* z, simd::unaligned_data_t) noexcept { for (std::size_t i = 0; i < W; ++i) { c[0].v[i] = x[i]; c[1].v[i] = y[i]; c[2].v[i] = z[i]; } }
    pqvec3_w_t(const pqvec3_t& p0, const pqvec3_t& p1, const pqvec3_t& p2, const pqvec3_t& p3) noexcept requires (W == 4) {
        const pqvec3_t* p[4] = { &p0, &p1, &p2, &p3 };
        for (int i = 0; i < 4; ++i) { c[0].v[i] = p[i]->x; c[1].v[i] = p[i]->y; c[2].v[i] = p[i]->z; }
    }
    const length_w_t<W>& x() const noexcept { return c[0]; }
    const length_w_t<W>& y() const noexcept { return c[1]; }
    const length_w_t<W>& z() const noexcept { return c[2]; }
    template <int I> pqvec3_t reads() const noexcept { return pqvec3_t{ c[0].v[I], c[1].v[I], c[2].v[I] }; }
    pqvec3_t read(int i) const noexcept { return pqvec3_t{ c[0].v[i], c[1].v[i], c[2].v[i] }; }
};
template <std::size_t W> inline pqvec3_w_t<W> operator-(const pqvec3_w_t<W>& a, const pqvec3_w_t<W>& b) noexcept { return pqvec3_w_t<W>{ a.c[0] - b.c[0], a.c[1] - b.c[1], a.c[2] - b.c[2] }; }
template <std::size_t W> inline pqvec3_w_t<W> operator+(const pqvec3_w_t<W>& a, const pqvec3_w_t<W>& b) noexcept { return pqvec3_w_t<W>{ a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2] }; }
template <std::size_t W> inline pqvec3_w_t<W> operator*(const pqvec3_w_t<W>& a, const vec3_w_t<W>& b) noexcept { return pqvec3_w_t<W>{ a.c[0] * b.c[0], a.c[1] * b.c[1], a.c[2] * b.c[2] }; }
// a mask made of the raw bits of a vector of numbers (wide_vector.hpp:225-228); blendv looks at the sign bit
template <std::size_t W> struct bvec3_w_t {
    b_w_t<W> c[3];
    explicit bvec3_w_t(const vec3_w_t<W>& s) noexcept { for (int k = 0; k < 3; ++k) for (std::size_t i = 0; i < W; ++i) c[k].v[i] = std::signbit(s.c[k].v[i]); }
};
using pqvec3_w4_t = pqvec3_w_t<4>;
using pqvec3_w8_t = pqvec3_w_t<8>; using vec3_w8_t = vec3_w_t<8>; using f_w8_t = f_w_t<8>; using length_w8_t = length_w_t<8>; using bvec3_w8_t = bvec3_w_t<8>; using b_w8_t = b_w_t<8>;
namespace m {
// simd/math.hpp:409-416 over vblendvps: b where the mask is set, else a
template <std::size_t W> inline f_w_t<W> selectv(const f_w_t<W>& a, const f_w_t<W>& b, const b_w_t<W>& mask) noexcept { f_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = mask.v[i] ? b.v[i] : a.v[i]; return r; }
template <std::size_t W> inline pqvec3_w_t<W> selectv(const pqvec3_w_t<W>& a, const pqvec3_w_t<W>& b, const bvec3_w_t<W>& mask) noexcept { return pqvec3_w_t<W>{ selectv<W>(a.c[0], b.c[0], mask.c[0]), selectv<W>(a.c[1], b.c[1], mask.c[1]), selectv<W>(a.c[2], b.c[2], mask.c[2]) }; }
// vmaxps / vminps: the SECOND operand when the comparison is false, i.e. also when either is NaN (simd_avx.hpp:244-260); four arguments pair up
// as (v1, v2), (v3, v4) (simd/math.hpp:333-356)
template <std::size_t W> inline f_w_t<W> max(const f_w_t<W>& a, const f_w_t<W>& b) noexcept { f_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = a.v[i] > b.v[i] ? a.v[i] : b.v[i]; return r; }
template <std::size_t W> inline f_w_t<W> min(const f_w_t<W>& a, const f_w_t<W>& b) noexcept { f_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = a.v[i] < b.v[i] ? a.v[i] : b.v[i]; return r; }
template <std::size_t W> inline f_w_t<W> max(const length_w_t<W>& a, const length_w_t<W>& b) noexcept { return max<W>(static_cast<const f_w_t<W>&>(a), static_cast<const f_w_t<W>&>(b)); }
template <std::size_t W> inline f_w_t<W> min(const length_w_t<W>& a, const length_w_t<W>& b) noexcept { return min<W>(static_cast<const f_w_t<W>&>(a), static_cast<const f_w_t<W>&>(b)); }
template <std::size_t W> inline f_w_t<W> max(const f_w_t<W>& a, const f_w_t<W>& b, const f_w_t<W>& c, const f_w_t<W>& d) noexcept { return max<W>(max<W>(a, b), max<W>(c, d)); }
template <std::size_t W> inline f_w_t<W> min(const f_w_t<W>& a, const f_w_t<W>& b, const f_w_t<W>& c, const f_w_t<W>& d) noexcept { return min<W>(min<W>(a, b), min<W>(c, d)); }
// simd/math.hpp:358-367: min(max(v, lo), hi) on vmaxps / vminps
template <std::size_t W> inline f_w_t<W> clamp(const f_w_t<W>& v, const f_w_t<W>& lo, const f_w_t<W>& hi) noexcept { return min<W>(max<W>(v, lo), hi); }
template <std::size_t W> inline f_w_t<W> fms(const f_w_t<W>& a, const f_w_t<W>& b, const f_w_t<W>& c) noexcept { f_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = std::fma(a.v[i], b.v[i], -c.v[i]); return r; }
// simd/math.hpp:494-504 (eft::diff_prod on wide vectors) and :543-551 (cross)
template <std::size_t W> inline f_w_t<W> diff_prod_w(const f_w_t<W>& a, const f_w_t<W>& b, const f_w_t<W>& c, const f_w_t<W>& d) noexcept { const auto cd = c * d; const auto diff = fms<W>(a, b, cd); const auto err = fms<W>(c, d, cd); return diff - err; }
template <std::size_t W, typename U, typename V> inline pqvec3_w_t<W> cross_w(const U& u, const V& v) noexcept {
    return pqvec3_w_t<W>{ diff_prod_w<W>(u.y(), v.z(), u.z(), v.y()), diff_prod_w<W>(u.z(), v.x(), u.x(), v.z()), diff_prod_w<W>(u.x(), v.y(), u.y(), v.x()) };
}
template <std::size_t W> inline pqvec3_w_t<W> cross(const vec3_w_t<W>& u, const pqvec3_w_t<W>& v) noexcept { return cross_w<W>(u, v); }
template <std::size_t W> inline pqvec3_w_t<W> cross(const pqvec3_w_t<W>& u, const pqvec3_w_t<W>& v) noexcept { return cross_w<W>(u, v); }
template <std::size_t W> inline f_w_t<W> fma(const f_w_t<W>& a, const f_w_t<W>& b, const f_w_t<W>& c) noexcept { f_w_t<W> r; for (std::size_t i = 0; i < W; ++i) r.v[i] = std::fma(a.v[i], b.v[i], c.v[i]); return r; }
// simd/math.hpp:536-541 (quantities of different kinds: lengths . pure numbers)
template <std::size_t W, typename U, typename V> inline f_w_t<W> dot_w(const U& u, const V& v) noexcept {
    f_w_t<W> sum = u.x() * v.x();
    sum = fma<W>(u.y(), v.y(), sum);
    return fma<W>(u.z(), v.z(), sum);
}
template <std::size_t W> inline f_w_t<W> dot(const pqvec3_w_t<W>& u, const vec3_w_t<W>& v) noexcept { return dot_w<W>(u, v); }
template <std::size_t W> inline f_w_t<W> dot(const vec3_w_t<W>& u, const pqvec3_w_t<W>& v) noexcept { return dot_w<W>(u, v); }
template <std::size_t W> inline f_w_t<W> dot(const pqvec3_w_t<W>& u, const pqvec3_w_t<W>& v) noexcept { return dot_w<W>(u, v); }
template <std::size_t W> inline f_w_t<W> dot(const vec3_w_t<W>& u, const vec3_w_t<W>& v) noexcept { return dot_w<W>(u, v); }
template <std::size_t W> inline bool any(const b_w_t<W>& m) noexcept { for (std::size_t i = 0; i < W; ++i) if (m.v[i]) return true; return false; }
}
}
#endif
