// ORACLE -- TEST INFRASTRUCTURE ONLY.
// Build shim for oracle/_ref: stands in for /root/reference/include/wt/math/common.hpp (an umbrella over glm and mp-units, both absent here --
// SURVEY.md 8c) with just the names the reference headers compiled into oracle/_ref use, so that those headers compile UNMODIFIED from where
// they lie.  Everything here has textbook semantics (std::complex, a 3-float vector, std::sqrt ...) except m::dot, which is the fma chain of
// the reference's include/wt/math/vecmath.hpp:21-66 as ot_math.h restates it.  The formulas under test are the reference's own text.
#pragma once
#include <type_traits>
#include <wt/util/concepts.hpp>
#include <concepts>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <complex>
#include <limits>
#include <algorithm>
namespace wt {
using f_t = float;                                   // the f32 build (CMakeLists.txt:214-216)
using c_t = std::complex<f_t>;
template <typename T> using limits = std::numeric_limits<T>;

// glm::vec2 / glm::mat2 as fsd.hpp uses them: mat2(c0, c1) takes COLUMNS; vec * mat is the row vector times the matrix,
// (v.x*m[0].x + v.y*m[0].y, v.x*m[1].x + v.y*m[1].y) (glm/detail/type_mat2x2.inl), plain multiply-adds
struct vec2_t {
    f_t x{}, y{};
    constexpr vec2_t() = default;
    constexpr vec2_t(f_t x_, f_t y_) : x(x_), y(y_) {}
#ifdef WT_SHIM_MAT4
    template <typename V> requires requires(const V& v) { v.x; v.y; v.z; } explicit constexpr vec2_t(const V& v) : x(v.x), y(v.y) {}      // glm: the xy part of a 3-vector
#endif
#ifdef WT_SHIM_DISTINCT_PQ
    template <typename V> requires (requires(const V& v) { v.x; v.y; } && !requires(const V& v) { v.z; } && !std::is_same_v<V, vec2_t>) explicit constexpr vec2_t(const V& v) : x(v.x), y(v.y) {}      // numbers out of a 2-vector of lengths (divided by the unit)
#endif
    constexpr vec2_t& operator/=(f_t s) { x /= s; y /= s; return *this; }
    constexpr f_t& operator[](std::size_t i) { return i == 0 ? x : y; }
    constexpr const f_t& operator[](std::size_t i) const { return i == 0 ? x : y; }
};
constexpr vec2_t operator*(f_t s, vec2_t v) { return { s * v.x, s * v.y }; }
constexpr vec2_t operator*(vec2_t v, f_t s) { return { v.x * s, v.y * s }; }
constexpr vec2_t operator*(vec2_t a, vec2_t b) { return { a.x * b.x, a.y * b.y }; }
constexpr vec2_t operator/(f_t s, vec2_t v) { return { s / v.x, s / v.y }; }
constexpr vec2_t operator/(vec2_t v, f_t s) { return { v.x / s, v.y / s }; }
constexpr vec2_t operator/(vec2_t a, vec2_t b) { return { a.x / b.x, a.y / b.y }; }
constexpr vec2_t operator+(vec2_t a, vec2_t b) { return { a.x + b.x, a.y + b.y }; }
constexpr vec2_t operator-(vec2_t a, vec2_t b) { return { a.x - b.x, a.y - b.y }; }
constexpr vec2_t operator-(vec2_t a) { return { -a.x, -a.y }; }
constexpr bool operator==(vec2_t a, vec2_t b) { return a.x == b.x && a.y == b.y; }
struct dir2_t : vec2_t {        // unit vector in the plane
    constexpr dir2_t() = default;
    constexpr dir2_t(f_t x_, f_t y_) : vec2_t(x_, y_) {}
    constexpr explicit dir2_t(const vec2_t& v) : vec2_t(v) {}
};
// glm::mat2: column-major; mat2(c0, c1) / mat2(x0, y0, x1, y1) take COLUMNS; m[i] is column i; vec * mat = row vector times matrix,
// mat * vec = matrix times column vector, mat * mat the usual product (glm/detail/type_mat2x2.inl: plain multiply-adds in this order)
struct mat2_t {
    vec2_t c[2];
    constexpr mat2_t() : c{ { 1, 0 }, { 0, 1 } } {}
    constexpr mat2_t(vec2_t c0, vec2_t c1) : c{ c0, c1 } {}
    constexpr mat2_t(f_t x0, f_t y0, f_t x1, f_t y1) : c{ { x0, y0 }, { x1, y1 } } {}
    constexpr vec2_t& operator[](std::size_t i) { return c[i]; }
    constexpr const vec2_t& operator[](std::size_t i) const { return c[i]; }
};
constexpr vec2_t operator*(vec2_t v, const mat2_t& m) { return { v.x * m.c[0].x + v.y * m.c[0].y, v.x * m.c[1].x + v.y * m.c[1].y }; }
constexpr vec2_t operator*(const mat2_t& m, vec2_t v) { return { m.c[0].x * v.x + m.c[1].x * v.y, m.c[0].y * v.x + m.c[1].y * v.y }; }
constexpr mat2_t operator*(const mat2_t& a, const mat2_t& b) {
    return { a.c[0].x * b.c[0].x + a.c[1].x * b.c[0].y, a.c[0].y * b.c[0].x + a.c[1].y * b.c[0].y,
             a.c[0].x * b.c[1].x + a.c[1].x * b.c[1].y, a.c[0].y * b.c[1].x + a.c[1].y * b.c[1].y };
}
constexpr mat2_t operator+(const mat2_t& a, const mat2_t& b) { return { a.c[0] + b.c[0], a.c[1] + b.c[1] }; }
constexpr mat2_t operator-(const mat2_t& a, const mat2_t& b) { return { a.c[0].x - b.c[0].x, a.c[0].y - b.c[0].y, a.c[1].x - b.c[1].x, a.c[1].y - b.c[1].y }; }
constexpr mat2_t operator*(const mat2_t& a, f_t s) { return { a.c[0] * s, a.c[1] * s }; }
constexpr mat2_t operator*(f_t s, const mat2_t& a) { return { a.c[0] * s, a.c[1] * s }; }
namespace u::ang { inline constexpr f_t rad = 1; }      // mp-units' radian: angles are plain f_t here
namespace u { inline constexpr f_t m = 1; }             // mp-units' metre: lengths are plain f_t here
using angle_t = f_t;
struct vec4_t { f_t x{}, y{}, z{}, w{}; };

struct vec3_t {
    f_t x{}, y{}, z{};
    constexpr vec3_t() = default;
    constexpr vec3_t(f_t x_, f_t y_, f_t z_) : x(x_), y(y_), z(z_) {}
    constexpr explicit vec3_t(f_t s) : x(s), y(s), z(s) {}
    constexpr vec3_t(const vec2_t& v, f_t z_) : x(v.x), y(v.y), z(z_) {}
    constexpr f_t& operator[](std::size_t i) { return i == 0 ? x : i == 1 ? y : z; }
    constexpr const f_t& operator[](std::size_t i) const { return i == 0 ? x : i == 1 ? y : z; }
};
constexpr vec3_t operator+(vec3_t a, vec3_t b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
constexpr vec3_t operator-(vec3_t a, vec3_t b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
constexpr vec3_t operator*(f_t s, vec3_t a) { return { s * a.x, s * a.y, s * a.z }; }
constexpr vec3_t operator*(vec3_t a, f_t s) { return { a.x * s, a.y * s, a.z * s }; }
constexpr bool operator==(const vec3_t& a, const vec3_t& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
// names math/linalg.hpp's one template (solve_linear_system2x2, not on the path and never instantiated here) is parsed against
template <int N, typename T> using vec = vec2_t;
template <typename T> using vec2 = vec2_t;
template <typename T> using mat2 = mat2_t;
template <typename T> using limits = std::numeric_limits<T>;
#ifndef WT_SHIM_DISTINCT_PQ
using pqvec2_t = vec2_t;
#else
// (ref_frame.cpp) math/frame.hpp overloads to_local / to_world on vectors of lengths vs plain vectors, so the two must be distinct types here too: a
// pq vector is three numbers in metres; the arithmetic on it is the plain vectors' (mp-units adds no operation, only the unit)
struct pqvec3_t;
struct pqvec2_t {
    f_t x{}, y{};
    constexpr pqvec2_t() = default;
    constexpr pqvec2_t(f_t x_, f_t y_) : x(x_), y(y_) {}
    constexpr pqvec2_t(const vec2_t& v) : x(v.x), y(v.y) {}            // (a plain vector times a length)
    explicit inline pqvec2_t(const pqvec3_t& v);                        // the xy part
};
constexpr pqvec2_t operator*(const pqvec2_t& a, f_t s) { return { a.x * s, a.y * s }; }
constexpr pqvec2_t operator/(const pqvec2_t& a, f_t s) { return { a.x / s, a.y / s }; }
constexpr vec2_t to_vec2(const pqvec2_t& a) { return { a.x, a.y }; }
constexpr vec2_t operator/(const pqvec2_t& a, const pqvec2_t& b) { return { a.x / b.x, a.y / b.y }; }      // lengths / lengths: numbers
constexpr pqvec2_t operator-(const pqvec2_t& a, const pqvec2_t& b) { return { a.x - b.x, a.y - b.y }; }
constexpr pqvec2_t operator+(const pqvec2_t& a, const pqvec2_t& b) { return { a.x + b.x, a.y + b.y }; }
constexpr pqvec2_t operator*(f_t s, const pqvec2_t& a) { return { s * a.x, s * a.y }; }
constexpr vec2_t operator/(f_t s, const pqvec2_t& a) { return { s / a.x, s / a.y }; }                    // 1 / lengths
constexpr vec2_t operator*(const pqvec2_t& a, const vec2_t& b) { return { a.x * b.x, a.y * b.y }; }      // lengths x (1 / lengths): numbers
constexpr pqvec2_t operator*(const vec2_t& a, const pqvec2_t& b) { return { a.x * b.x, a.y * b.y }; }    // numbers x lengths
#endif
using length_t = f_t;           // metres
using angle_t = f_t;            // radians
// mp-units semantics the pinned UTD code relies on: a wavenumber is held in 1/mm and lengths in m, so the dimensionless number k * l taken
// with u::to_num is (k * l) scaled by the unit ratio m/mm = 1000 (one f32 product by 1000 after the product of the two numerical values)
struct wavenumber_t { f_t per_mm; };
struct wavenumber_length_t { f_t mm_per_m_scaled; };
constexpr wavenumber_length_t operator*(wavenumber_t k, length_t l) { return { k.per_mm * l }; }
template <typename T> concept Angle = std::is_floating_point_v<T>;
template <typename T> concept Length = std::is_floating_point_v<T>;
template <typename T> concept Area = std::is_floating_point_v<T>;      // a length is a plain f_t here
template <typename T> concept Wavenumber = std::is_same_v<T, wavenumber_t>;
namespace u { constexpr f_t to_m(f_t v) { return v; } constexpr vec2_t to_num(const vec2_t& v) { return v; } constexpr f_t to_num(f_t v) { return v; } constexpr f_t to_num(wavenumber_length_t v) { return v.mm_per_m_scaled * f_t(1000); } }
#ifndef WT_SHIM_DISTINCT_PQ
using pqvec3_t = vec3_t;        // mp-units' vector of lengths: plain floats here
#else
struct pqvec3_t {
    f_t x{}, y{}, z{};
    static constexpr pqvec3_t zero() { return {}; }
    static constexpr pqvec3_t infinity() { return { std::numeric_limits<f_t>::infinity(), std::numeric_limits<f_t>::infinity(), std::numeric_limits<f_t>::infinity() }; }
    constexpr pqvec3_t() = default;
    constexpr pqvec3_t(f_t x_, f_t y_, f_t z_) : x(x_), y(y_), z(z_) {}
    constexpr pqvec3_t(const vec3_t& v) : x(v.x), y(v.y), z(v.z) {}
    constexpr pqvec3_t(const vec2_t& v, f_t z_) : x(v.x), y(v.y), z(z_) {}
    explicit constexpr pqvec3_t(f_t s) : x(s), y(s), z(s) {}       // (lengths: a plain 2-vector times a length, and a z)         // (a plain vector times a length, e.g. t * v.x with v.x in metres)
};
constexpr pqvec3_t operator-(const pqvec3_t& a, const vec3_t& b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
constexpr pqvec3_t operator-(const pqvec3_t& a, const pqvec3_t& b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
constexpr pqvec3_t operator+(const pqvec3_t& a, const pqvec3_t& b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
constexpr pqvec3_t& operator-=(pqvec3_t& a, const pqvec3_t& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
inline pqvec2_t::pqvec2_t(const pqvec3_t& v) : x(v.x), y(v.y) {}
constexpr pqvec3_t operator*(const pqvec3_t& a, const vec3_t& b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
constexpr pqvec3_t operator-(const pqvec3_t& a) { return { -a.x, -a.y, -a.z }; }
constexpr bool operator==(const pqvec3_t& a, const pqvec3_t& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
constexpr bool operator!=(const pqvec3_t& a, const pqvec3_t& b) { return !(a == b); }
constexpr pqvec3_t operator+(const pqvec3_t& a) { return a; }
} namespace glm { struct bvec3_standin { bool x, y, z; }; } namespace wt {
// boolean vectors and the glm-style select of the scalar ray-AABB entry point (math/intersect/ray.hpp:244-263; parsed, not under pin)
struct vec3b_t { bool x, y, z; };
} namespace glm { constexpr wt::vec3b_t equal(const wt::vec3_t& a, const wt::vec3_t& b) { return { a.x == b.x, a.y == b.y, a.z == b.z }; } } namespace wt {
constexpr vec3b_t operator<(const vec3_t& a, const vec3_t& b) { return { a.x < b.x, a.y < b.y, a.z < b.z }; }
constexpr bool operator<=(const pqvec3_t& a, const pqvec3_t& b) { return a.x <= b.x && a.y <= b.y && a.z <= b.z; }
constexpr bool operator>=(const pqvec3_t& a, const pqvec3_t& b) { return a.x >= b.x && a.y >= b.y && a.z >= b.z; }
constexpr pqvec3_t& operator+=(pqvec3_t& a, const pqvec3_t& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
constexpr pqvec3_t operator*(f_t s, const pqvec3_t& a) { return { s * a.x, s * a.y, s * a.z }; }
constexpr pqvec3_t operator/(const pqvec3_t& a, f_t s) { return { a.x / s, a.y / s, a.z / s }; }       // (by a number, or by a length: see dir3_t's constructor below)
constexpr vec3_t operator/(const pqvec3_t& a, const pqvec3_t& b) { return { a.x / b.x, a.y / b.y, a.z / b.z }; }      // lengths / lengths: numbers
struct dir2_t_tag {};
#endif
struct mat3_t {                          // glm::mat3, column-major: mat3(x0,y0,z0, x1,...) takes COLUMNS; m[i] is column i
    vec3_t c[3];
    constexpr mat3_t() : c{ { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } } {}
    constexpr mat3_t(f_t x0, f_t y0, f_t z0, f_t x1, f_t y1, f_t z1, f_t x2, f_t y2, f_t z2) : c{ { x0, y0, z0 }, { x1, y1, z1 }, { x2, y2, z2 } } {}
    constexpr vec3_t& operator[](std::size_t i) { return c[i]; }
    constexpr const vec3_t& operator[](std::size_t i) const { return c[i]; }
};
constexpr mat3_t operator*(const mat3_t& a, f_t s) { mat3_t r; for (int i = 0; i < 3; ++i) r.c[i] = vec3_t{ a.c[i].x * s, a.c[i].y * s, a.c[i].z * s }; return r; }
constexpr mat3_t operator+(const mat3_t& a, const mat3_t& b) { mat3_t r; for (int i = 0; i < 3; ++i) r.c[i] = vec3_t{ a.c[i].x + b.c[i].x, a.c[i].y + b.c[i].y, a.c[i].z + b.c[i].z }; return r; }
constexpr vec3_t operator*(const mat3_t& m, const vec3_t& v) { return { m.c[0].x * v.x + m.c[1].x * v.y + m.c[2].x * v.z, m.c[0].y * v.x + m.c[1].y * v.y + m.c[2].y * v.z, m.c[0].z * v.x + m.c[1].z * v.y + m.c[2].z * v.z }; }
// unit vector (include/wt/math/unit_vector/unit_vector.hpp): a vec3 with explicit construction from one
struct dir3_t : vec3_t {
    constexpr dir3_t() = default;
    constexpr dir3_t(f_t x_, f_t y_, f_t z_) : vec3_t(x_, y_, z_) {}
    constexpr explicit dir3_t(const vec3_t& v) : vec3_t(v) {}
#ifdef WT_SHIM_DISTINCT_PQ
    constexpr explicit dir3_t(const pqvec3_t& v) : vec3_t{ v.x, v.y, v.z } {}      // a vector of lengths divided by its length (plain floats cannot tell: the quotient arrives as lengths)
#endif
    constexpr dir3_t operator-() const { return dir3_t{ -x, -y, -z }; }
};

// mp-units' `zero`: compares with any quantity; quantities are plain f_t here
struct zero_t { constexpr operator f_t() const noexcept { return 0; } };
template <std::size_t W> struct bvec3_w_t;
inline constexpr zero_t zero{};
template <typename T> requires std::is_same_v<T, c_t> inline bool operator==(const T& a, zero_t) noexcept { return a.real() == f_t(0) && a.imag() == f_t(0); }
constexpr vec3_t operator/(f_t s, const vec3_t& v) { return { s / v.x, s / v.y, s / v.z }; }
namespace m {
inline f_t fma(f_t a, f_t b, f_t c) noexcept { return std::fma(a, b, c); }
inline bool isnan(f_t v) noexcept { return std::isnan(v); }
template <typename A, typename B, typename C> auto selectv(const A&, const B&, const C&);      // (named by never-instantiated wide templates)
template <typename T> constexpr T pow(T base, std::size_t e) noexcept { T r = 1; for (std::size_t i = 0; i < e; ++i) r *= base; return r; }
using std::ceil; using std::log; using std::abs;
template <typename T> constexpr T min(T a, T b) noexcept { return std::min(a, b); }
template <typename T> constexpr T max(T a, T b) noexcept { return std::max(a, b); }
template <typename T> constexpr T sqr(T v) noexcept { return v * v; }
inline f_t sqrt(f_t v) noexcept { return std::sqrt(v); }
inline c_t sqrt(c_t v) noexcept { return std::sqrt(v); }                 // common.hpp:38-40: glm::sqrt(c) == std::sqrt
inline constexpr f_t two_pi = f_t(2. * 3.141592653589793238462643383279502884);            // math/defs.hpp:40
inline constexpr f_t inv_two_pi = f_t(0.318309886183790671537767526745028724 / 2.);        // math/defs.hpp:49
inline constexpr f_t pi_2 = f_t(3.141592653589793238462643383279502884 / 2.), pi_4 = f_t(3.141592653589793238462643383279502884 / 4.);   // math/defs.hpp
inline constexpr f_t inv_pi = f_t(0.318309886183790671537767526745028724), inv_four_pi = f_t(0.318309886183790671537767526745028724 / 4.);
// glm::inverse(mat2) (glm/detail/func_matrix.inl compute_inverse<2,2>): one reciprocal of the determinant, then four products
inline mat2_t inverse(const mat2_t& m) noexcept {
    const f_t ood = f_t(1) / (m.c[0].x * m.c[1].y - m.c[1].x * m.c[0].y);
    return mat2_t{ vec2_t{ m.c[1].y * ood, -m.c[0].y * ood }, vec2_t{ -m.c[1].x * ood, m.c[0].x * ood } };
}
inline constexpr f_t pi = f_t(3.141592653589793238462643383279502884);                     // math/defs.hpp
inline f_t cos(f_t v) noexcept { return std::cos(v); }
inline f_t sin(f_t v) noexcept { return std::sin(v); }
inline f_t fract(f_t v) noexcept { return v - std::floor(v); }
inline constexpr f_t inf = std::numeric_limits<f_t>::infinity();
inline constexpr f_t sqrt_pi = f_t(1.772453850905516027298167483341145183), inv_sqrt_pi = f_t(0.564189583547756286948079451560772586);   // math/defs.hpp
template <typename T> constexpr T min(T a, T b, T c) noexcept { return std::min(a, std::min(b, c)); }
template <typename T> constexpr T max(T a, T b, T c) noexcept { return std::max(a, std::max(b, c)); }
inline bool isfinite(f_t v) noexcept { return std::isfinite(v); }
inline f_t pow(f_t b, f_t e) noexcept { return std::pow(b, e); }
inline f_t determinant(const mat2_t& m) noexcept { return m.c[0].x * m.c[1].y - m.c[1].x * m.c[0].y; }          // glm::determinant
inline mat3_t outer(const vec3_t& c, const vec3_t& r) noexcept { mat3_t m; for (int i = 0; i < 3; ++i) m.c[i] = vec3_t{ c.x * r[i], c.y * r[i], c.z * r[i] }; return m; }     // glm::outerProduct
inline mat3_t transpose(const mat3_t& a) noexcept { return mat3_t{ a.c[0].x, a.c[1].x, a.c[2].x, a.c[0].y, a.c[1].y, a.c[2].y, a.c[0].z, a.c[1].z, a.c[2].z }; }
inline f_t determinant(const mat3_t& m) noexcept { return m.c[0].x * (m.c[1].y * m.c[2].z - m.c[2].y * m.c[1].z) - m.c[1].x * (m.c[0].y * m.c[2].z - m.c[2].y * m.c[0].z) + m.c[2].x * (m.c[0].y * m.c[1].z - m.c[1].y * m.c[0].z); }   // (only inside a debug assertion)
inline mat2_t transpose(const mat2_t& m) noexcept { return mat2_t{ m.c[0].x, m.c[1].x, m.c[0].y, m.c[1].y }; }
inline vec2_t iszero(vec2_t v) noexcept { return { f_t(v.x == 0), f_t(v.y == 0) }; }
inline bool all(vec2_t v) noexcept { return v.x != 0 && v.y != 0; }
namespace eft {     // math/eft/eft.hpp: compensated a*b - c*d (Kahan), as ot_math.h restates it
inline f_t diff_prod(f_t a, f_t b, f_t c, f_t d) noexcept { const f_t cd = c * d; const f_t r = std::fma(a, b, -cd); return r + std::fma(-c, d, cd); }
inline f_t sum_prod(f_t a, f_t b, f_t c, f_t d) noexcept { return diff_prod(a, b, -c, d); }          // eft.hpp:153-159
// eft.hpp:33-51,184-197 (Graillat / Menissier-Morain compensated dot product: two_prod by fma, two_sum, errors summed in order); the pinned
// intersect_cone_edge calls it on 3-vectors for the `b` coefficient of its quadratic
template <typename A, typename B> inline f_t dot(const A& v1, const B& v2) noexcept {
    f_t d = 0, err = 0;
    const f_t a[3] = { v1.x, v1.y, v1.z }, b[3] = { v2.x, v2.y, v2.z };
    for (int i = 0; i < 3; ++i) {
        const f_t prod = a[i] * b[i]; const f_t err1 = std::fma(a[i], b[i], -prod);
        const f_t sum = d + prod; const f_t e1 = sum - d; const f_t e2 = sum - e1; const f_t err2 = (prod - e1) + (d - e2);
        d = sum; err = err + err1 + err2;
    }
    return d + err;
}
}
inline constexpr f_t sqrt_two = f_t(1.41421356237309504880168872420969808);
inline constexpr f_t inv_sqrt_two = f_t(1. / 1.41421356237309504880168872420969808);                   // math/defs.hpp:57
inline constexpr f_t sqrt_pi_2 = f_t(1.253314137315500251207882642405522627), inv_sqrt_two_pi = f_t(0.398942280401432677939946059934381868);  // math/defs.hpp
inline f_t round(f_t v) noexcept { return std::round(v); }                                      // common.hpp:104-106 glm::round
inline f_t atan2(f_t y, f_t x) noexcept { return std::atan2(y, x); }
inline f_t acos(f_t v) noexcept { return std::acos(v); }                            // quantity/math.hpp:213-216
inline f_t cot(f_t a) noexcept { return f_t(1) / std::tan(a); }                                 // quantity/math.hpp:199-202
inline f_t mod(f_t a, f_t b) noexcept { return a - b * std::floor(a / b); }                      // quantity/math.hpp:84-88 glm::mod
inline bool isfinite(const c_t& v) noexcept { return std::isfinite(v.real()) && std::isfinite(v.imag()); }
inline f_t sign(f_t t) noexcept { return f_t((f_t(0) < t) - (t < f_t(0))); }                   // common.hpp:128-131 glm::sign
// common.hpp:257-264: the end points are returned exactly, otherwise glm::mix = a (1 - x) + b x
inline f_t mix(f_t a, f_t b, f_t x) noexcept { if (x == f_t(0)) return a; if (x == f_t(1)) return b; return a * (f_t(1) - x) + b * x; }                              // common.hpp:228 glm::fract
// glm::clamp = min(max(v, lo), hi) (common.hpp:238-243); clamp01 (:508-511)
template <typename T> constexpr T clamp(const T& v, const std::type_identity_t<T>& lo, const std::type_identity_t<T>& hi) noexcept { return std::min(std::max(v, lo), hi); }
template <typename T> constexpr T clamp01(const T& v) noexcept { return clamp<T>(v, 0, 1); }
inline double sqrt(double v) noexcept { return std::sqrt(v); }
inline double sqr(double v) noexcept { return v * v; }
inline bool isfinite(double v) noexcept { return std::isfinite(v); }
inline vec2_t mix(const vec2_t& a, const vec2_t& b, f_t x) noexcept {        // common.hpp:264-271: glm::mix on a vector = a (1 - x) + b x per component
    if (x == f_t(0)) return a; if (x == f_t(1)) return b;
    return { a.x * (f_t(1) - x) + b.x * x, a.y * (f_t(1) - x) + b.y * x };
}
inline f_t exp(f_t v) noexcept { return std::exp(v); }
// common.hpp:414-434 ("From boost")
inline f_t sinc(const f_t x) noexcept {
    constexpr f_t taylor_0_bound = std::numeric_limits<f_t>::epsilon();
    constexpr f_t taylor_2_bound = static_cast<f_t>(0.00034526698300124390839884978618400831996329879769945L);
    constexpr f_t taylor_n_bound = static_cast<f_t>(0.018581361171917516667460937040007436176452688944747L);
    if (std::abs(x) >= taylor_n_bound) return std::sin(x) / x;
    f_t result = 1;
    if (std::abs(x) >= taylor_0_bound) { const f_t x2 = x * x; result -= x2 / f_t(6); if (std::abs(x) >= taylor_2_bound) result += (x2 * x2) / f_t(120); }
    return result;
}
inline f_t dot(const vec2_t& a, const vec2_t& b) noexcept { return std::fma(a.y, b.y, a.x * b.x); }                          // vecmath.hpp:21-66
inline f_t length2(const vec2_t& v) noexcept { return dot(v, v); }
inline f_t dot(const vec3_t& a, const vec3_t& b) noexcept { return std::fma(a.z, b.z, std::fma(a.y, b.y, a.x * b.x)); }   // vecmath.hpp:21-66
inline f_t length2(const vec3_t& v) noexcept { return dot(v, v); }
inline f_t length(const vec2_t& v) noexcept { return std::sqrt(length2(v)); }                    // vecmath.hpp:38-44
inline f_t length(const vec3_t& v) noexcept { return std::sqrt(length2(v)); }
inline vec3_t cross(const vec3_t& x, const vec3_t& y) noexcept {                                 // vecmath.hpp:53-64
    return { eft::diff_prod(x.y, y.z, x.z, y.y), eft::diff_prod(x.z, y.x, x.x, y.z), eft::diff_prod(x.x, y.y, x.y, y.x) };
}
inline dir3_t normalize(const vec3_t& v) noexcept { const f_t l = std::sqrt(dot(v, v)); return dir3_t{ v.x / l, v.y / l, v.z / l }; }
#ifdef WT_SHIM_DISTINCT_PQ
inline f_t dot(const pqvec3_t& a, const pqvec3_t& b) noexcept { return std::fma(a.z, b.z, std::fma(a.y, b.y, a.x * b.x)); }
inline pqvec3_t cross(const vec3_t& x, const pqvec3_t& y) noexcept { return { eft::diff_prod(x.y, y.z, x.z, y.y), eft::diff_prod(x.z, y.x, x.x, y.z), eft::diff_prod(x.x, y.y, x.y, y.x) }; }
inline pqvec3_t cross(const pqvec3_t& x, const pqvec3_t& y) noexcept { return { eft::diff_prod(x.y, y.z, x.z, y.y), eft::diff_prod(x.z, y.x, x.x, y.z), eft::diff_prod(x.x, y.y, x.y, y.x) }; }
inline pqvec3_t cross(const pqvec3_t& x, const vec3_t& y) noexcept { return { eft::diff_prod(x.y, y.z, x.z, y.y), eft::diff_prod(x.z, y.x, x.x, y.z), eft::diff_prod(x.x, y.y, x.y, y.x) }; }
inline pqvec3_t mix(const pqvec3_t& a, const pqvec3_t& b, const vec3b_t& s) noexcept { return { s.x ? b.x : a.x, s.y ? b.y : a.y, s.z ? b.z : a.z }; }
inline pqvec3_t mix(const pqvec3_t& a, const pqvec3_t& b, bool s) noexcept { return s ? b : a; }
inline f_t max_element(const pqvec3_t& v) noexcept { return std::max(v.x, std::max(v.y, v.z)); }
inline pqvec3_t abs(const pqvec3_t& v) noexcept { return { std::fabs(v.x), std::fabs(v.y), std::fabs(v.z) }; }
inline f_t length2(const pqvec3_t& v) noexcept { return std::fma(v.z, v.z, std::fma(v.y, v.y, v.x * v.x)); }
inline f_t length(const pqvec3_t& v) noexcept { return std::sqrt(length2(v)); }
inline bool isfinite(const pqvec3_t& v) noexcept { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); }
inline pqvec3_t mix(const pqvec3_t& a, const pqvec3_t& b, f_t x) noexcept { if (x == f_t(0)) return a; if (x == f_t(1)) return b; return { a.x * (f_t(1) - x) + b.x * x, a.y * (f_t(1) - x) + b.y * x, a.z * (f_t(1) - x) + b.z * x }; }
inline f_t min_element(const pqvec3_t& v) noexcept { return std::min(v.x, std::min(v.y, v.z)); }
inline f_t dot(const pqvec2_t& a, const vec2_t& b) noexcept { return std::fma(a.y, b.y, a.x * b.x); }
inline f_t length(const pqvec2_t& v) noexcept { return std::sqrt(std::fma(v.y, v.y, v.x * v.x)); }
inline f_t max_element(const pqvec2_t& v) noexcept { return std::max(v.x, v.y); }
inline pqvec2_t mix(const pqvec2_t& a, const pqvec2_t& b, f_t x) noexcept { if (x == f_t(0)) return a; if (x == f_t(1)) return b; return { a.x * (f_t(1) - x) + b.x * x, a.y * (f_t(1) - x) + b.y * x }; }      // as the 2-vector of numbers above
inline f_t dot(const pqvec3_t& a, const vec3_t& b) noexcept { return std::fma(a.z, b.z, std::fma(a.y, b.y, a.x * b.x)); }
inline f_t dot(const vec3_t& a, const pqvec3_t& b) noexcept { return std::fma(a.z, b.z, std::fma(a.y, b.y, a.x * b.x)); }
inline dir3_t normalize(const pqvec3_t& v) noexcept { const f_t l = std::sqrt(std::fma(v.z, v.z, std::fma(v.y, v.y, v.x * v.x))); return dir3_t{ v.x / l, v.y / l, v.z / l }; }
struct bvec3_t { bool x, y, z; };
inline bvec3_t iszero(const pqvec3_t& v) noexcept { return { v.x == 0, v.y == 0, v.z == 0 }; }
inline bool all(const bvec3_t& b) noexcept { return b.x && b.y && b.z; }
inline bvec3_t operator&&(const bvec3_t& a, const bvec3_t& b) noexcept { return { a.x && b.x, a.y && b.y, a.z && b.z }; }       // glm: componentwise
#endif
}
}

#ifdef WT_SHIM_MM_UNIT
// WT_SHIM_MM_UNIT (oracle/ref_traverse.cpp only): `f_t(1) * u::mm`, the unit the Fraunhofer aperture is expressed in.  Lengths are plain floats in METRES
// here; a length held in millimetres is a type of its own, and the two places the pinned constructor uses it follow mp-units: a length in metres divided
// by one in millimetres is a number carrying the unit ratio m/mm, converted to a plain number by the exact factor 1000 (one f32 product after the
// quotient of the numerical values); a wavenumber in 1/mm times a length in mm is a plain number (no factor).
namespace wt {
struct length_mm_t { f_t mm; };
namespace u { struct mm_unit_t {}; inline constexpr mm_unit_t mm{}; }
constexpr length_mm_t operator*(f_t v, u::mm_unit_t) { return { v }; }
constexpr vec2_t operator/(const pqvec2_t& a, const length_mm_t& b) { return { a.x / b.mm * f_t(1000), a.y / b.mm * f_t(1000) }; }
constexpr f_t operator*(const wavenumber_t& k, const length_mm_t& l) { return k.per_mm * l.mm; }
}
#endif

#ifdef WT_SHIM_MAT4
// WT_SHIM_MAT4 (oracle/ref_mueller.cpp only): glm's vec4 / column-major mat4 and the quantity-vector aliases, as far as
// interaction/polarimetric/{stokes,mueller}.hpp use them.  glm semantics: mat4(s) = s on the diagonal; the 16-scalar constructor fills COLUMN by column;
// m[i] is column i; mat * mat: column c of the result = sum_k a[k] * b[c][k] accumulated left to right (type_mat4x4.inl, no contraction);
// mat4 +, scalar *, / are elementwise.  m::dot on 4-vectors is the fma chain of the reference's vecmath.hpp, like the 2- and 3-vector ones above.
namespace wt {
template <typename T> concept Quantity = std::is_arithmetic_v<T>;
struct vec4q_t {
    f_t x{}, y{}, z{}, w{};
    constexpr vec4q_t() = default;
    constexpr vec4q_t(f_t x_, f_t y_, f_t z_, f_t w_) : x(x_), y(y_), z(z_), w(w_) {}
    constexpr f_t& operator[](std::size_t i) { return i == 0 ? x : i == 1 ? y : i == 2 ? z : w; }
    constexpr const f_t& operator[](std::size_t i) const { return i == 0 ? x : i == 1 ? y : i == 2 ? z : w; }
    constexpr bool operator==(const vec4q_t&) const = default;
};
constexpr vec4q_t operator*(const vec4q_t& a, f_t s) { return { a.x * s, a.y * s, a.z * s, a.w * s }; }
constexpr vec4q_t operator*(f_t s, const vec4q_t& a) { return { s * a.x, s * a.y, s * a.z, s * a.w }; }
constexpr vec4q_t operator/(const vec4q_t& a, f_t s) { return { a.x / s, a.y / s, a.z / s, a.w / s }; }
constexpr vec4q_t operator+(const vec4q_t& a, const vec4q_t& b) { return { a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w }; }
constexpr vec4q_t operator-(const vec4q_t& a, const vec4q_t& b) { return { a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w }; }
template <typename T> concept ScalarOrUnit = std::is_arithmetic_v<T>;
using QE_t = f_t; using QE_solid_angle_t = f_t; using QE_area_t = f_t; using QE_flux_t = f_t; using radiant_flux_t = f_t; using irradiance_t = f_t; using radiant_intensity_t = f_t; using radiance_t = f_t;
using spectral_radiant_flux_t = f_t; using spectral_irradiance_t = f_t; using spectral_radiant_intensity_t = f_t; using spectral_radiance_t = f_t;      // quantities are plain f_t here
template <Quantity Q> using qvec4 = vec4q_t;
template <Quantity Q> using qvec3 = vec3_t;
template <Quantity Q> using qvec2 = vec2_t;
struct mat4_t {
    vec4q_t c[4];
    constexpr mat4_t() = default;
    constexpr mat4_t(f_t s) : c{ { s, 0, 0, 0 }, { 0, s, 0, 0 }, { 0, 0, s, 0 }, { 0, 0, 0, s } } {}
    constexpr mat4_t(f_t a0, f_t a1, f_t a2, f_t a3, f_t b0, f_t b1, f_t b2, f_t b3, f_t c0, f_t c1, f_t c2, f_t c3, f_t d0, f_t d1, f_t d2, f_t d3)
        : c{ { a0, a1, a2, a3 }, { b0, b1, b2, b3 }, { c0, c1, c2, c3 }, { d0, d1, d2, d3 } } {}
    constexpr vec4q_t& operator[](std::size_t i) { return c[i]; }
    constexpr const vec4q_t& operator[](std::size_t i) const { return c[i]; }
};
constexpr mat4_t operator*(const mat4_t& a, f_t s) { mat4_t r; for (int i = 0; i < 4; ++i) r.c[i] = a.c[i] * s; return r; }
constexpr mat4_t operator*(f_t s, const mat4_t& a) { mat4_t r; for (int i = 0; i < 4; ++i) r.c[i] = s * a.c[i]; return r; }
constexpr mat4_t operator/(const mat4_t& a, f_t s) { mat4_t r; for (int i = 0; i < 4; ++i) r.c[i] = a.c[i] / s; return r; }
constexpr mat4_t operator+(const mat4_t& a, const mat4_t& b) { mat4_t r; for (int i = 0; i < 4; ++i) r.c[i] = a.c[i] + b.c[i]; return r; }
constexpr mat4_t operator*(const mat4_t& a, const mat4_t& b) { mat4_t r; for (int i = 0; i < 4; ++i) r.c[i] = a.c[0] * b.c[i].x + a.c[1] * b.c[i].y + a.c[2] * b.c[i].z + a.c[3] * b.c[i].w; return r; }
namespace m {
inline f_t dot(const vec4q_t& a, const vec4q_t& b) noexcept { return std::fma(a.w, b.w, std::fma(a.z, b.z, std::fma(a.y, b.y, a.x * b.x))); }
inline bool isfinite(const vec4q_t& v) noexcept { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z) && std::isfinite(v.w); }
inline bool isnan(const vec4q_t& v) noexcept { return std::isnan(v.x) || std::isnan(v.y) || std::isnan(v.z) || std::isnan(v.w); }
inline bool isfinite(const mat4_t& a) noexcept { return isfinite(a.c[0]) && isfinite(a.c[1]) && isfinite(a.c[2]) && isfinite(a.c[3]); }
inline bool isnan(const mat4_t& a) noexcept { return isnan(a.c[0]) || isnan(a.c[1]) || isnan(a.c[2]) || isnan(a.c[3]); }
inline mat4_t transpose(const mat4_t& a) noexcept { mat4_t r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.c[i][j] = a.c[j][i]; return r; }
}
}
#include <format>
template <> struct std::formatter<wt::mat4_t> : std::formatter<int> { auto format(const wt::mat4_t&, std::format_context& ctx) const { return ctx.out(); } };       // (never called)
#endif
