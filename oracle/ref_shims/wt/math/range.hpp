// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: the closed interval of include/wt/math/range.hpp (a template over mp-units
// quantities, inclusiveness and the wide-vector types), as far as the pinned headers use it: math/intersect/clip.hpp reads the two bounds, the
// 1-D distributions use length / contains / all / & / == (range.hpp:64-69, :141-158) and m::mix(range, t) (:309-314).
#pragma once
#include <limits>
#include <wt/math/common.hpp>
#ifdef WT_SHIM_WIDE_LANES
#include <wt/math/simd/wide_vector.hpp>
#endif
namespace wt {
template <typename T = f_t> struct range_t {
    T min, max;
    constexpr T length() const noexcept { return max - min; }
    constexpr bool empty() const noexcept { return !(min <= max); }
    constexpr T centre() const noexcept { return (max + min) / T(2); }      // range.hpp:175-177
    constexpr range_t grow(const T extent) const noexcept { return range_t{ min - extent, max + extent }; }      // range.hpp:179-181
    constexpr bool overlaps(const range_t& o) const noexcept { return !(*this & o).empty(); }
    constexpr bool contains(T pt) const noexcept { return (pt < max && min < pt) || pt == min || pt == max; }
#ifdef WT_SHIM_WIDE_LANES
    template <std::size_t W> b_w_t<W> contains(const f_w_t<W>& pt) const noexcept { return (f_w_t<W>{ min } <= pt) && (f_w_t<W>{ max } >= pt); }      // range.hpp:89-99, inclusive ends
#endif
    constexpr range_t operator&(const range_t& o) const noexcept { return { m::max(min, o.min), m::min(max, o.max) }; }
    constexpr bool operator==(const range_t& o) const noexcept { return (min == o.min && max == o.max) || (empty() && o.empty()); }
    constexpr bool operator!=(const range_t& o) const noexcept { return !(*this == o); }
    static constexpr range_t null() noexcept { return { +std::numeric_limits<T>::infinity(), -std::numeric_limits<T>::infinity() }; }      // range.hpp:250-256
    static constexpr range_t positive() noexcept { return { T(0), +std::numeric_limits<T>::infinity() }; }
    static constexpr range_t all() noexcept { return { -std::numeric_limits<T>::infinity(), +std::numeric_limits<T>::infinity() }; }
};
template <typename T> constexpr range_t<T> operator*(T s, const range_t<T>& r) noexcept { return s >= 0 ? range_t<T>{ s * r.min, s * r.max } : range_t<T>{ s * r.max, s * r.min }; }
template <typename T = f_t> using pqrange_t = range_t<T>;     // lengths are plain f_t here
namespace m {
template <typename S, typename T> constexpr S mix(const range_t<S>& r, const T& x) noexcept { if (x == T(0)) return r.min; if (x == T(1)) return r.max; return m::mix(r.min, r.max, S(x)); }
}
}
