// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: include/wt/math/type_traits.hpp (concepts over mp-units quantity vectors).
// math/util.hpp constrains its length-vector overloads with QuantityVectorOf<isq::length>; here that is "one of the shim's length-vector types"
// (only with WT_SHIM_DISTINCT_PQ, where those are types of their own).
#pragma once
#include <concepts>
#include <cstddef>
#include <wt/math/common.hpp>
namespace wt {
namespace isq { struct length_tag {}; inline constexpr length_tag length{}; }
#ifdef WT_SHIM_DISTINCT_PQ
template <typename T, auto Q> concept QuantityVectorOfImpl = std::same_as<std::remove_cvref_t<T>, pqvec2_t> || std::same_as<std::remove_cvref_t<T>, pqvec3_t>;
template <typename T> inline constexpr std::size_t element_count_v = std::same_as<std::remove_cvref_t<T>, pqvec2_t> ? 2 : 3;
namespace u { constexpr vec3_t to_m(const pqvec3_t& v) { return { v.x, v.y, v.z }; } constexpr vec2_t to_m(const pqvec2_t& v) { return { v.x, v.y }; } }
#else
template <typename T, auto Q> concept QuantityVectorOfImpl = false;
template <typename T> inline constexpr std::size_t element_count_v = 0;
#endif
template <auto Q> struct qv_of { template <typename T> static constexpr bool value = QuantityVectorOfImpl<T, Q>; };
}
// `template <QuantityVectorOf<isq::length> V>` needs a concept whose first parameter is the constrained type
namespace wt { template <typename T, auto Q> concept QuantityVectorOf = QuantityVectorOfImpl<T, Q>; }
