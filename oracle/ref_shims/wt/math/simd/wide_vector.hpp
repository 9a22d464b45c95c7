// ORACLE -- TEST INFRASTRUCTURE ONLY.
// Build shim for oracle/_ref: stands in for /root/reference/include/wt/math/simd/wide_vector.hpp (AVX 4- / 8- / 16-wide vectors over mp-units
// quantities).  The headers compiled into oracle/_ref only NAME the wide types -- in member templates and in result structs of the wide
// (AVX) entry points, none of which is instantiated here -- so declarations suffice.
#pragma once
#include <cstddef>
#include <wt/math/common.hpp>
namespace wt {
template <std::size_t W> struct f_w_t;
template <std::size_t W> struct b_w_t;
template <std::size_t W> struct length_w_t;
template <std::size_t W> struct vec3_w_t;
template <std::size_t W> struct pqvec3_w_t;
}
