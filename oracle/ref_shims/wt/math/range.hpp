// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: the closed interval of include/wt/math/range.hpp (a template over mp-units
// quantities and the wide-vector types), as far as math/intersect/clip.hpp reads it: two bounds.
#pragma once
#include <wt/math/common.hpp>
namespace wt {
template <typename T = f_t> struct range_t { T min, max; };
template <typename T = f_t> using pqrange_t = range_t<T>;     // lengths are plain f_t here
}
