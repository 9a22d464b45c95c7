// ORACLE -- TEST INFRASTRUCTURE ONLY.
// Build shim for oracle/_ref: stands in for /root/reference/include/wt/math/common.hpp (glm / mp-units) with the handful of names the
// reference's sobolld_sampler.hpp uses: f_t (the f32 build, CMakeLists.txt:214-216), limits<>, m::pow / ceil / log / min.
#pragma once
#include <cmath>
#include <cstddef>
#include <limits>
#include <algorithm>
namespace wt {
using f_t = float;
template <typename T> using limits = std::numeric_limits<T>;
namespace m {
template <typename T> constexpr T pow(T base, std::size_t e) noexcept { T r = 1; for (std::size_t i = 0; i < e; ++i) r *= base; return r; }
using std::ceil; using std::log;
template <typename T> constexpr T min(T a, T b) noexcept { return std::min(a, b); }
}
}
