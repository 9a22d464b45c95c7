// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: the debug assertions of include/wt/util/assert.hpp compile to nothing
// (the reference's release build does the same).
#pragma once
#include <cassert>
namespace wt {
template <typename... A> constexpr void assert_iszero(A&&...) noexcept {}
template <typename... A> constexpr void assert_isfinite(A&&...) noexcept {}
template <typename... A> constexpr void assert_unit_vector(A&&...) noexcept {}
}
