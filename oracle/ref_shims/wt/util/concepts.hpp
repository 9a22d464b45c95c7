// ORACLE -- TEST INFRASTRUCTURE ONLY.
// Build shim for oracle/_ref: stands in for /root/reference/include/wt/util/concepts.hpp, which pulls in glm and mp-units (absent here,
// SURVEY.md 8c), so that the reference's own sobolld headers compile UNMODIFIED from where they lie.  Only the one concept they use.
#pragma once
#include <type_traits>
namespace wt {
template <typename T> concept FloatingPoint = std::is_floating_point_v<T>;
}
