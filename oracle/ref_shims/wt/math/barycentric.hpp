// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: barycentric_t (include/wt/math/barycentric.hpp, which pulls in the mesh
// headers) as sampler.hpp's uniform_triangle returns it: the two free coordinates.
#pragma once
#include <wt/math/common.hpp>
namespace wt { struct barycentric_t { vec2_t uv; constexpr explicit barycentric_t(vec2_t v) : uv(v) {} }; }
