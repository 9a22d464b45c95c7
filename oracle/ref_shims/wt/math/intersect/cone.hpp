// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: reference sources that include <wt/math/intersect/cone.hpp> get the part of that
// header the Makefile's `ref` target cut out at build time (oracle/_ref/cone_scalar_part.hpp: everything but the 8-wide cone-AABB test, :272-475).
#pragma once
#include <_ref/cone_scalar_part.hpp>
