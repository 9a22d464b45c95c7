// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: glm is absent here; the vector / matrix types the compiled reference files
// use are the plain ones of the common.hpp shim.
#pragma once
#include <wt/math/common.hpp>
