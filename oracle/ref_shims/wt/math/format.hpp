// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: include/wt/math/format.hpp (std::format support for vectors and quantities:
// printing, no arithmetic) is not needed by the pinned code.
#pragma once
