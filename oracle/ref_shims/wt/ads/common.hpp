// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: the one name of include/wt/ads/common.hpp that interaction/fsd/common.hpp uses.
// oracle/ref_traverse.cpp (WT_SHIM_WIDE_LANES) compiles the traversal loops and takes the reference's own header instead.
#pragma once
#include <cstdint>
#ifdef WT_SHIM_WIDE_LANES
#include "/root/reference/include/wt/ads/common.hpp"
#else
namespace wt::ads { using tuid_t = std::uint32_t; }
#endif
