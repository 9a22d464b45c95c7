// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: the one name of include/wt/ads/common.hpp that interaction/fsd/common.hpp uses.
#pragma once
#include <cstdint>
namespace wt::ads { using tuid_t = std::uint32_t; }
