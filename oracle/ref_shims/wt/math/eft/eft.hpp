// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: include/wt/math/eft/eft.hpp is a template library over mp-units quantities and
// the wide-vector types; the two scalar functions the pinned headers call -- diff_prod (:117-125) and sum_prod (:153-159) -- live in the
// wt/math/common.hpp shim (m::eft).
#pragma once
#include <wt/math/common.hpp>
