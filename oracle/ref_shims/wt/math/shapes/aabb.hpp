// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: src/math/gaussian2d.cpp includes the AABB header and uses nothing from it.
#pragma once
