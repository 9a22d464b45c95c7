// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: src/math/gaussian2d.cpp includes this glm extension and uses nothing from it.
#pragma once
