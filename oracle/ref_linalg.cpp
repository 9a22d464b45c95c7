// ORACLE -- TEST INFRASTRUCTURE ONLY.  Part of oracle/_ref/libref_gaussian2d.so: the REFERENCE'S OWN 2x2 QR / SVD (include/wt/math/linalg.hpp:24-135,
// including the `n*x*y` denominator at :84), compiled unmodified.  Every beam-footprint transform on the path goes through it: the projection of
// an elliptic cone's cross-section onto a surface and the wavefront's principal axes.  The shims supply the glm
// column-major 2x2 matrix and eft's diff_prod / sum_prod (eft.hpp:117-125, :153-159).
// Pins ot_math.h's QR / SVD: tests/test_oracle_kats.py::test_svd_equals_the_reference_code.
#include <wt/math/common.hpp>
#include <wt/math/linalg.hpp>

extern "C" {
// A: n x 4 floats, glm order (column 0 row 0, column 0 row 1, column 1 row 0, column 1 row 1); out: n x 6 (Ucos Usin Vcos Vsin sigma1 sigma2)
void ref_svd(unsigned n, const float* A, float* out) {
    for (unsigned i = 0; i < n; ++i) {
        const float* a = A + 4 * i; float* o = out + 6 * i;
        const auto s = wt::SVD(wt::mat2_t{ a[0], a[1], a[2], a[3] });
        o[0] = s.Ucos; o[1] = s.Usin; o[2] = s.Vcos; o[3] = s.Vsin; o[4] = s.sigma1; o[5] = s.sigma2;
    }
}
}
