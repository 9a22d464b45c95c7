// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: libcerf (deps/libcerf @ 09b98c1, an empty submodule here) declares
// `_cerf_cmplx cerfc(_cerf_cmplx)`, _cerf_cmplx = std::complex<double> in C++.  The stand-in forwards to a function the TEST installs
// (scipy.special.erfc on complex128: an independent double-precision implementation, like libcerf's), so that the reference's UTD code is
// exercised with a complementary error function that owes nothing to the oracle's own series (ot_integrator.h cerfc_rot45).
#pragma once
#include <complex>
extern "C" { typedef void (*ref_cerfc_fn)(double re, double im, double* out); extern ref_cerfc_fn ref_cerfc_hook; }
inline std::complex<double> cerfc(const std::complex<double>& z) { double o[2] = { 0, 0 }; ref_cerfc_hook(z.real(), z.imag(), o); return { o[0], o[1] }; }
