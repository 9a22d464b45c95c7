"""Regenerates the Fraunhofer free-space-diffraction importance-sampling tables.

The reference loads data/fsd/iCDFa{1,2}.fp64 (3072x3072 f64) and iCDFa{1,2}theta.fp64 (2048 f64) (src/interaction/fsd/fraunhofer/fsd_lut.cpp:28-74);
in the reference tree these are Git-LFS pointer stubs, and their construction is not in the repository.  What the sampler needs is fixed by how
the tables are read (include/wt/interaction/fsd/fraunhofer/fsd_lut.hpp:54-69): with zeta = r (cos t, sin t) in the first quadrant,
    t = iCDFtheta(u1),  r = iCDF[t][u2],  then a random quadrant,
i.e. inverse CDFs of the densities  p_j(zeta) ~ chi_e(zeta) |alpha_j(zeta)|^2  (j = 1, 2; include/wt/interaction/fsd/fraunhofer/fsd.hpp:59-84).
integrate() evaluates the normalisations of these densities: 0.004827 and 0.16252 (polar and Cartesian quadratures agree to 1e-3).
The reference quotes PA1 = 0.0049361, PA2 = 0.21899 (fsd.hpp:55-57) for "power contained in chi_e x |alpha_j|^2": +2% / +35% off what
its published formulas integrate to, so the original tables used a mask or cut-off that is not in the repository.  The constants are
used verbatim where the reference uses them (edge selection pdfs); the tables only shape the rejection-sampling proposal, whose pdf
is never assumed exact (fsd_sampler.cpp:81-113) -- oracle and device share one set of tables, so parity is unaffected.
"""
import os
import numpy as np

_CACHE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cache")
R_MAX = 2.0e3


def _sinc(x):
    return np.sinc(x / np.pi)


def alpha1(x, y):
    with np.errstate(divide="ignore", invalid="ignore"):
        v = (1 / (2 * np.pi)) * y / (x * (x * x + y * y)) * (np.cos(x / 2) - _sinc(x / 2))
    return np.where(x == 0, 0.0, v)


def alpha2(x, y):
    with np.errstate(divide="ignore", invalid="ignore"):
        v = (1 / (2 * np.pi)) * y / (x * x + y * y) * _sinc(x / 2)
    return np.where(x == 0, 0.0, v)


def chi_e(r2):
    t = 1 + 0.830092714835359 * r2
    return np.maximum(0.0, 1 - (3 / t ** 2 - 2 / t ** 3))


def _r_grid():
    return np.unique(np.concatenate([np.linspace(0.0, 60.0, 24001), np.geomspace(60.0, R_MAX, 12000)]))


def _polar_density(alpha, thetas, r):
    th = thetas[:, None]; rr = r[None, :]
    a = alpha(rr * np.cos(th), rr * np.sin(th))
    return chi_e(rr * rr) * a * a * rr          # Jacobian r


def integrate(which, n_theta=1025):
    """4 * int_0^{pi/2} int_0^R chi_e |alpha_j|^2 r dr dtheta  ->  PA_j."""
    alpha = alpha1 if which == 1 else alpha2
    r = _r_grid(); th = np.linspace(0, np.pi / 2, n_theta)
    q = _polar_density(alpha, th, r)
    return 4 * np.trapezoid(np.trapezoid(q, r, axis=1), th)


def build(n=2048, m=1024, use_cache=True):
    """Returns (icdf_theta1[n], icdf_theta2[n], icdf1[m,m], icdf2[m,m]) as float32."""
    path = os.path.join(_CACHE, f"fsd_lut_{n}_{m}.npz")
    if use_cache and os.path.exists(path):
        z = np.load(path); return z["t1"], z["t2"], z["c1"], z["c2"]
    r = _r_grid()
    out = {}
    for j, alpha in ((1, alpha1), (2, alpha2)):
        th_rows = np.linspace(0, np.pi / 2, m)
        icdf = np.zeros((m, m), np.float32)
        marg = np.zeros(m)
        u = np.linspace(0, 1, m)
        for lo in range(0, m, 128):
            q = _polar_density(alpha, th_rows[lo:lo + 128], r)
            cdf = np.concatenate([np.zeros((q.shape[0], 1)), np.cumsum(.5 * (q[:, 1:] + q[:, :-1]) * np.diff(r)[None, :], axis=1)], axis=1)
            marg[lo:lo + 128] = cdf[:, -1]
            for i in range(q.shape[0]):
                tot = cdf[i, -1]
                icdf[lo + i] = np.interp(u, cdf[i] / tot, r) if tot > 0 else 0.0
        # marginal over theta (finer grid for the inverse)
        cth = np.concatenate([[0], np.cumsum(.5 * (marg[1:] + marg[:-1]) * np.diff(th_rows))])
        icdft = np.interp(np.linspace(0, 1, n), cth / cth[-1], th_rows).astype(np.float32)
        out[f"t{j}"], out[f"c{j}"] = icdft, icdf
    if use_cache:
        os.makedirs(_CACHE, exist_ok=True)
        np.savez(path, **out)
    return out["t1"], out["t2"], out["c1"], out["c2"]
