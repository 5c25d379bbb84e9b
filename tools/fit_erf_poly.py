"""Coefficients of the FMA-only erf used by the fused-GELU GEMM epilogues (csrc/common.cuh: erf_poly2).
erf(z) = z * P(t), t = 2 z^2 / zmax^2 - 1; weighted least squares iterated towards the minimax of the absolute error; the
error is then measured with the fp32 Horner evaluation the kernel performs."""
import numpy as np
from numpy.polynomial import chebyshev as C
from scipy.special import erf


def fit(zmax, nterms, iters=40):
    U = zmax ** 2
    k = np.arange(8000)
    u = 0.5 * U * (1 - np.cos(np.pi * (k + 0.5) / 8000))
    z = np.sqrt(u)
    g = np.where(z > 1e-8, erf(z) / np.maximum(z, 1e-30), 2 / np.sqrt(np.pi))
    V = C.chebvander(2 * u / U - 1, nterms - 1)
    w = z.copy()
    for _ in range(iters):
        coef, *_ = np.linalg.lstsq(V * w[:, None], g * w, rcond=None)
        err = np.abs(z * (V @ coef - g))
        w = w * (1 + 2 * err / err.max())
    return C.cheb2poly(coef)


if __name__ == "__main__":
    zmax, n = 3.7, 13
    m = fit(zmax, n)
    print("zmax", zmax, "coefficients (t^0 .. t^%d):" % (n - 1))
    print(", ".join("%.9ef" % c for c in m))
    x = np.linspace(-8, 8, 800001).astype(np.float32)
    z = np.clip(x * np.float32(0.70710678118654752), -np.float32(zmax), np.float32(zmax))
    t = (z * z) * np.float32(2.0 / zmax ** 2) - np.float32(1.0)
    acc = np.full_like(t, np.float32(m[-1]))
    for c in m[-2::-1]:
        acc = acc * t + np.float32(c)
    e = acc * z
    x64 = x.astype(np.float64)
    print("max |erf error| (fp32 Horner): %.3e" % np.abs(e - erf(x64 / np.sqrt(2))).max())
    print("max |gelu error|: %.3e" % np.abs(0.5 * x * (1 + e) - 0.5 * x64 * (1 + erf(x64 / np.sqrt(2)))).max())
