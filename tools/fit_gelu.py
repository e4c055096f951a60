"""Fit of the epilogue GELU: gelu(x) = x * sigmoid(x * (a0 + a1 x^2 + a2 x^4)), x^2 clamped at L^2, against the exact
x * Phi(x); prints the coefficients used by gemm_tcgen05.cu:gelu_fast and the maximum absolute error (float32 evaluation)."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf

L = 5.0
xs = np.linspace(-8, 8, 400001)
g = xs * 0.5 * (1 + erf(xs / np.sqrt(2)))

def approx(c, x, dt=np.float64):
    x = x.astype(dt); c = c.astype(dt)
    x2 = np.minimum(x * x, dt(L * L))
    q = (c[2] * x2 + c[1]) * x2 + c[0]
    return x / (dt(1) + np.exp(-(q * x)))

c = np.array([1.5958, 0.0714, 0.0])
for p in (2, 4, 8, 16):
    c = least_squares(lambda c: np.sign(approx(c, xs) - g) * np.abs(approx(c, xs) - g) ** (p / 2), c, xtol=1e-15, ftol=1e-15).x
print("coefficients", c)
print("max abs error (float64 eval)", np.abs(approx(c, xs) - g).max())
print("max abs error (float32 eval)", np.abs(approx(c, xs, np.float32).astype(np.float64) - g).max())
