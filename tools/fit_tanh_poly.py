"""Coefficients of the odd polynomial used by spa_tanh_half_f32 (csrc/step_kernels.cuh) for |h| < 0.6:
tanh(h) ~ h + h * u * p(u), u = h^2, p of degree 3, fitted by iteratively re-weighted least squares on Chebyshev nodes
(a cheap stand-in for Remez) and checked in float32 arithmetic against numpy's float64 tanh."""
import numpy as np


def f(u):
    h = np.sqrt(u)
    return np.where(u > 1e-12, (np.tanh(h) / h - 1) / u, -1 / 3)


def main():
    a, b = 0.0, 0.36
    k = np.arange(400)
    u = (np.cos(np.pi * (k + 0.5) / 400) + 1) / 2 * (b - a) + a
    w = u.copy()
    coef = np.polyfit(u, f(u), 3, w=w + 1e-3)
    for _ in range(30):
        e = np.abs(u * (np.polyval(coef, u) - f(u)))
        w = w * (1 + e / e.max())
        coef = np.polyfit(u, f(u), 3, w=w)
    c32 = coef.astype(np.float32)
    h = np.linspace(1e-4, 0.6, 200001).astype(np.float32)
    uu = (h * h).astype(np.float32)
    p = np.float32(c32[0])
    for c in c32[1:]:
        p = (p * uu + np.float32(c)).astype(np.float32)
    t = (h * (p * uu).astype(np.float32) + h).astype(np.float32)
    ref = np.tanh(h.astype(np.float64))
    print("coefficients (highest power first):", [float(c) for c in c32])
    print("max relative error: %.3g (%.2f ulp of float32)" % (np.max(np.abs(t - ref) / ref), np.max(np.abs(t - ref) / ref) / 2 ** -24))


if __name__ == "__main__":
    main()
