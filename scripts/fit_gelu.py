#!/usr/bin/env python
"""Fits Q(z) ~ log2(erfc(z)) - 1 on [0, 4.2] (degree 7) for the GELU epilogue in csrc/common.cuh::gelu_erf and
reports the worst-case error of x*Phi(x) evaluated in fp32 against scipy's erfc."""
import numpy as np
from scipy.special import erfc

ZMAX, DEG = 4.2, 7
z = np.cos(np.linspace(0, np.pi, 4001)) * 0.5 * ZMAX + 0.5 * ZMAX
y = np.log2(erfc(z)) - 1.0
w = np.ones_like(z)
for _ in range(60):                                   # iteratively re-weighted least squares -> near-minimax
    c = np.polynomial.polynomial.polyfit(z, y, DEG, w=w)
    err = np.polynomial.polynomial.polyval(z, c) - y
    w = w * (1 + 4 * np.abs(err) / np.abs(err).max())
    w /= w.mean()
c32 = c.astype(np.float32)
zz = np.linspace(0, ZMAX, 200001).astype(np.float32)
q = np.zeros_like(zz)
for k in range(DEG, -1, -1):
    q = q * zz + c32[k]
wv = np.exp2(q.astype(np.float64))
x = zz.astype(np.float64) * np.sqrt(2)
ref = 0.5 * erfc(zz.astype(np.float64))
print("coefficients (c0..c7):", ", ".join(f"{v:.9e}f" for v in c32))
print("max |gelu err|:", max(np.abs((x - x * wv) - x * (1 - ref)).max(), np.abs(x * wv - x * ref).max()))
print("max rel err (x>0):", np.abs(((x - x * wv) - x * (1 - ref))[1:] / (x * (1 - ref))[1:]).max())
