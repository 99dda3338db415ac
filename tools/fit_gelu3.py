"""Derives the coefficients of gelu_bf16 (llamole_b200/csrc/llb_common.cuh): degree-3 fit of log2(Phi(-a)) on [0, 7], Lawson-weighted
towards the minimax of the absolute GELU error, checked in emulated fp32 against the fp64 erf form.  Used where the result is
rounded to bf16 right away (GEMM epilogues with bf16 output, GIN messages): max |error| 5.5e-5, i.e. 1/35 of a bf16 half-ulp at 1."""
import numpy as np
from numpy.polynomial import polynomial as Pn
from scipy.special import erf, log_ndtr

A, d = 7.0, 3
a = np.cos(np.linspace(0, np.pi, 4001)) * A / 2 + A / 2
y = log_ndtr(-a) / np.log(2.0)
w = np.ones_like(a)
for _ in range(200):
    coef = Pn.polyfit(a, y, d, w=w)
    err = np.abs(a * 2.0 ** Pn.polyval(a, coef) - a * 2.0 ** y)
    w = w * (1 + 3 * err / err.max())
    w /= w.mean()
c32 = coef.astype(np.float32)
xs = np.linspace(-12, 12, 600001).astype(np.float32)
aa = np.abs(xs)
p = np.full_like(aa, c32[-1])
for k in range(d - 1, -1, -1):
    p = (p * aa + c32[k]).astype(np.float32)
g = (np.maximum(xs, 0) - aa * np.exp2(p.astype(np.float64)).astype(np.float32)).astype(np.float32)
ref = 0.5 * xs.astype(np.float64) * (1 + erf(xs.astype(np.float64) / np.sqrt(2)))
print("coefficients (c0..c3):", [float(c) for c in coef])
print("max |gelu_bf16 - gelu_erf| on [-12, 12]: %.3e" % np.abs(g - ref).max())
