"""Fit Q(a) ~= acos(a)/sqrt(1-a) on a in [0,1] (relative minimax via Remez-like iterated weighted LSQ).
acos(c) = sqrt(1-|c|) * Q(|c|) for c>=0, pi - that for c<0.  Prints fp32 coefficients and error measured
with fp32 Horner evaluation against float64 acos."""
import numpy as np
from numpy.polynomial import chebyshev as Ch

def target(a):
    a = np.asarray(a, dtype=np.float64)
    w = 1.0 - a
    out = np.empty_like(a)
    small = w < 1e-9
    out[~small] = np.arccos(a[~small]) / np.sqrt(w[~small])
    # series: acos(a)/sqrt(1-a) -> sqrt(2)*(1 + w/12 + 3w^2/160 ...)
    ws = w[small]
    out[small] = np.sqrt(2.0) * (1 + ws / 12 + 3 * ws**2 / 160)
    return out

def fit(deg, iters=30):
    n = 4000
    xs = 0.5 * (1 - np.cos(np.pi * (np.arange(n) + 0.5) / n))  # cheb nodes on [0,1]
    f = target(xs)
    wts = np.ones(n)
    best = None
    for _ in range(iters):
        V = np.vander(xs, deg + 1, increasing=True)
        A = V * (wts / f)[:, None]
        b = wts
        coef, *_ = np.linalg.lstsq(A, b, rcond=None)
        err = np.abs(V @ coef / f - 1)
        m = err.max()
        if best is None or m < best[0]:
            best = (m, coef.copy())
        wts = wts * (1 + 2.0 * err / (m + 1e-300))
        wts /= wts.mean()
    return best

for deg in range(3, 9):
    m, coef = fit(deg)
    c32 = coef.astype(np.float32)
    # fp32 evaluation of full acos over c in [-1,1]
    c = np.linspace(-1, 1, 2000001).astype(np.float32)
    a = np.abs(c)
    q = np.full_like(a, c32[-1])
    for k in range(deg - 1, -1, -1):
        q = (q * a + c32[k]).astype(np.float32)
    sq = np.sqrt((np.float32(1) - a).astype(np.float32)).astype(np.float32)
    h = (np.float32(np.pi / 2) - sq * q).astype(np.float32)
    th = (np.float32(np.pi / 2) - np.copysign(h, c)).astype(np.float32)
    ref = np.arccos(c.astype(np.float64))
    abs_err = np.abs(th - ref)
    rel = abs_err[ref > 1e-3] / ref[ref > 1e-3]
    print(f"deg {deg}: minimax rel(Q) {m:.3e}  fp32 acos abs err max {abs_err.max():.3e}  rel max(theta>1e-3) {rel.max():.3e}")
    if deg in (5, 6, 7):
        print("   coef:", ", ".join(f"{x:.9e}f" for x in c32))
