"""Coefficients and accuracy of the circle coordinate used by csrc/sliced.cu::circle_key (atan q = q P8(q^2), interpolated at
Chebyshev nodes of q^2 in [0, 1]): float32 emulation in numpy against the exact value and against the reference's own float32
chain (F.normalize, atan2, (theta + pi) / (2 pi) -- max_spherical_sliced_w.py:270-279) on the CPU."""
import numpy as np, torch
from numpy.polynomial import chebyshev as C
k = np.arange(4000); sn = 0.5 * (1 - np.cos(np.pi * (k + 0.5) / 4000)); qn = np.sqrt(sn)
c = C.chebfit(2 * sn - 1, np.arctan(qn) / qn, 8)
mono = C.cheb2poly(c)
res = np.zeros(9)
for i, ci in enumerate(mono):
    term = np.array([1.0])
    for _ in range(i):
        term = np.convolve(term, [-1.0, 2.0])
    res[:len(term)] += ci * term
f32 = np.float32
co = [f32(x) for x in res]
print("coefficients s^0..s^8:", [float(x) for x in co])
def fast_key(a, c):
    x, y = (-a).astype(f32), (-c).astype(f32)
    ax, ay = np.abs(x), np.abs(y)
    mx, mn = np.maximum(ax, ay), np.minimum(ax, ay)
    q = (mn / np.where(mx > 0, mx, f32(1))).astype(f32)
    s = (q * q).astype(f32)
    p = np.full_like(s, co[8])
    for i in range(7, -1, -1):
        p = (p * s + co[i]).astype(f32)
    r = (p * q).astype(f32)
    r = np.where(ay > ax, (f32(np.pi / 2) - r).astype(f32), r)
    r = np.where(np.signbit(x), (f32(np.pi) - r).astype(f32), r)
    r = np.copysign(r, y)
    return (r * f32(1 / (2 * np.pi)) + f32(0.5)).astype(f32)
rng = np.random.default_rng(0)
a = rng.standard_normal(2_000_000).astype(f32); c = rng.standard_normal(2_000_000).astype(f32)
exact = (np.arctan2(-c.astype(np.float64), -a.astype(np.float64)) + np.pi) / (2 * np.pi)
fk = fast_key(a, c)
v = torch.nn.functional.normalize(torch.stack([torch.from_numpy(a), torch.from_numpy(c)], -1), p=2, dim=-1)
ref = ((torch.atan2(-v[:, 1], -v[:, 0]) + np.pi) / (2 * np.pi)).numpy()
print("this key   vs exact: max %.3e mean %.3e" % (np.abs(fk - exact).max(), np.abs(fk - exact).mean()))
print("reference  vs exact: max %.3e mean %.3e" % (np.abs(ref - exact).max(), np.abs(ref - exact).mean()))
print("this key vs reference: max %.3e, bit-equal on %.1f %% of the points" % (np.abs(fk - ref).max(), 100 * (fk == ref).mean()))
