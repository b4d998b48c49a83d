"""CPU study for DESIGN.md section 8 item 1 (certified truncation of the OT sweeps): with both clouds Morton-ordered on the
sphere, which fraction of (32-owner group, 16-point streamed sub-tile) pairs could a warp skip in a converged half-step,
using only the bounding cones, the tile's maximum potential and the group's minimum log-sum-exp?  float64, one pair."""
import math, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

def morton(p):
    q = np.clip(((p + 1) * 0.5 * 1023).astype(np.int64), 0, 1023)
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF; v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3; v = (v | (v << 2)) & 0x09249249
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)

def cones(p, size):
    g = p.reshape(-1, size, 3)
    c = g.mean(1); c /= np.linalg.norm(c, axis=1, keepdims=True)
    cosr = np.einsum("gsk,gk->gs", g, c).min(1)
    return c, np.arccos(np.clip(cosr, -1, 1))

for N in (1024, 4096):
    L, eps, T = 100, 0.01, 50.0
    t, s = bench.registration_pairs(1, N, 1234)
    x = t[0] - t[0].mean(0); y = s[0] - s[0].mean(0)
    x = (x / x.norm(dim=1, keepdim=True)).double().numpy(); y = (y / y.norm(dim=1, keepdim=True)).double().numpy()
    x = x[np.argsort(morton(x))]; y = y[np.argsort(morton(y))]
    C = np.arccos(np.clip(x @ y.T, -1, 1)) ** 2
    k = math.log2(math.e) / eps
    la = math.log2(1.0 / N + 1e-8)
    a = np.zeros(N); b = np.zeros(N)
    def lse2(M, axis):
        m = M.max(axis, keepdims=True)
        return (m + np.log2(np.exp2(M - m).sum(axis, keepdims=True))).squeeze(axis)
    for it in range(L):
        lse_a = lse2(b[None, :] - k * C, 1); a = la - lse_a
        lse_b = lse2(a[:, None] - k * C, 0); b = la - lse_b
    # row half-step at convergence: owner i streams j, exponent m_ij = b_j - kC_ij - lse_a_i  (sum_j 2^m = 1)
    m = b[None, :] - k * C - lse_a[:, None]
    need = (m > -T)
    cg, rg = cones(x, 32); ct, rt = cones(y, 16)
    d = np.arccos(np.clip(cg @ ct.T, -1, 1))
    th_min = np.maximum(d - rg[:, None] - rt[None, :], 0.0)
    bound = b.reshape(-1, 16).max(1)[None, :] - k * 2 * (1 - np.cos(th_min)) - lse_a.reshape(-1, 32).min(1)[:, None]
    skip = bound < -T
    # ground truth: tiles with no element above the threshold
    truth = ~need.reshape(N // 32, 32, N // 16, 16).any(axis=(1, 3))
    assert not (skip & ~truth).any(), "the bound must be conservative"
    # the packed layout pairs sub-tile t with sub-tile t + (N/16)/2: a record block is skipped only if both are
    half = skip.shape[1] // 2
    both = skip[:, :half] & skip[:, half:]
    per_group = skip.mean(1)
    print("N=%5d  elements above 2^-%d: %.1f %%   tiles skippable (truth) %.1f %%   by the cone bound %.1f %%   "
          "both halves of a packed block %.1f %%   slowest group skips %.1f %%   mean patch radius: group %.1f deg, tile %.1f deg"
          % (N, int(T), 100 * need.mean(), 100 * truth.mean(), 100 * skip.mean(), 100 * both.mean(), 100 * per_group.min(),
             np.degrees(rg.mean()), np.degrees(rt.mean())))
