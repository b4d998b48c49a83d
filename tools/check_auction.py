"""Diagnostic: exact assignment (auction kernel) vs scipy's linear_sum_assignment on the reference's cost matrix."""
import os, sys, time
import numpy as np
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd, oracle, bench
dev = torch.device("cuda:0")
for (B, N, kind, p, seed) in [(3, 7, "sqeuclid", 2, 0), (2, 64, "geodesic", 2, 1), (2, 257, "sqeuclid", 2, 2), (2, 300, "geodesic", 1, 3),
                              (1, 1024, "geodesic", 2, 4), (1, 1024, "sqeuclid", 2, 5), (2, 1, "sqeuclid", 2, 6), (1, 2048, "sqeuclid", 2, 7)]:
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, N, 3, generator=g); y = torch.randn(B, N, 3, generator=g) * 0.9 + 0.1
    if kind == "geodesic":
        x, y = F.normalize(x, dim=-1), F.normalize(y, dim=-1)
    torch.cuda.synchronize(); t0 = time.time()
    sig, prices, rounds, status = shwd.ops.exact_assignment(x.to(dev), y.to(dev), kind, float(p), return_info=True)
    torch.cuda.synchronize(); t1 = time.time()
    C = oracle.cost_matrix(x, y, kind, p).double().numpy()
    ok, worst = True, 0.0
    for b in range(B):
        r, c = linear_sum_assignment(C[b])
        ref = C[b][r, c].sum()
        s = sig[b].cpu().numpy()
        assert sorted(s.tolist()) == list(range(N)), "not a permutation"
        ours = C[b][np.arange(N), s].sum()
        worst = max(worst, (ours - ref) / max(ref, 1e-30))
        ok = ok and np.array_equal(s, c)
    print("B%d N%d %s p%s: same permutation %s  rel. excess cost %.2e  rounds %s  status %d  [%.2f ms]" % (
        B, N, kind, p, ok, worst, rounds.cpu().tolist(), int(status.item()), (t1 - t0) * 1e3))
# training-shaped batch
tmpl, src = bench.registration_pairs(32, 1024, 1234, dev)
tmpl = tmpl - tmpl.mean(1, keepdim=True); src = src - src.mean(1, keepdim=True)
for kind in ("sqeuclid", "geodesic"):
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        sig, prices, rounds, status = shwd.ops.exact_assignment(tmpl, src, kind, 2.0, return_info=True)
        torch.cuda.synchronize(); t1 = time.time()
    print("B=32 N=1024 %s: %.1f ms, rounds min/max %d/%d, status %d" % (kind, (t1 - t0) * 1e3, rounds.min().item(), rounds.max().item(), int(status.item())))
xg = src.clone().requires_grad_(True)
crit = shwd.losses.Cos_disimilarity_W(dev, p=2, solver="exact")
torch.cuda.synchronize(); t0 = time.time()
loss = crit(tmpl, xg); loss.backward()
torch.cuda.synchronize(); t1 = time.time()
print("Cos_disimilarity_W(solver=exact) loss %.6f fwd+bwd %.1f ms; entropic (eps=0.01, L=100): %.6f" % (
    loss.item(), (t1 - t0) * 1e3, shwd.losses.Cos_disimilarity_W(dev, p=2, solver="sinkhorn")(tmpl, src).item()))
