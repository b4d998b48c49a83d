"""Randomised check of the exact-assignment kernel against scipy.optimize.linear_sum_assignment: sizes 2 .. 700, point clouds
(all cost kinds, clustered / duplicated points) and explicit matrices (continuous, heavily tied integers, negative entries).
Compares the optimal VALUE (permutations may differ on ties).  Usage: python tools/fuzz_auction.py [seconds] [seed]"""
import os, sys, time
import numpy as np
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shwd, oracle
dev = torch.device("cuda:0")
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
g = torch.Generator().manual_seed(seed)
t0 = time.time()
n_cases, worst = 0, 0.0
while time.time() - t0 < budget:
    N = int(rng.choice([2, 3, 5, 17, 33, 64, 100, 257, 400, 700])) if rng.random() < 0.7 else int(rng.integers(2, 300))
    mode = rng.choice(["cloud", "cloud_dup", "cloud_cluster", "dense", "dense_ties", "dense_neg"])
    if mode.startswith("cloud"):
        kind, p = [("sqeuclid", 2), ("geodesic", 2), ("geodesic", 1), ("sqeuclid", 1), ("euclid", 2), ("one_minus_cos", 2), ("sqeuclid", 3)][int(rng.integers(0, 7))]
        x = torch.randn(1, N, 3, generator=g)
        y = torch.randn(1, N, 3, generator=g) * float(rng.uniform(0.2, 2.0)) + float(rng.uniform(-0.5, 0.5))
        if mode == "cloud_dup":
            k = max(1, N // int(rng.integers(2, 9)))
            x = x[:, torch.randint(0, k, (N,), generator=g)]
            y = y[:, torch.randint(0, k, (N,), generator=g)]
        if mode == "cloud_cluster":
            x = x * 1e-3 + torch.randn(1, 1, 3, generator=g)
            y = y * 1e-3 + torch.randn(1, 1, 3, generator=g)
        xn, yn = (F.normalize(x, dim=-1), F.normalize(y, dim=-1)) if kind in ("geodesic", "one_minus_cos") else (x, y)
        sig, _, rounds, status = shwd.ops.exact_assignment(x.to(dev), y.to(dev), kind, float(p), return_info=True)
        C = oracle.cost_matrix(xn, yn, kind, p)[0].double().numpy()
    else:
        if mode == "dense":
            M = torch.rand(N, N, generator=g) * float(10 ** rng.uniform(-3, 3))
        elif mode == "dense_ties":
            M = torch.randint(0, int(rng.integers(2, 6)), (N, N), generator=g).float()
        else:
            M = torch.randn(N, N, generator=g)
        sig, _, rounds, status = shwd.exact_assignment_dense(M.to(dev), return_info=True)
        C = M.double().numpy()
        kind, p = "-", 0
    assert int(status.item()) == 0, (mode, N, kind, p)
    s = sig[0].cpu().numpy()
    assert sorted(s.tolist()) == list(range(N)), (mode, N, kind, p, "not a permutation")
    r, c = linear_sum_assignment(C)
    ours, ref = C[np.arange(N), s].sum(), C[r, c].sum()
    # float32 evaluation of the cost in the kernel vs torch: ~1e-6 relative on the optimum, far below any wrong matching
    err = abs(ours - ref) / max(abs(ref), np.abs(C).max() * 1e-6, 1e-30)
    worst = max(worst, err)
    assert err < 2e-5, (mode, N, kind, p, ours, ref, err)
    n_cases += 1
print("fuzz ok: %d cases in %.0f s, worst relative excess %.2e" % (n_cases, time.time() - t0, worst))
