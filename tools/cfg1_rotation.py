"""cfg1 (BASELINE.json configs[0]): the loss comparison of Comparison_Wasserstein_with_Chamfer_distance/main_rotation.py
-- Chamfer (:203), log-Sinkhorn 'L2' (:207-211) and the exact W2 through the script's own POT_loss (:63-79, ot.emd2) --
on synthetic sphere clouds N = 1024 rotated about x by 90 ... 180 degrees (Data_set_transformation.py:159), evaluated
with the drop-ins (losses.chamfer_distance / log_Sinkhorn_Distance_Loss / dropin_ot.ot.emd2) on the GPU; Chamfer and the
exact W2 are re-evaluated on a bounded CPU sample (plain torch broadcast; scipy's assignment solver for the LP) for a
sanity column and the CPU timing.  The reference's per-sample means (CD_loss / len(test_dataset), ...) are tabulated."""
import math
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dropin_ot"))
import ot  # noqa: E402  (the drop-in)
import shwd  # noqa: E402

dev = torch.device("cuda:0")
B, N, EPS, ITERS = 32, 1024, 0.01, 100  # main_rotation.py:116-124 defaults
CPU_PAIRS = 2                            # bounded CPU sample per angle


def cost_matrix(x, y, p=2):  # main_rotation.py:82-92 (restated)
    return torch.pow(torch.sum((torch.abs(x.unsqueeze(-2) - y.unsqueeze(-3))) ** p, -1), 1.0 / p)


def pot_loss(x, y, criteria, p=2):  # main_rotation.py:63-79 (restated): sum_b emd2_b^(1/p)
    C = cost_matrix(x, y, p)
    n = x.shape[-2]
    w = torch.full((n,), 1.0 / n, device=x.device)
    return sum(torch.pow(criteria(w, w, C[i]), 1.0 / p) for i in range(x.shape[0]))


def rot_x(deg):
    a = math.radians(deg)
    return torch.tensor([[1.0, 0.0, 0.0], [0.0, math.cos(a), -math.sin(a)], [0.0, math.sin(a), math.cos(a)]])


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v = fn()
    torch.cuda.synchronize()
    return v, (time.perf_counter() - t0) * 1e3


torch.manual_seed(1234)
tmpl = F.normalize(torch.randn(B, N, 3), dim=-1)
sd = shwd.losses.log_Sinkhorn_Distance_Loss(eps=EPS, max_iter=ITERS, batch_reduction="sum", type_of_cost_norm="L2")
print("| angle | CD / pair | Sinkhorn / pair | W2 (exact) / pair | GPU ms: CD, SD, WD (32 pairs) | CPU ms per pair: CD, WD (scipy LP) | max rel. dev. GPU vs CPU sample (CD, WD) |")
print("|---|---|---|---|---|---|---|")
for warm in (True, False):
    for deg in ((90,) if warm else (90, 105, 120, 135, 150, 165, 180)):
        src = tmpl @ rot_x(deg).T
        t, s = tmpl.to(dev), src.to(dev)
        with torch.no_grad():
            cd, t_cd = timed(lambda: shwd.losses.chamfer_distance(s, t, batch_reduction="sum")[0].item())
            sk, t_sd = timed(lambda: sd(t, s, dev)[0].item())
            wd, t_wd = timed(lambda: pot_loss(t, s, ot.emd2).item())
        if warm:
            continue
        # CPU oracle on a bounded sample (same pairs), for parity and the timing column
        tc, sc = tmpl[:CPU_PAIRS], src[:CPU_PAIRS]
        c0 = time.perf_counter()
        d = ((sc.unsqueeze(2) - tc.unsqueeze(1)) ** 2).sum(-1)
        cd_c = (d.min(2).values.mean(1) + d.min(1).values.mean(1)).sum().item()
        c1 = c2 = time.perf_counter()
        from scipy.optimize import linear_sum_assignment
        wd_c = 0.0
        for i in range(CPU_PAIRS):
            C = cost_matrix(tc[i], sc[i]).double().numpy()
            r, c = linear_sum_assignment(C)
            wd_c += math.sqrt(C[r, c].sum() / N)
        c3 = time.perf_counter()
        with torch.no_grad():
            g = (shwd.losses.chamfer_distance(s[:CPU_PAIRS], t[:CPU_PAIRS], batch_reduction="sum")[0].item(),
                 pot_loss(t[:CPU_PAIRS], s[:CPU_PAIRS], ot.emd2).item())
        dev_rel = max(abs(a - b) / abs(b) for a, b in zip(g, (cd_c, wd_c)))
        print("| %d | %.5f | %.5f | %.5f | %.2f, %.2f, %.2f | %.1f, %.1f | %.1e |" % (
            deg, cd / B, sk / B, wd / B, t_cd, t_sd, t_wd, (c1 - c0) / CPU_PAIRS * 1e3, (c3 - c2) / CPU_PAIRS * 1e3, dev_rel))
