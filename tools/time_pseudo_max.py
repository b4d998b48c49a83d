"""pseudo_max_cos_disimilarity_wassersten_distance (s2_wasserstein.py:272-344) on one B200: ten random Residual flows, B = 32,
N = 1024; one solver launch for all 320 pairs against one criterion call per flow (the reference's order).
    gpurun -- python tools/time_pseudo_max.py > gpurun_out/<tag>_pseudo_max.md
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import shwd  # noqa: E402

L = shwd.losses


def wall(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    dev = torch.device("cuda:0")
    a, b = bench.registration_pairs(32, 1024, 1234, dev)
    a = a - a.mean(1, keepdim=True)
    b = b - b.mean(1, keepdim=True)
    print("| criterion | mode | one call per flow ms | one launch ms | value (per flow / one launch) |")
    print("|---|---|---|---|---|")
    for name, csw in (("Cos_disimilarity_W(p=2): exact ot.emd2", L.Cos_disimilarity_W(dev, p=2)),
                      ("Geodesic_distance_W(p=2, eps=0.01, L=100): entropic", L.Geodesic_distance_W(dev, p=2, eps=0.01, max_iter=100))):
        torch.manual_seed(0)
        pm = L.pseudo_max_cos_disimilarity_wassersten_distance(csw, dev, phi_num=10, n_flow_layer=5, flow_name="Residual")
        for mode in ("max", "mean"):
            pm.mean_or_max_or_softmax = mode
            pm.batched = False
            t0, v0 = wall(lambda: pm(a, b)[0].item()), pm(a, b)[0].item()
            pm.batched = True
            t1, v1 = wall(lambda: pm(a, b)[0].item()), pm(a, b)[0].item()
            print("| %s | %s | %.1f | %.1f | %.6f / %.6f |" % (name, mode, t0, t1, v0, v1))


if __name__ == "__main__":
    main()
