"""Eager call vs CUDA-graph replay (shwd.graphed_loss) of loss forward + backward at the launch-bound shapes.
Wall clock per call with the host free-running (what a training loop sees), synchronised at both ends."""
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
U, _ = torch.linalg.qr(torch.randn(512, 3, 2, generator=g))
U = U.to(dev)
th = F.normalize(torch.randn(512, 3, generator=g), dim=-1).to(dev)
geo = shwd.losses.Geodesic_distance_W(device=dev, p=2, eps=0.01, max_iter=100)
CASES = [
    ("Chamfer B=32 N=1024 (train_CD.py:161)", lambda a, b: shwd.losses.chamfer_distance(a, b)[0], (32, 1024, 3)),
    ("Chamfer B=1 N=16384 (flow notebooks)", lambda a, b: shwd.losses.chamfer_distance(a, b)[0], (1, 16384, 3)),
    ("SSW p=1 P=512, one pair N=4096 (cfg3)", lambda a, b: shwd.losses.sliced_cost(a, b, U, p=1), (4096, 3)),
    ("SSW p=2 P=512, one pair N=4096 (cfg3)", lambda a, b: shwd.losses.sliced_cost(a, b, U, p=2), (4096, 3)),
    ("Euclid SW p=2 P=512, one pair N=4096", lambda a, b: shwd.losses.sliced_wasserstein_distance(a, b, p=2, device=dev, projections=th), (4096, 3)),
    ("W_COS geodesic L=100, one pair N=1024 (cfg1)", lambda a, b: geo(a, b), (1, 1024, 3)),
    ("W_COS geodesic L=100, B=32 N=1024 (cfg2)", lambda a, b: geo(a, b), (32, 1024, 3)),
]


def wall(fn, x, y, reps):
    for _ in range(5):
        fn(x, y)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn(x, y)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print("| loss call (forward + backward) | eager ms | graphed (autograd node) ms | graphed value_and_grad ms |")
print("|---|---|---|---|")
for name, fn, shape in CASES:
    x = F.normalize(torch.randn(*shape, generator=g), dim=-1).to(dev)
    y = F.normalize(torch.randn(*shape, generator=g) + 0.2, dim=-1).to(dev)
    reps = 20 if "cfg2" in name else 200

    def eager(a, b):
        a = a.detach().requires_grad_(True)
        b = b.detach().requires_grad_(True)
        fn(a, b).backward()

    gfn = shwd.graphed_loss(fn)

    def graphed(a, b):
        a = a.detach().requires_grad_(True)
        b = b.detach().requires_grad_(True)
        gfn(a, b).backward()

    xr, yr = x.clone().requires_grad_(True), y.clone().requires_grad_(True)

    def vg(a, b):
        gfn.value_and_grad(xr, yr)

    print("| %s | %.3f | %.3f | %.3f |" % (name, wall(eager, x, y, reps), wall(graphed, x, y, reps), wall(vg, x, y, reps)))
