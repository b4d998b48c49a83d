"""Loss-kernel sweep (BASELINE.json config 5) plus the cfg3 / cfg4 shapes: fwd+bwd time of the W_COS loss (geodesic
entropic OT, eps=0.01, L=100) and of Chamfer over N x B, and of the sliced losses at N=4096, P=512.  Prints a markdown
table; run on the GPU box:  python tools/sweep.py > gpurun_out/sweep.md"""
import os, sys, time
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd
dev = torch.device("cuda:0")
L, EPS = 100, 0.01


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def clouds(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    x = F.normalize(torch.randn(B, N, 3, generator=g), dim=-1)
    y = F.normalize(torch.randn(B, N, 3, generator=g) + 0.2, dim=-1)
    return x.to(dev).requires_grad_(True), y.to(dev).requires_grad_(True)


print("| loss | B | N | fwd+bwd ms | pairs/s | FP32-roofline frac (54 lane-ops x (2L+1) N^2) |")
print("|---|---|---|---|---|---|")
peak = 36.6e12
for N in (256, 1024, 4096, 16384, 65536):
    for B in (1, 32, 256):
        work = (2 * L + 1) * N * N * B
        if work > 2.0e12:
            continue
        x, y = clouds(B, N, N + B)

        def step():
            x.grad = None; y.grad = None
            res = shwd.entropic_ot(x, y, "geodesic", 2.0, EPS, L, center=True)
            res.cost.sum().backward()
        ms = timed(step, 3 if work > 2e11 else 6)
        print("| W_COS (geodesic OT, L=100) | %d | %d | %.3f | %.1f | %.3f |" % (B, N, ms, B / (ms * 1e-3), 54.0 * work / (ms * 1e-3) / peak))
        sys.stdout.flush()
for N in (256, 1024, 4096, 16384, 65536):
    for B in (1, 32, 256):
        if B * N * N > 3e11:
            continue
        x, y = clouds(B, N, N + B)

        def step():
            x.grad = None; y.grad = None
            out, _ = shwd.losses.chamfer_distance(x, y)
            out.backward()
        ms = timed(step, 6)
        print("| Chamfer | %d | %d | %.3f | %.1f | - |" % (B, N, ms, B / (ms * 1e-3)))
        sys.stdout.flush()
g = torch.Generator().manual_seed(11)
U, _ = torch.linalg.qr(torch.randn(512, 3, 2, generator=g)); U = U.to(dev)
th = F.normalize(torch.randn(512, 3, generator=g), dim=-1).to(dev)
for B in (1, 8):
    x, y = clouds(B, 4096, 77)
    for name, fn in (("SSW p=1 (level median), P=512", lambda: shwd.ops.spherical_sliced_w1(x, y, U).sum()),
                     ("SSW p=2 (bisection), P=512", lambda: shwd.ops.spherical_sliced_wp(x, y, U, 2.0).sum()),
                     ("Euclid SW p=2, P=512", lambda: shwd.ops.euclid_sliced_w(x, y, th, 2.0).sum())):
        def step():
            x.grad = None; y.grad = None
            fn().backward()
        ms = timed(step, 6)
        print("| %s | %d | 4096 | %.3f | %.1f | - |" % (name, B, ms, B / (ms * 1e-3)))
        sys.stdout.flush()
