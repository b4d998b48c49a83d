"""Where the end-to-end leg of bench.py spends its time beyond the two sweep kernels (diagnostic, cfg2 workload).
Each variant is timed like bench.py's e2e leg: CUDA events around every step, L2 flushed between steps, host running ahead."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import shwd  # noqa: E402

dev = torch.device("cuda:0")
B, N = bench.B_PER_GPU, bench.N_PTS
tmpl, src = bench.registration_pairs(B, N, 1234, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
crit = shwd.losses.Geodesic_distance_W(device=dev, p=int(bench.P_COST), eps=bench.EPS, max_iter=bench.ITERS)
h_t, h_s = tmpl.cpu().pin_memory(), src.cpu().pin_memory()
h_gx, h_gy = torch.empty_like(h_t).pin_memory(), torch.empty_like(h_s).pin_memory()
h_loss = torch.empty(1).pin_memory()
print("pinned:", h_t.is_pinned(), h_gx.is_pinned(), h_loss.is_pinned())


def resident():
    x = tmpl.detach().requires_grad_(True)
    y = src.detach().requires_grad_(True)
    res = shwd.entropic_ot(x, y, "geodesic", bench.P_COST, bench.EPS, bench.ITERS, center=True)
    res.cost.mean().backward()


def crit_only():
    x = tmpl.detach().requires_grad_(True)
    y = src.detach().requires_grad_(True)
    crit(x, y).backward()


def crit_centred(h2d=False, d2h=False):
    x = h_t.to(dev, non_blocking=True) if h2d else tmpl
    y = h_s.to(dev, non_blocking=True) if h2d else src
    x = x - x.mean(dim=1, keepdim=True)
    y = y - y.mean(dim=1, keepdim=True)
    x.requires_grad_(True)
    y.requires_grad_(True)
    loss = crit(x, y)
    loss.backward()
    if d2h:
        h_loss.copy_(loss.detach().reshape(1), non_blocking=True)
        h_gx.copy_(x.grad, non_blocking=True)
        h_gy.copy_(y.grad, non_blocking=True)


def timed(name, fn, steps=12, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = []
    t0 = time.perf_counter()
    host = 0.0
    for _ in range(steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h0 = time.perf_counter()
        fn()
        host += time.perf_counter() - h0
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    print("%-34s %8.3f ms/step (events)   host enqueue %6.3f ms/step   wall %7.3f ms/step" % (name, ms, host / steps * 1e3, wall / steps * 1e3))


timed("resident entropic_ot(center)", resident)
timed("crit (no centring, no copies)", crit_only)
timed("centring + crit", lambda: crit_centred(False, False))
timed("H2D + centring + crit", lambda: crit_centred(True, False))
timed("centring + crit + D2H", lambda: crit_centred(False, True))
timed("full e2e", lambda: crit_centred(True, True))
