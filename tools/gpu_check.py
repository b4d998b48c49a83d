"""Diagnostic run on a B200 (not a test): parity of the CUDA path against the CPU oracle / golden fixtures with
verbose numbers, plus a quick timing.  Usage: python tools/gpu_check.py [quick]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import shwd  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
dev = torch.device("cuda:0")


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def check_sphere_map():
    torch.manual_seed(0)
    for (B, N) in ((3, 64), (2, 50), (4, 1024)):
        x = torch.randn(B, N, 3) * 2 + 0.3
        for center in (True, False):
            for normalize in (True, False):
                xr = x.clone().requires_grad_(True)
                ref = oracle.sphere_map(xr, center, normalize)
                w = torch.randn_like(ref)
                (ref * w).sum().backward()
                xg = x.clone().to(dev).requires_grad_(True)
                out = shwd.sphere_map(xg, center, normalize)
                (out * w.to(dev)).sum().backward()
                print(f"sphere_map B{B} N{N} c{int(center)} n{int(normalize)}: out rel {rel(out, ref):.2e} max|d| "
                      f"{(out.cpu() - ref).abs().max().item():.2e}  grad rel {rel(xg.grad, xr.grad):.2e}")
        xr = x.clone().requires_grad_(True)
        r = oracle.flow_regularization(xr)
        r.backward()
        xg = x.clone().to(dev).requires_grad_(True)
        rg = shwd.flow_regularization(xg)
        rg.backward()
        print(f"  regulariser: {r.item():.6f} vs {rg.item():.6f}  grad rel {rel(xg.grad, xr.grad):.2e}")


def run_case(tag, x, y, kind, p, eps, iters, n_power=1.0, thresh=0.0, ref=None):
    xg = x.clone().to(dev).requires_grad_(True)
    yg = y.clone().to(dev).requires_grad_(True)
    t0 = time.time()
    res = shwd.entropic_ot(xg, yg, kind, p, eps, iters, n_power, thresh)
    cost = res.cost
    if n_power != 1.0:
        cost = cost.pow(1.0 / n_power)
    tot = cost.sum()
    tot.backward()
    torch.cuda.synchronize()
    t1 = time.time()
    st = res.status()
    if ref is None:
        xr = x.clone().requires_grad_(True)
        yr = y.clone().requires_grad_(True)
        c = oracle.log_sinkhorn(xr, yr, kind, p, eps, iters, thresh if thresh > 0 else None, "none", n_power)
        c.sum().backward()
        ref = (c.detach().reshape(-1), xr.grad.reshape(xg.shape), yr.grad.reshape(yg.shape))
    c_ref, gx_ref, gy_ref = ref
    print(f"{tag}: status {st} it {res.iterations()} cost rel {rel(cost, c_ref):.2e} (|cost| {c_ref.abs().mean().item():.4g})  "
          f"gx rel {rel(xg.grad, gx_ref):.2e}  gy rel {rel(yg.grad, gy_ref):.2e}  [{(t1 - t0) * 1e3:.1f} ms]")
    sys.stdout.flush()
    return res


def check_sinkhorn():
    def gold(name):
        d = dict(np.load(os.path.join(G, name + ".npz")))
        return d

    for name, kind, p, npow, thresh in (
            ("geodesic_sinkhorn_p2", "geodesic", 2, 1, 0), ("geodesic_sinkhorn_p2_ragged", "geodesic", 2, 1, 0),
            ("geodesic_sinkhorn_p1", "geodesic", 1, 1, 0), ("sinkhorn_cmp_L2", "sqeuclid", 2, 1, 1e-9),
            ("sinkhorn_cmp_L1_sum", "sqeuclid", 1, 1, 1e-9), ("sinkhorn_plain_L2_none", "sqeuclid", 2, 1, 0),
            ("sinkhorn_fixed_L2", "euclid", 2, 1, 1e-9), ("sinkhorn_logN_2", "sqeuclid", 2, 2, 1e-9)):
        d = gold(name)
        x, y = torch.from_numpy(d["x"]), torch.from_numpy(d["y"])
        red = str(d["batch_reduction"])
        B = x.shape[0]
        # fixture grads are of the reduced loss; undo the reduction so everything is compared per pair-sum
        scale = B if red == "mean" else 1.0
        xr = x.clone().requires_grad_(True)
        yr = y.clone().requires_grad_(True)
        c = oracle.log_sinkhorn(xr, yr, kind, p, float(d["eps"]), int(d["max_iter"]), thresh if thresh > 0 else None, "none", npow)
        c.sum().backward()
        print(f"  [fixture {name}: oracle-vs-reference gx rel {rel(xr.grad, torch.from_numpy(d['gx']) * scale):.1e}]")
        run_case(name, x, y, kind, p, float(d["eps"]), int(d["max_iter"]), npow, thresh, ref=(c.detach(), xr.grad, yr.grad))
    # more shapes against the oracle
    torch.manual_seed(5)
    for (B, N, M, L, eps) in ((1, 33, 70, 10, 0.05), (5, 256, 256, 20, 0.01), (2, 300, 1000, 15, 0.02), (3, 2500, 700, 6, 0.05)):
        x = torch.nn.functional.normalize(torch.randn(B, N, 3), dim=-1)
        y = torch.nn.functional.normalize(torch.randn(B, M, 3) + 0.3, dim=-1)
        run_case(f"geo2 B{B} N{N} M{M} L{L}", x, y, "geodesic", 2, eps, L)
    x = torch.randn(2, 100, 3)
    y = torch.randn(2, 90, 3)
    run_case("one_minus_cos p2", x, y, "one_minus_cos", 2, 0.05, 20)
    run_case("geodesic p1.5", x, y, "geodesic", 1.5, 0.05, 20)
    run_case("euclid p1", x, y, "euclid", 1, 0.1, 20)
    run_case("sqeuclid p3", x, y, "sqeuclid", 3, 0.1, 20)


def timing():
    torch.manual_seed(1234)
    B, N, L = 32, 1024, 100
    x = torch.nn.functional.normalize(torch.randn(B, N, 3), dim=-1).to(dev).requires_grad_(True)
    y = torch.nn.functional.normalize(torch.randn(B, N, 3), dim=-1).to(dev).requires_grad_(True)
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        res = shwd.entropic_ot(x, y, "geodesic", 2.0, 0.01, L, center=True)
        e1.record()
        res.cost.sum().backward()
        e2.record()
        torch.cuda.synchronize()
        f, b = e0.elapsed_time(e1), e1.elapsed_time(e2)
        print(f"timing B{B} N{N} L{L}: fwd {f:.2f} ms  bwd {b:.2f} ms  -> {B / ((f + b) * 1e-3):.0f} pairs/s  status {res.status()}")
    sys.stdout.flush()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "SMs", shwd._lib.lib().shwd_device_sm_count())
    check_sphere_map()
    check_sinkhorn()
    timing()
