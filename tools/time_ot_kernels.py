"""Device time of the two OT launches alone (shwd_sinkhorn_fwd / shwd_sinkhorn_bwd through ctypes on preallocated
buffers; CUDA events on the launching stream; the sentinel memsets the launchers enqueue are inside the timed region),
per shape and kernel family (flat = flattened deal, lean = dedicated CTAs).  Geodesic cost, p = 2, eps = 0.01, L = 100.
    python tools/time_ot_kernels.py [BxN ...] > gpurun_out/ot_kernels.md"""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd
from shwd_b200 import _lib
dev = torch.device("cuda:0")
lib = _lib.lib()
L, EPS, PEAK = 100, 0.01, 36.6e12


def ptr(t):
    return t.data_ptr()


def run(B, N, mode, reps=9):
    lib.shwd_sinkhorn_set_path(mode)
    g = torch.Generator().manual_seed(B + N)
    x = F.normalize(torch.randn(B, N, 3, generator=g), dim=-1).to(dev)
    y = F.normalize(torch.randn(B, N, 3, generator=g) + 0.2, dim=-1).to(dev)
    f32 = dict(device=dev, dtype=torch.float32)
    x4, y4 = torch.empty(B, N, 4, **f32), torch.empty(B, N, 4, **f32)
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.shwd_sphere_map_fwd(ptr(x), ptr(x4), None, B, N, 3, s), "map")
    _lib.check(lib.shwd_sphere_map_fwd(ptr(y), ptr(y4), None, B, N, 3, s), "map")
    HL = L + 1
    alpha, beta = torch.empty(2, B, HL, N, **f32), torch.empty(2, B, HL, N, **f32)
    rpc, cpc, cost = torch.empty(B, N, **f32), torch.empty(B, N, **f32), torch.empty(B, **f32)
    it = torch.empty(1, device=dev, dtype=torch.int32)
    wsb = lib.shwd_sinkhorn_workspace_bytes(B, N, N, L)
    ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
    gc = torch.ones(B, **f32)
    g4x, g4y = torch.empty(B, N, 4, **f32), torch.empty(B, N, 4, **f32)

    def fwd():
        _lib.check(lib.shwd_sinkhorn_fwd(ptr(x4), ptr(y4), B, N, N, 0, 2.0, 1.0, EPS, L, 0.0, HL, ptr(alpha), ptr(beta), ptr(rpc),
                                         ptr(cpc), ptr(cost), ptr(it), ptr(ws), wsb, s), "fwd")

    def bwd():
        _lib.check(lib.shwd_sinkhorn_bwd(ptr(x4), ptr(y4), B, N, N, 0, 2.0, 1.0, EPS, L, ptr(alpha), ptr(beta), ptr(rpc), ptr(cpc),
                                         ptr(it), ptr(gc), ptr(g4x), ptr(g4y), ptr(ws), wsb, s), "bwd")

    def timed(fn):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]
    tf = timed(fwd)
    fwd()
    tb = timed(bwd)
    st = int(ws[:4].view(torch.int32).item())
    lib.shwd_sinkhorn_set_path(0)
    return tf, tb, st, float(cost.sum().item())


shapes = [(1, 256), (1, 1024), (1, 2048), (4, 1024), (8, 1024), (16, 1024), (32, 1024), (32, 256), (32, 512), (64, 256), (256, 256)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
print("| B | N | path | fwd ms | bwd ms | fwd+bwd ms | us per half-step (fwd / bwd) | frac of FP32 roofline | status | cost |")
print("|---|---|---|---|---|---|---|---|---|---|")
for B, N in shapes:
    work = (2 * L + 1) * N * N * B
    for name, mode in (("flat", 1), ("lean", 2)):
        if mode == 2:
            lib.shwd_sinkhorn_set_path(2)
            ok = lib.shwd_sinkhorn_lean_regime(B, N, N)
            lib.shwd_sinkhorn_set_path(0)
            if not ok:
                continue
        tf, tb, st, c = run(B, N, mode)
        print("| %d | %d | %s | %.3f | %.3f | %.3f | %.2f / %.2f | %.3f | %d | %.6f |" % (
            B, N, name, tf, tb, tf + tb, tf * 1e3 / (2 * L + 1), tb * 1e3 / (2 * L + 1), 54.0 * work / ((tf + tb) * 1e-3) / PEAK, st, c))
        sys.stdout.flush()
