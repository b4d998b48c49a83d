"""Small-problem regime of the W_COS loss: fwd, bwd and fwd+bwd time of the geodesic entropic OT
(eps=0.01, L=100) per shape, with the flattened-deal kernels (path 1) and the lean dedicated-CTA kernels (path 2) of
csrc/sinkhorn_lean.cu, plus what the automatic selection picks.  Kernel-only times (CUDA events around the C-ABI
launches).  Run on the GPU box:  python tools/sweep_small.py > gpurun_out/sweep_small.md"""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd
dev = torch.device("cuda:0")
L, EPS = 100, 0.01
lib = shwd._lib.lib()
PEAK = 36.6e12


def timed(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def clouds(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    x = F.normalize(torch.randn(B, N, 3, generator=g), dim=-1)
    y = F.normalize(torch.randn(B, N, 3, generator=g) + 0.2, dim=-1)
    return x.to(dev).requires_grad_(True), y.to(dev).requires_grad_(True)


shapes = [(1, 256), (1, 512), (1, 1024), (1, 2048), (2, 1024), (4, 1024), (8, 1024), (16, 1024), (32, 1024), (4, 512), (32, 128), (32, 256),
          (32, 512), (64, 256), (148, 256), (256, 256), (256, 128), (512, 256)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
print("| B | N | path | lean regime (auto) | fwd ms | bwd ms | fwd+bwd ms | pairs/s | frac of FP32 roofline |")
print("|---|---|---|---|---|---|---|---|---|")
for B, N in shapes:
    x, y = clouds(B, N, N + B)
    work = (2 * L + 1) * N * N * B
    for name, mode in (("flat", 1), ("lean", 2)):
        lib.shwd_sinkhorn_set_path(mode)
        if mode == 2 and not lib.shwd_sinkhorn_lean_regime(B, N, N):
            continue
        state = {}

        def fwd():
            state["res"] = shwd.entropic_ot(x, y, "geodesic", 2.0, EPS, L, center=True)
            state["loss"] = state["res"].cost.sum()

        def bwd():
            x.grad = None; y.grad = None
            state["loss"].backward(retain_graph=True)

        def both():
            fwd(); bwd()
        tf = timed(fwd)
        fwd()
        tb = timed(bwd)
        tt = timed(both)
        assert state["res"].status() == 0
        lib.shwd_sinkhorn_set_path(0)
        auto = lib.shwd_sinkhorn_lean_regime(B, N, N)
        print("| %d | %d | %s | %d | %.3f | %.3f | %.3f | %.1f | %.3f |" % (B, N, name, auto, tf, tb, tt, B / (tt * 1e-3), 54.0 * work / (tt * 1e-3) / PEAK))
        sys.stdout.flush()
lib.shwd_sinkhorn_set_path(0)
