"""One full SHWD training-loss step as train_W_COS.py runs it (s2_wasserstein.py:234-262): inner ascent on phi
(max_iter=1) + outer loss + backward to the cloud, B=32, N=1024, Residual x3 phi, geodesic p=2, eps=0.01, L=100.
Times the step with the fused phi kernel and with the eager torch modules (diagnostic)."""
import os, sys, time
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd
import bench
dev = torch.device("cuda:0")
B, N = 32, 1024
tmpl, src = bench.registration_pairs(B, N, 1234, dev)
tmpl = tmpl - tmpl.mean(1, keepdim=True)
src = (src - src.mean(1, keepdim=True)).requires_grad_(True)
for mode in ("fused", "eager"):
    torch.manual_seed(0)
    phi = shwd.losses.Norm_Flow_structure(flow_name="Residual", n_flow_layer=3).to(dev)
    if mode == "eager":
        phi.forward = phi.forward_eager
    opt = torch.optim.Adam(phi.parameters(), lr=1e-3)
    crit = shwd.losses.max_cos_disimilarity_wassersten_distance(phi, shwd.losses.Geodesic_distance_W(dev, p=2, solver="sinkhorn"), dev, opt,
                                                                max_iter=1, lam=0.1)
    ts, pts = [], []
    for it in range(8):
        src.grad = None
        torch.cuda.synchronize(); t0 = time.perf_counter()
        loss, _, _ = crit(tmpl, src, "train")
        loss.backward()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        if it >= 3:
            ts.append((t1 - t0) * 1e3)
    # phi alone: forward + backward on both clouds
    for it in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        y = phi(tmpl); z = phi(src)
        (y.sum() + z.sum()).backward()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        if it >= 3:
            pts.append((t1 - t0) * 1e3)
    print("%s phi: training-loss step %.2f ms (%.0f pairs/s), phi fwd+bwd on both clouds alone %.3f ms, loss %.6f" % (
        mode, min(ts), B / (min(ts) * 1e-3), min(pts), loss.item()))

# The criterion train_W_COS.py:393 builds -- Cos_disimilarity_W(device, p=2), nothing else -- i.e. the exact solve (auction
# kernel) inside the same max-over-phi wrapper: two exact B=32 solves per step (inner ascent + outer loss).
torch.manual_seed(0)
phi = shwd.losses.Norm_Flow_structure(flow_name="Residual", n_flow_layer=3).to(dev)
opt = torch.optim.Adam(phi.parameters(), lr=1e-3)
crit = shwd.losses.max_cos_disimilarity_wassersten_distance(phi, shwd.losses.Cos_disimilarity_W(dev, p=2), dev, opt, max_iter=1, lam=0.1)
ts = []
for it in range(8):
    src.grad = None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loss, _, _ = crit(tmpl, src, "train")
    loss.backward()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if it >= 3:
        ts.append((t1 - t0) * 1e3)
one = shwd.losses.Cos_disimilarity_W(dev, p=2)
t1s = []
for it in range(8):
    src.grad = None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    one(tmpl, src).backward()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if it >= 3:
        t1s.append((t1 - t0) * 1e3)
print("exact criterion (reference default): training-loss step %.2f ms (%.0f pairs/s); one Cos_disimilarity_W call fwd+bwd %.2f ms, loss %.6f" % (
    min(ts), B / (min(ts) * 1e-3), min(t1s), loss.item()))
