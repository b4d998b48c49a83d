"""Diagnostic: circular W_p kernel vs the CPU oracle (loss, theta, gradients) on a few shapes."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd, oracle
from oracle import sliced as osl
dev = torch.device("cuda:0")
def rel(a, b):
    a = a.detach().double().cpu().reshape(-1); b = b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
for (S, n, m, p, seed) in [(9, 70, 55, 2, 0), (5, 64, 64, 2, 1), (6, 500, 333, 2, 2), (4, 1024, 1024, 2, 3), (3, 300, 300, 3, 4), (3, 200, 257, 1.5, 5), (2, 4096, 4096, 2, 6), (4, 1, 5, 2, 7), (4, 7, 1, 2, 8)]:
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(S, n, generator=g); v = (torch.rand(S, m, generator=g) * 0.7 + 0.2) % 1.0
    ur = u.clone().requires_grad_(True); vr = v.clone().requires_grad_(True)
    t0 = time.time()
    wr, thr = osl.binary_search_circle(ur, vr, p=p, return_theta=True)
    t1 = time.time()
    wr.sum().backward()
    ug = u.clone().to(dev).requires_grad_(True); vg = v.clone().to(dev).requires_grad_(True)
    us, _ = shwd.ops.SegmentedSortFn.apply(ug); vs, _ = shwd.ops.SegmentedSortFn.apply(vg)
    torch.cuda.synchronize(); t2 = time.time()
    w, th = shwd.ops.CircularWpFn.apply(us, vs, float(p), -1.0, 1.0, 1e-7)
    torch.cuda.synchronize(); t3 = time.time()
    w.sum().backward()
    print("S%d n%d m%d p%s: w rel %.2e  max|dtheta| %.2e  gu rel %.2e  gv rel %.2e   [cpu %.1f ms, gpu %.2f ms]" % (
        S, n, m, p, rel(w, wr), (th.cpu() - thr).abs().max().item(), rel(ug.grad, ur.grad), rel(vg.grad, vr.grad), (t1 - t0) * 1e3, (t3 - t2) * 1e3))
    if rel(vg.grad, vr.grad) > 1e-5 or rel(ug.grad, ur.grad) > 1e-5:
        for name, a, b in (("gv", vg.grad.cpu(), vr.grad), ("gu", ug.grad.cpu(), ur.grad)):
            dd = (a - b).abs()
            idx = torch.nonzero(dd > 1e-9)
            print("   ", name, "mismatches:", idx.shape[0], "max", dd.max().item(), "|g|max", b.abs().max().item())
            for r, c in idx[:6].tolist():
                srt, perm = torch.sort((v if name == "gv" else u)[r])
                rank = (perm == c).nonzero().item()
                print("      row %d col %d (sorted rank %d of %d): ours %.6e ref %.6e  theta %.9f" % (r, c, rank, a.shape[1], a[r, c].item(), b[r, c].item(), thr[r].item()))
# fixture frozen from the reference
import numpy as np
d = dict(np.load(os.path.join(ROOT, "tests/golden/binary_search_circle_p2.npz")))
w = shwd.losses.binary_search_circle(torch.from_numpy(d["u"]).to(dev), torch.from_numpy(d["v"]).to(dev), p=2)
print("fixture binary_search_circle_p2: rel %.2e" % rel(w, torch.from_numpy(d["w"])))
d = dict(np.load(os.path.join(ROOT, "tests/golden/ssw_p2.npz")))
xs = torch.from_numpy(d["Xs"]).to(dev).requires_grad_(True); xt = torch.from_numpy(d["Xt"]).to(dev).requires_grad_(True)
loss = shwd.losses.sliced_cost(xs, xt, torch.from_numpy(d["U"]).to(dev), p=2)
loss.backward()
print("fixture ssw_p2: loss rel %.2e gx rel %.2e gy rel %.2e" % (abs(loss.item() - float(d["loss"])) / float(d["loss"]), rel(xs.grad, torch.from_numpy(d["gx"])), rel(xt.grad, torch.from_numpy(d["gy"]))))
# cfg3-sized timing: N=4096, P=512
g = torch.Generator().manual_seed(11)
Xs = torch.nn.functional.normalize(torch.randn(4096, 3, generator=g), dim=-1).to(dev).requires_grad_(True)
Xt = torch.nn.functional.normalize(torch.randn(4096, 3, generator=g) + 0.3, dim=-1).to(dev)
U, _ = torch.linalg.qr(torch.randn(512, 3, 2, generator=g)); U = U.to(dev)
for p in (1, 2):
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        l = shwd.losses.sliced_cost(Xs, Xt, U, p=p); l.backward()
        torch.cuda.synchronize(); t1 = time.time()
    print("cfg3 N=4096 P=512 p=%d: loss %.6f  fwd+bwd %.2f ms" % (p, l.item(), (t1 - t0) * 1e3))
