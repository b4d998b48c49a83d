"""Exploratory edge-case sweep of the drop-in losses (not a test: prints what each call does)."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import shwd, oracle
L = shwd.losses
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def attempt(name, fn):
    try:
        out = fn()
        torch.cuda.synchronize()
        print("OK   %-70s %s" % (name, out))
    except Exception as e:
        print("EXC  %-70s %s: %s" % (name, type(e).__name__, str(e)[:160]))


x = torch.randn(4, 200, 3, generator=g)
y = torch.randn(4, 160, 3, generator=g) + 0.2
crit = L.Geodesic_distance_W(device=dev, p=2, eps=0.05, max_iter=20)
ref = oracle.entropic_w(x - x.mean(1, keepdim=True), y - y.mean(1, keepdim=True), "geodesic", 2, 0.05, 20)
attempt("geodesic sinkhorn float64 inputs", lambda: rel(crit((x - x.mean(1, keepdim=True)).double().to(dev), (y - y.mean(1, keepdim=True)).double().to(dev)), ref))
xt = (x - x.mean(1, keepdim=True)).transpose(1, 2).contiguous().transpose(1, 2).to(dev)  # non-contiguous view
attempt("geodesic sinkhorn non-contiguous x", lambda: rel(crit(xt, (y - y.mean(1, keepdim=True)).to(dev)), ref))
attempt("geodesic sinkhorn CPU tensors (moved by .to(device))", lambda: rel(crit(x - x.mean(1, keepdim=True), y - y.mean(1, keepdim=True)), ref))
attempt("geodesic sinkhorn half inputs", lambda: crit(x.half().to(dev), y.half().to(dev)).item())
attempt("geodesic sinkhorn B=0", lambda: crit(torch.empty(0, 10, 3, device=dev), torch.empty(0, 10, 3, device=dev)))
attempt("geodesic sinkhorn N=1", lambda: crit(torch.randn(2, 1, 3, device=dev), torch.randn(2, 1, 3, device=dev)).item())
attempt("geodesic sinkhorn zero point (norm 0)", lambda: crit(torch.zeros(1, 8, 3, device=dev), torch.randn(1, 8, 3, device=dev)).item())
attempt("geodesic sinkhorn dims != 3", lambda: crit(torch.randn(2, 8, 2, device=dev), torch.randn(2, 8, 2, device=dev)).item())
attempt("Cos_disimilarity_W exact default, N=64", lambda: L.Cos_disimilarity_W(dev, p=2)(torch.randn(2, 64, 3, device=dev), torch.randn(2, 64, 3, device=dev)).item())
attempt("Cos_disimilarity_W exact default, 64 vs 96", lambda: L.Cos_disimilarity_W(dev, p=2)(torch.randn(2, 64, 3, device=dev), torch.randn(2, 96, 3, device=dev)).item())
attempt("Cos_disimilarity_W p=1 un-batched", lambda: L.Cos_disimilarity_W(dev, p=1)(torch.randn(64, 3, device=dev), torch.randn(64, 3, device=dev)).item())
attempt("chamfer B=0", lambda: L.chamfer_distance(torch.empty(0, 10, 3, device=dev), torch.empty(0, 10, 3, device=dev)))
attempt("chamfer float64", lambda: L.chamfer_distance(x.double().to(dev), y.double().to(dev))[0].item() - oracle.chamfer_distance(x, y)[0].item())
attempt("chamfer N=1 M=5", lambda: L.chamfer_distance(torch.randn(2, 1, 3, device=dev), torch.randn(2, 5, 3, device=dev))[0].item())
attempt("chamfer batch_reduction=None", lambda: L.chamfer_distance(x.to(dev), y.to(dev), batch_reduction=None)[0].shape)
attempt("chamfer point_reduction=sum", lambda: L.chamfer_distance(x.to(dev), y.to(dev), point_reduction="sum")[0].item())
Xs, Xt = F.normalize(torch.randn(300, 3, generator=g), dim=-1), F.normalize(torch.randn(260, 3, generator=g), dim=-1)
attempt("sliced_wasserstein_sphere p=2", lambda: L.sliced_wasserstein_sphere(Xs.to(dev), Xt.to(dev), 50, dev, p=2).item())
attempt("sliced_wasserstein_sphere p=1", lambda: L.sliced_wasserstein_sphere(Xs.to(dev), Xt.to(dev), 50, dev, p=1).item())
attempt("sliced_wasserstein_sphere p=1.5", lambda: L.sliced_wasserstein_sphere(Xs.to(dev), Xt.to(dev), 50, dev, p=1.5).item())
attempt("sliced_wasserstein_sphere float64", lambda: L.sliced_wasserstein_sphere(Xs.double().to(dev), Xt.double().to(dev), 50, dev, p=2).item())
attempt("sliced_wasserstein_sphere identical clouds p=2", lambda: L.sliced_wasserstein_sphere(Xs.to(dev), Xs.to(dev), 50, dev, p=2).item())
attempt("sliced_wasserstein_sphere identical clouds p=1", lambda: L.sliced_wasserstein_sphere(Xs.to(dev), Xs.to(dev), 50, dev, p=1).item())
attempt("sliced_wasserstein_sphere 1 projection", lambda: L.sliced_wasserstein_sphere(Xs.to(dev), Xt.to(dev), 1, dev, p=2).item())
attempt("sliced_wasserstein_sphere N=1", lambda: L.sliced_wasserstein_sphere(Xs[:1].to(dev), Xt[:1].to(dev), 4, dev, p=2).item())
attempt("sliced_wasserstein_distance (notebook) n != m", lambda: L.sliced_wasserstein_distance(Xs.to(dev), Xt.to(dev), 20, 2, dev).item())
attempt("binary_search_circle keys == 1.0 exactly", lambda: L.binary_search_circle(torch.tensor([[0.0, 0.5, 1.0]], device=dev), torch.tensor([[0.25, 0.75, 1.0]], device=dev), p=2).item())
attempt("emd1D_circle p=2 (reference returns None)", lambda: L.emd1D_circle(torch.rand(2, 5, device=dev), torch.rand(2, 5, device=dev), p=2))
sk = L.log_Sinkhorn_Distance_Loss(0.05, 30, 'mean', 'L2')
attempt("log_Sinkhorn 'L3' norm", lambda: L.log_Sinkhorn_Distance_Loss(0.05, 30, 'mean', 'L3')(x.to(dev), y.to(dev), dev))
attempt("log_Sinkhorn float64", lambda: sk(x.double().to(dev), y.double().to(dev), dev)[0].item())
attempt("log_Sinkhorn returns (cost,P,C) shapes", lambda: [tuple(t.shape) for t in sk(x.to(dev), y.to(dev), dev)])
attempt("NaN cloud through geodesic sinkhorn", lambda: crit(torch.full((1, 8, 3), float("nan"), device=dev), torch.randn(1, 8, 3, device=dev)).item())
attempt("NaN cloud through exact", lambda: L.Cos_disimilarity_W(dev, p=2)(torch.full((1, 8, 3), float("nan"), device=dev), torch.randn(1, 8, 3, device=dev)).item())
attempt("after NaN: context still healthy", lambda: crit(torch.randn(1, 8, 3, device=dev), torch.randn(1, 8, 3, device=dev)).item())
