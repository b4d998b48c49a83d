"""How far the entropic loss (eps=0.01, L=100: the north-star kernel) sits from the exact EMD the reference's W_COS path
computes (s2_wasserstein.py:41-44), on the benchmark inputs (cfg2 registration pairs, centred).  Information only -- the
drop-in's default is the exact solve; the entropic solver is what `eps=` / `max_iter=` select.
    python tools/entropic_gap.py > gpurun_out/entropic_gap.md"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd
import bench
dev = torch.device("cuda:0")
L = shwd.losses
t, s = bench.registration_pairs(32, 1024, 1234, dev)
t = t - t.mean(1, keepdim=True)
s = s - s.mean(1, keepdim=True)
print("| loss object (p=2) | exact (default) | entropic eps=0.01 L=100 | entropic eps=0.01 L=1000 | entropic eps=0.001 L=1000 | rel. gap of the benchmark setting |")
print("|---|---|---|---|---|---|")
for name, cls in (("Geodesic_distance_W", L.Geodesic_distance_W), ("Cos_disimilarity_W", L.Cos_disimilarity_W)):
    ex = cls(dev, p=2)(t, s).item()
    e1 = cls(dev, p=2, eps=0.01, max_iter=100)(t, s).item()
    e2 = cls(dev, p=2, eps=0.01, max_iter=1000)(t, s).item()
    e3 = cls(dev, p=2, eps=0.001, max_iter=1000)(t, s).item()
    print("| %s | %.6f | %.6f | %.6f | %.6f | %+.2f %% |" % (name, ex, e1, e2, e3, 100 * (e1 - ex) / ex))
