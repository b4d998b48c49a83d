"""Host time of one eager loss call (fwd + bwd) at sizes where the kernels are negligible, and where it goes (cProfile)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import shwd
L = shwd.losses
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randn(4, 64, 3, generator=g).to(dev).requires_grad_(True)
y = torch.randn(4, 64, 3, generator=g).to(dev).requires_grad_(True)
U, _ = torch.linalg.qr(torch.randn(8, 3, 2, generator=g)); U = U.to(dev)
crit = L.Geodesic_distance_W(device=dev, p=2, eps=0.05, max_iter=5)
cases = {
    "chamfer_distance": lambda: L.chamfer_distance(x, y)[0],
    "sliced_cost p=2": lambda: L.sliced_cost(x, y, U, p=2).mean(),
    "sliced_cost p=1": lambda: L.sliced_cost(x, y, U, p=1).mean(),
    "Geodesic_distance_W (sinkhorn, 5 iterations)": lambda: crit(x, y),
    "Cos_disimilarity_W (exact)": lambda: L.Cos_disimilarity_W(dev, p=2)(x, y),
}
for name, fn in cases.items():
    def step():
        x.grad = y.grad = None
        fn().backward()
    for _ in range(20):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("%-46s host %.1f us per fwd+bwd call" % (name, (t1 - t0) / 200 * 1e6))
    if "--profile" in sys.argv:
        pr = cProfile.Profile(); pr.enable()
        for _ in range(200):
            step()
        pr.disable()
        s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14)
        print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:26]))
