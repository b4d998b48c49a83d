"""cfg4-sized (N=16384, one pair) timings of every loss on the path (diagnostic): fwd+bwd ms."""
import os, sys
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd
dev = torch.device("cuda:0")
N, P = 16384, 512
g = torch.Generator().manual_seed(5)
x = F.normalize(torch.randn(1, N, 3, generator=g), dim=-1).to(dev).requires_grad_(True)
y = F.normalize(torch.randn(1, N, 3, generator=g) + 0.2, dim=-1).to(dev).requires_grad_(True)
U, _ = torch.linalg.qr(torch.randn(P, 3, 2, generator=g)); U = U.to(dev)
th = F.normalize(torch.randn(P, 3, generator=g), dim=-1).to(dev)
cases = {
    "SSW p=1 (level median)": lambda: shwd.ops.spherical_sliced_w1(x, y, U),
    "SSW p=2 (bisection)": lambda: shwd.ops.spherical_sliced_wp(x, y, U, 2.0),
    "Euclid SW p=2": lambda: shwd.ops.euclid_sliced_w(x, y, th, 2.0),
    "Chamfer": lambda: shwd.losses.chamfer_distance(x, y)[0],
    "W_COS geodesic p=2 L=100": lambda: shwd.entropic_ot(x, y, "geodesic", 2.0, 0.01, 100, center=True).cost,
}
for name, fn in cases.items():
    def step():
        x.grad = None; y.grad = None
        fn().sum().backward()
    step(); step(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("N=%d P=%d  %-75s %9.3f ms" % (N, P, name, min(ts)))
