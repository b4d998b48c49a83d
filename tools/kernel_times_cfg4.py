"""Per-kernel CUDA time (torch profiler) of one fwd+bwd of each cfg4-size loss (one pair, N = 16384, 512 slices)."""
import os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shwd
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
N, P = 16384, 512
g = torch.Generator().manual_seed(5)
x = F.normalize(torch.randn(1, N, 3, generator=g), dim=-1).to(dev).requires_grad_(True)
y = F.normalize(torch.randn(1, N, 3, generator=g) + 0.2, dim=-1).to(dev).requires_grad_(True)
U, _ = torch.linalg.qr(torch.randn(P, 3, 2, generator=g)); U = U.to(dev)
th = F.normalize(torch.randn(P, 3, generator=g), dim=-1).to(dev)
cases = {
    "SSW p=1": lambda: shwd.ops.spherical_sliced_w1(x, y, U),
    "SSW p=2": lambda: shwd.ops.spherical_sliced_wp(x, y, U, 2.0),
    "Euclid SW p=2": lambda: shwd.ops.euclid_sliced_w(x, y, th, 2.0),
    "Chamfer": lambda: shwd.losses.chamfer_distance(x, y)[0],
}
for name, fn in cases.items():
    def step():
        x.grad = None; y.grad = None
        fn().sum().backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): step()
        torch.cuda.synchronize()
    evs = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    print("%s: %.1f us of kernels per step" % (name, sum(e.device_time_total for e in evs) / 3))
    for ev in evs[:7]:
        print("  %-70s n=%3d  %9.1f us/step" % (ev.key[:70], ev.count, ev.device_time_total / 3))
