"""Per-kernel CUDA time via torch profiler for one fwd+bwd of the sliced loss."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch, torch.nn.functional as F
import shwd
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
B = 8
for P_ in (1.0, 2.0):
    g = torch.Generator().manual_seed(11)
    x = F.normalize(torch.randn(B, 4096, 3, generator=g), dim=-1).to(dev).requires_grad_(True)
    y = F.normalize(torch.randn(B, 4096, 3, generator=g) + 0.2, dim=-1).to(dev).requires_grad_(True)
    U, _ = torch.linalg.qr(torch.randn(512, 3, 2, generator=g)); U = U.to(dev)
    def step():
        x.grad = None; y.grad = None
        w = shwd.ops.spherical_sliced_w1(x, y, U) if P_ == 1 else shwd.ops.spherical_sliced_wp(x, y, U, P_)
        w.sum().backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): step()
    e1.record(); torch.cuda.synchronize()
    print("p=%s B=%d fwd+bwd %.3f ms/step" % (P_, B, e0.elapsed_time(e1) / 5))
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): step()
        torch.cuda.synchronize()
    for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:10]:
        print("  %-60s n=%3d  %9.1f us/step" % (ev.key[:60], ev.count, ev.device_time_total / 3))
