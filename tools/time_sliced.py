"""Per-kernel time of the spherical sliced losses at BASELINE config 3 (N=4096, 512 slices), B pairs (diagnostic;
run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split)."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
P = float(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator().manual_seed(11)
x = F.normalize(torch.randn(B, 4096, 3, generator=g), dim=-1).to(dev).requires_grad_(True)
y = F.normalize(torch.randn(B, 4096, 3, generator=g) + 0.2, dim=-1).to(dev).requires_grad_(True)
U, _ = torch.linalg.qr(torch.randn(512, 3, 2, generator=g)); U = U.to(dev)
for it in range(3):
    x.grad = None; y.grad = None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    w = shwd.ops.spherical_sliced_w1(x, y, U) if P == 1 else shwd.ops.spherical_sliced_wp(x, y, U, P)
    w.sum().backward()
    e1.record(); torch.cuda.synchronize()
print("B=%d p=%s fwd+bwd %.3f ms (%.3f ms/pair)" % (B, P, e0.elapsed_time(e1), e0.elapsed_time(e1) / B))
