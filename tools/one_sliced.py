"""Two fwd+bwd steps of the cfg3 sliced loss (p from argv, default 2) -- the command profiled by ncu for the sliced kernels."""
import os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shwd
dev = torch.device("cuda:0")
p = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
g = torch.Generator().manual_seed(11)
x = F.normalize(torch.randn(8, 4096, 3, generator=g), dim=-1).to(dev).requires_grad_(True)
y = F.normalize(torch.randn(8, 4096, 3, generator=g) + 0.2, dim=-1).to(dev).requires_grad_(True)
U, _ = torch.linalg.qr(torch.randn(512, 3, 2, generator=g)); U = U.to(dev)
for _ in range(2):
    x.grad = y.grad = None
    w = shwd.ops.spherical_sliced_w1(x, y, U) if p == 1 else shwd.ops.spherical_sliced_wp(x, y, U, p)
    w.sum().backward()
torch.cuda.synchronize()
print("ok", w.mean().item())
