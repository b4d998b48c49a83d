"""One fwd+bwd of the spherical sliced W1 loss at cfg4 size (N = 16384, 512 slices) -- the command profiled for
profiles/r01h_ncu_cfg4_sliced_summary.txt (compact sort kernel, circular_w1_kernel<64>)."""
import os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shwd
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)
x = F.normalize(torch.randn(1, 16384, 3, generator=g), dim=-1).to(dev).requires_grad_(True)
y = F.normalize(torch.randn(1, 16384, 3, generator=g) + 0.2, dim=-1).to(dev).requires_grad_(True)
U, _ = torch.linalg.qr(torch.randn(512, 3, 2, generator=g))
for _ in range(2):
    x.grad = None
    y.grad = None
    shwd.ops.spherical_sliced_w1(x, y, U.to(dev)).sum().backward()
torch.cuda.synchronize()
print("ok")
