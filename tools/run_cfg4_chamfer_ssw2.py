"""One fwd+bwd of Chamfer and of the spherical sliced W_2 loss at cfg4 size (one pair, N = 16384, 512 slices) -- the command
profiled for profiles/r01h_ncu_cfg4_chamfer_wp_summary.txt (chamfer_fwd_kernel<8>, circular_wp_kernel<true, 1024>)."""
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
    shwd.losses.chamfer_distance(x, y)[0].backward()
    shwd.ops.spherical_sliced_wp(x, y, U.to(dev), 2.0).sum().backward()
torch.cuda.synchronize()
print("ok")
