"""Three forward calls of chamfer_distance on one pair of 16384-point clouds -- the command profiled for the single-pair kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shwd
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(16384)
x = torch.randn(1, 16384, 3, generator=g).to(dev)
y = (torch.randn(1, 16384, 3, generator=g) * 1.1 + 0.1).to(dev)
for _ in range(3):
    l = shwd.losses.chamfer_distance(x, y)[0]
torch.cuda.synchronize()
print("ok", l.item())
