"""Chamfer forward / backward kernel times (CUDA events around the C-ABI launches) at the sweep shapes (diagnostic)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd
dev = torch.device("cuda:0")
for B, N in ((1, 1024), (1, 4096), (1, 16384), (2, 16384), (1, 65536), (32, 1024), (256, 1024), (32, 4096), (32, 16384), (4, 65536)):
    g = torch.Generator().manual_seed(N)
    x = torch.randn(B, N, 3, generator=g).to(dev).requires_grad_(True)
    y = (torch.randn(B, N, 3, generator=g) * 1.1 + 0.1).to(dev).requires_grad_(True)
    def fwd():
        return shwd.losses.chamfer_distance(x, y)[0]
    for _ in range(3):
        x.grad = None; y.grad = None
        fwd().backward()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    R = 10
    for _ in range(R):
        x.grad = None; y.grad = None
        torch.cuda.synchronize()
        ev[0].record(); l = fwd(); ev[1].record(); l.backward(); ev[2].record()
        torch.cuda.synchronize()
        tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
    pairs = 2.0 * B * N * N
    print("B=%d N=%d  fwd %.3f ms (%.1f Tlane-op/s of 8 per candidate)  bwd %.3f ms" % (B, N, tf / R, 8 * pairs / (tf / R * 1e-3) / 1e12, tb / R))
