"""Kernel-only time of circular_w1 (level median of the circular W1) at cfg3 and cfg4 row sizes and a few ragged shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import shwd
dev = torch.device("cuda:0")
for S, n, m in ((4096, 4096, 4096), (4096, 4000, 3500), (4096, 1024, 1024), (512, 16384, 16384), (1024, 5000, 5000), (1024, 6000, 6000), (1024, 8192, 8192), (4096, 300, 250)):
    g = torch.Generator().manual_seed(11)
    us = torch.sort(torch.rand(S, n, generator=g), -1)[0].to(dev)
    vs = torch.sort((torch.rand(S, m, generator=g) * 0.8 + 0.15) % 1.0, -1)[0].to(dev)
    if n + m > shwd.ops.CIRCULAR_W1_MAX:
        continue
    for _ in range(2):
        shwd.ops.CircularW1Fn.apply(us, vs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        w = shwd.ops.CircularW1Fn.apply(us, vs)
    e1.record(); torch.cuda.synchronize()
    print("S=%d n=%d m=%d: %.1f us   mean W %.7e" % (S, n, m, e0.elapsed_time(e1) / 5 * 1e3, w.mean().item()))
