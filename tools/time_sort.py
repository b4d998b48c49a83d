"""Segmented sort timing by row length (diagnostic): ms per call and Gkeys/s, int32-permutation entry point."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd
dev = torch.device("cuda:0")
for segs, length in ((2048, 6500), (1024, 8192), (512, 8192), (512, 10240), (512, 16384), (1024, 16384), (512, 32768), (128, 65536)):
    k = torch.rand(segs, length, device=dev)
    for _ in range(3):
        shwd.ops._sort_i32(k)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); shwd.ops._sort_i32(k); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print("%5d segments x %6d keys  %8.3f ms  %6.1f Gkeys/s" % (segs, length, t, segs * length / t * 1e-6))
