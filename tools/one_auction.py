"""One exact-assignment launch (B = 1, N = 1024, squared Euclidean, bench clouds) -- the command profiled under ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shwd, bench
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
tmpl, src = bench.registration_pairs(32, 1024, 1234, dev)
tmpl = tmpl - tmpl.mean(1, keepdim=True); src = src - src.mean(1, keepdim=True)
for _ in range(2):
    sig = shwd.ops.exact_assignment(tmpl[:B], src[:B], "sqeuclid", 2.0)
torch.cuda.synchronize()
print("ok", sig.shape)
