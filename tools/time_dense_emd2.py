"""Single-pair exact solve on an explicit cost matrix (the ot.emd2 drop-in's call shape): ms and bidding rounds."""
import os, sys, time
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shwd
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1234)
for N in (256, 1024, 2048):
    x = F.normalize(torch.randn(N, 3, generator=g), dim=-1)
    y = F.normalize(torch.randn(N, 3, generator=g), dim=-1)
    M = torch.pow(torch.sum((x.unsqueeze(-2) - y.unsqueeze(-3)).abs() ** 2, -1), 0.5).to(dev)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sig, prices, rounds, status = shwd.exact_assignment_dense(M, return_info=True)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    val = M[torch.arange(N, device=dev), sig[0]].double().sum().item() / N
    print("N=%d  %.2f ms  rounds %d  status %d  value %.9f" % (N, min(ts), rounds.item(), status.item(), val))
