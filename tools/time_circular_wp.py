"""Kernel-only time of circular_wp at cfg3 size (4096 slices of 4096 + 4096 keys) as a function of the stopping width:
the slope between two widths is the cost of the bisection rounds in between (early rounds: full dCost; late rounds:
searches only, partial sums reused from a bracket end)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import shwd
dev = torch.device("cuda:0")
S = int(os.environ.get("S", 4096))
for n, m in ((4096, 4096), (4000, 4000), (4096, 3500)):
    g = torch.Generator().manual_seed(11)
    x = F.normalize(torch.randn(8, n, 3, generator=g), dim=-1).to(dev)
    y = F.normalize(torch.randn(8, m, 3, generator=g) + 0.2, dim=-1).to(dev)
    U, _ = torch.linalg.qr(torch.randn(S // 8, 3, 2, generator=g)); U = U.to(dev)
    ku = shwd.ops.ProjectCircleFn.apply(x, U).reshape(S, n)
    kv = shwd.ops.ProjectCircleFn.apply(y, U).reshape(S, m)
    us, vs = torch.sort(ku, -1)[0].contiguous(), torch.sort(kv, -1)[0].contiguous()
    prev = None
    for tol in (1.5, 2.0 ** -5, 2.0 ** -11, 2.0 ** -17, 1e-7):
        for _ in range(2):
            shwd.ops.CircularWpFn.apply(us, vs, 2.0, -1.0, 1.0, tol)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            w, th = shwd.ops.CircularWpFn.apply(us, vs, 2.0, -1.0, 1.0, tol)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 5 * 1e3
        print("n=%d m=%d tol=%.3g: %.1f us%s   mean W %.6e" % (n, m, tol, t, "" if prev is None else "  (+%.1f)" % (t - prev), w.mean().item()))
        prev = t
