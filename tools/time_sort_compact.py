"""A/B of the compact-layout sort (rows of 4225 .. 16384 keys): bucket pass + in-bucket rank vs the four radix passes, on circle
coordinates of a projected cloud (what the sliced losses sort at cfg4 size) and on uniform keys; torch.sort beside it."""
import os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shwd
dev = torch.device("cuda:0")
lib = shwd._lib.lib()


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


print("| rows x keys | input | torch.sort ms | radix ms | bucket ms | Gkeys/s radix / bucket |\n|---|---|---|---|---|---|")
for rows, n in ((512, 16384), (1024, 8192), (2048, 5000), (512, 10000)):
    g = torch.Generator().manual_seed(n)
    x = F.normalize(torch.randn(1, n, 3, generator=g), dim=-1).to(dev)
    U, _ = torch.linalg.qr(torch.randn(rows, 3, 2, generator=g))
    kc = shwd.ops.ProjectCircleFn.apply(x, U.to(dev)).reshape(rows, n).contiguous()
    ku = torch.rand(rows, n, device=dev)
    for name, k in (("projected sphere cloud", kc), ("uniform keys", ku)):
        t = []
        for method in (1, 0):
            lib.shwd_sort_set_method(method)
            t.append(timeit(lambda: shwd.ops._sort_i32(k)))
        lib.shwd_sort_set_method(0)
        tt = timeit(lambda: torch.sort(k, dim=-1, stable=True))
        print("| %d x %d | %s | %.3f | %.3f | %.3f | %.1f / %.1f |" % (rows, n, name, tt, t[0], t[1], rows * n / t[0] * 1e-6, rows * n / t[1] * 1e-6))
