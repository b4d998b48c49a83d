"""The segmented sort against torch.sort (CUB segmented radix sort) on the same B200, same keys: ms per call and Gkeys/s
for the sliced path's row shapes.  `fused` = shwd_sort_projected (the sort CTAs compute their keys from
the cloud: to be compared with projection + sort).   python tools/time_sort_baseline.py > gpurun_out/sort_baseline.md"""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd
dev = torch.device("cuda:0")
lib = shwd._lib.lib()


def timed(fn, reps=7):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


print("| rows x keys | torch.sort (values+int64 idx) ms | torch.sort stable ms | ours int32 ms | Gkeys/s torch / ours | projection + ours ms | fused project+sort ms |")
print("|---|---|---|---|---|---|---|")
for B, P, N in ((8, 512, 4096), (1, 512, 4096), (8, 512, 1024), (1, 512, 16384), (8, 512, 2048)):
    g = torch.Generator().manual_seed(N)
    x = F.normalize(torch.randn(B, N, 3, generator=g), dim=-1).to(dev)
    U, _ = torch.linalg.qr(torch.randn(P, 3, 2, generator=g))
    U = U.contiguous().to(dev)
    keys = shwd.ops.ProjectCircleFn.apply(x, U).reshape(B * P, N).contiguous()
    t_torch = timed(lambda: torch.sort(keys, dim=-1))
    t_stable = timed(lambda: torch.sort(keys, dim=-1, stable=True))
    t_ours = timed(lambda: shwd.ops._sort_i32(keys))
    t_ps = timed(lambda: shwd.ops._sort_i32(shwd.ops.ProjectCircleFn.apply(x, U).reshape(B * P, N)))
    so = torch.empty(B * P, N, device=dev)
    pe = torch.empty(B * P, N, device=dev, dtype=torch.int32)
    if N <= lib.shwd_sort_projected_max_points():
        st = torch.cuda.current_stream().cuda_stream
        t_f = timed(lambda: shwd._lib.check(lib.shwd_sort_projected(x.data_ptr(), U.data_ptr(), B, N, P, 1, so.data_ptr(), pe.data_ptr(), st), "f"))
        tf = "%.3f" % t_f
    else:
        tf = "-"
    nk = B * P * N
    print("| %d x %d | %.3f | %.3f | %.3f | %.1f / %.1f | %.3f | %s |" % (B * P, N, t_torch, t_stable, t_ours, nk / t_torch * 1e-6, nk / t_ours * 1e-6, t_ps, tf))
    sys.stdout.flush()
