"""A/B of the sliced losses' sort: bucket pass + in-bucket rank (automatic) vs radix passes only, on circle coordinates of
projected clouds (the fused projection + sort entry point) and on uniform keys."""
import os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shwd
dev = torch.device("cuda:0")
lib = shwd._lib.lib()
def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
print("| rows x keys | input | radix ms | bucket ms | Gkeys/s radix / bucket |\n|---|---|---|---|---|")
for B, N, P in ((8, 4096, 512), (8, 1024, 512), (8, 2048, 512), (4, 8192, 256), (32, 256, 512)):
    g = torch.Generator().manual_seed(N)
    x = F.normalize(torch.randn(B, N, 3, generator=g), dim=-1).to(dev)
    fr, _ = torch.linalg.qr(torch.randn(P, 3, 2, generator=g)); fr = fr.contiguous().to(dev)
    so = torch.empty(B * P, N, device=dev); pe = torch.empty(B * P, N, device=dev, dtype=torch.int32)
    if N > lib.shwd_sort_projected_max_points():
        continue
    def proj():
        shwd._lib.check(lib.shwd_sort_projected(x.data_ptr(), fr.data_ptr(), B, N, P, 1, so.data_ptr(), pe.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream), "sort_projected")
    k = torch.rand(B * P, N, device=dev)
    res = {}
    for name, fn in (("projected sphere cloud", proj), ("uniform keys", lambda: shwd.ops._sort_i32(k))):
        t = []
        for method in (1, 0):
            lib.shwd_sort_set_method(method)
            t.append(timeit(fn))
        lib.shwd_sort_set_method(0)
        print("| %d x %d | %s | %.3f | %.3f | %.1f / %.1f |" % (B * P, N, name, t[0], t[1], B * P * N / t[0] * 1e-6, B * P * N / t[1] * 1e-6))
