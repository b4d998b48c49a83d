"""Achieved HBM bandwidth of the memory-bound stages (north_star: sphere-map, sort, Chamfer I/O; plus the sliced
projection): algorithmic bytes / CUDA-event time at sizes well above the 126 MB L2, against the measured copy
bandwidth in MEASURED_PEAKS.json.  Prints a markdown table.   python tools/bench_hbm_stages.py > gpurun_out/hbm_stages.md"""
import json, os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd
from shwd_b200 import _lib
lib = _lib.lib()
dev = torch.device("cuda:0")
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    src = "MEASURED_PEAKS.json"
except Exception:
    PEAK, src = 6552.3, "fallback 6552 GB/s"
p = lambda t: t.data_ptr()
s = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


print("peak: %.1f GB/s (%s)\n" % (PEAK, src))
print("| stage | shape | algorithmic bytes | ms | GB/s | frac of measured HBM peak |")
print("|---|---|---|---|---|---|")


def row(name, shape, nbytes, ms):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print("| %s | %s | %.1f MB | %.3f | %.0f | %.3f |" % (name, shape, nbytes / 1e6, ms, gbs, gbs / PEAK))
    sys.stdout.flush()


# sphere map: 12 B read + 16 B written per point
B, N = 16384, 1024
x = torch.randn(B, N, 3, device=dev)
x4 = torch.empty(B, N, 4, device=dev)
for flags, tag in ((3, "centre+normalise"), (2, "normalise")):
    ms = timed(lambda: _lib.check(lib.shwd_sphere_map_fwd(p(x), p(x4), None, B, N, flags, s), "map"))
    row("sphere_map_fwd (%s)" % tag, "B=%d N=%d" % (B, N), B * N * 28, ms)
g4 = torch.randn(B, N, 4, device=dev)
gx = torch.empty(B, N, 3, device=dev)
ms = timed(lambda: _lib.check(lib.shwd_sphere_map_bwd(p(x), p(x4), p(g4), None, p(gx), B, N, 3, s), "mapb"))
row("sphere_map_bwd", "B=%d N=%d" % (B, N), B * N * (12 + 16 + 16 + 12), ms)
del x, x4, g4, gx
# sliced projection: 12 B read per point per 32-frame tile, 4 B written per key
B, N, P = 8, 4096, 512
X = F.normalize(torch.randn(B, N, 3, device=dev), dim=-1)
U, _ = torch.linalg.qr(torch.randn(P, 3, 2, device=dev))
keys = torch.empty(B, P, N, device=dev)
ms = timed(lambda: _lib.check(lib.shwd_project_circle(p(X), p(U), B, N, P, p(keys), s), "proj"))
row("project_circle", "B=%d N=%d P=%d" % (B, N, P), B * P * N * 4 + B * N * 12 * (P // 32), ms)
# segmented sort: 4 B key read, 4 B sorted + 8 B perm written per key (cfg3 shape x 16 pairs)
S, Ln = 16 * 512, 4096
k = torch.rand(S, Ln, device=dev)
out = torch.empty_like(k)
perm = torch.empty(S, Ln, device=dev, dtype=torch.int64)
ms = timed(lambda: _lib.check(lib.shwd_segmented_sort(p(k), S, Ln, p(out), p(perm), None, 0, s), "sort"))
row("segmented_sort (stable, smem radix)", "%d segments x %d" % (S, Ln), S * Ln * 16, ms)
ms = timed(lambda: _lib.check(lib.shwd_segmented_sort(p(k), S, Ln, p(out), None, None, 0, s), "sort"))
row("segmented_sort (values only)", "%d segments x %d" % (S, Ln), S * Ln * 8, ms)
del k, out, perm
# Chamfer: 12 B read per point of both clouds, 8 B (distance + index) written per point
for B, N in ((4096, 1024), (32768, 128)):
    x = torch.randn(B, N, 3, device=dev); y = torch.randn(B, N, 3, device=dev)
    dxy = torch.empty(B, N, device=dev); dyx = torch.empty(B, N, device=dev)
    ixy = torch.empty(B, N, device=dev, dtype=torch.int32); iyx = torch.empty(B, N, device=dev, dtype=torch.int32)
    ms = timed(lambda: _lib.check(lib.shwd_chamfer_fwd(p(x), p(y), B, N, N, p(dxy), p(ixy), p(dyx), p(iyx), s), "cham"), 5)
    row("chamfer_fwd (I/O; %.0f Glane-op/s in the N^2 loop)" % (16.0 * B * N * N / (ms * 1e-3) / 1e9), "B=%d N=M=%d" % (B, N), B * 2 * N * 20, ms)
    del x, y, dxy, dyx, ixy, iyx
