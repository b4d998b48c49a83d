"""Phase timing of the lean Sinkhorn kernels (diagnostic; needs a library built with -DSHWD_PROFILE, e.g.
tools/build_variant.sh prof -DSHWD_PROFILE, then SHWD_B200_LIB=tools/variants/prof.so python tools/phase_profile_lean.py BxN ...).
Thread 0 of every CTA accumulates clock64 deltas per phase; printed as cycles per half-step per CTA."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd  # noqa: E402
lib = shwd._lib.lib()
raw = ctypes.CDLL(shwd._lib.LIB_PATH)
raw.shwd_prof_read_lean.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
dev = torch.device("cuda:0")
L = 100
buf = (ctypes.c_ulonglong * 8)()
names_f = ["poll", "barrier1", "compute", "or-barrier", "log2+publish", "loop head", "barrier2"]
names_b = ["poll", "PRE stage", "compute", "barrier2", "merge+publish", "loop head", "barrier1"]
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(1, 256), (1, 1024), (4, 1024), (32, 256)]
lib.shwd_sinkhorn_set_path(2)
for B, N in shapes:
    torch.manual_seed(0)
    x = torch.nn.functional.normalize(torch.randn(B, N, 3), dim=-1).to(dev).requires_grad_(True)
    y = torch.nn.functional.normalize(torch.randn(B, N, 3), dim=-1).to(dev).requires_grad_(True)
    for it in range(2):
        raw.shwd_prof_read_lean(buf, 1)
        res = shwd.entropic_ot(x, y, "geodesic", 2.0, 0.01, L, center=True)
        torch.cuda.synchronize()
        raw.shwd_prof_read_lean(buf, 1)
        f = list(buf)[:7]
        res.cost.sum().backward()
        torch.cuda.synchronize()
        raw.shwd_prof_read_lean(buf, 1)
        b = list(buf)[:7]
    import math
    G = 148
    q = max(1, G // B)
    for tag, v, names in (("fwd", f, names_f), ("bwd", b, names_b)):
        tot = sum(v)
        print("B=%d N=%d %s: total thread-0 cycles over all CTAs %.2fM | " % (B, N, tag, tot / 1e6) +
              "  ".join("%s %.1f%%" % (n, 100 * c / max(tot, 1)) for n, c in zip(names, v)))
lib.shwd_sinkhorn_set_path(0)
