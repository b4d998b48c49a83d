"""Phase timing of the persistent Sinkhorn kernels (diagnostic; needs a library built with -DSHWD_PROFILE):
    python <pkg>/build.py --force --profile && python tools/phase_profile.py
Prints the share of CTA time spent waiting on other CTAs, staging, computing, finalising and signalling."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd  # noqa: E402

lib = shwd._lib.lib()
raw = ctypes.CDLL(shwd._lib.LIB_PATH)
raw.shwd_prof_read.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
dev = torch.device("cuda:0")
B, N, L = 32, 1024, 100
torch.manual_seed(0)
x = torch.nn.functional.normalize(torch.randn(B, N, 3), dim=-1).to(dev).requires_grad_(True)
y = torch.nn.functional.normalize(torch.randn(B, N, 3), dim=-1).to(dev).requires_grad_(True)
buf = (ctypes.c_ulonglong * 8)()
names = ["wait", "pre-stage", "compute", "finalize", "signal", "item-setup", "post-stage"]
for it in range(2):
    res = shwd.entropic_ot(x, y, "geodesic", 2.0, 0.01, L, center=True)
    torch.cuda.synchronize()
    raw.shwd_prof_read(buf, 1)
    f = list(buf)[:7]
    res.cost.sum().backward()
    torch.cuda.synchronize()
    raw.shwd_prof_read(buf, 1)
    b = list(buf)[:7]
    for tag, v in (("fwd", f), ("bwd", b)):
        tot = sum(v)
        print(tag, "cycles/CTA %.2fM  " % (tot / 148 / 1e6) + "  ".join("%s %.1f%%" % (n, 100 * c / tot) for n, c in zip(names, v)))
