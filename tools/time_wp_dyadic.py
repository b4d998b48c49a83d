"""A/B of the closed-form (dyadic) searches of circular_wp for equal power-of-two cloud sizes: kernel-only time at cfg3 size
(4096 slices of 4096 + 4096 keys, p = 2) and the fused cfg3 loss fwd+bwd, shortcut on / off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import shwd
dev = torch.device("cuda:0")
lib = shwd._lib.lib()


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for n in (1024, 4096, 16384, 1000, 4000):
    S = 4096 if n <= 4096 else 512
    g = torch.Generator().manual_seed(11)
    B = 8 if n <= 4096 else 1
    x = F.normalize(torch.randn(B, n, 3, generator=g), dim=-1).to(dev)
    y = F.normalize(torch.randn(B, n, 3, generator=g) + 0.2, dim=-1).to(dev)
    U, _ = torch.linalg.qr(torch.randn(S // B, 3, 2, generator=g)); U = U.to(dev)
    ku = shwd.ops.ProjectCircleFn.apply(x, U).reshape(S, n)
    kv = shwd.ops.ProjectCircleFn.apply(y, U).reshape(S, n)
    us, vs = torch.sort(ku, -1)[0].contiguous(), torch.sort(kv, -1)[0].contiguous()
    xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)

    def step():
        xg.grad = yg.grad = None
        shwd.ops.spherical_sliced_wp(xg, yg, U, 2.0).sum().backward()

    for on in (0, 1):
        lib.shwd_circular_wp_set_dyadic(on)
        tk = timed(lambda: shwd.ops.CircularWpFn.apply(us, vs, 2.0, -1.0, 1.0, 1e-7))
        ts = timed(step)
        print("n = m = %d, %d slices, dyadic %s: circular_wp kernel (value + gradients) %.1f us, fused sliced loss fwd+bwd %.1f us"
              % (n, S, "on" if on else "off", tk, ts))
lib.shwd_circular_wp_set_dyadic(1)
