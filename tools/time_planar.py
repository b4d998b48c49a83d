"""Planar sphere map (csrc/planar.cu) against the eager PlanarFlow modules on one B200: forward + backward, CUDA events.
    gpurun -- python tools/time_planar.py > gpurun_out/<tag>_planar.md
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shwd  # noqa: E402


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3  # us


def graphed(fn, reps=200):
    """Device time of ``fn`` without the host: captured once in a CUDA graph, replayed."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return timed(g.replay, reps)


def main():
    dev = torch.device("cuda:0")
    print("fwd+bwd = phi(x), (y * w).sum(), autograd.grad w.r.t. x and every parameter; 'graph' = the same step replayed as a CUDA")
    print("graph (device time without the Python / autograd host work)\n")
    print("| input | layers | eager fwd+bwd us | fused fwd+bwd us | eager graph us | fused graph us | fused fwd only us | max abs(y - y_eager) |")
    print("|---|---|---|---|---|---|---|---|")
    for shape, layers in (((32, 1024, 3), 3), ((32, 1024, 3), 5), ((256, 256, 3), 3), ((1, 16384, 3), 3), ((1, 262144, 3), 3),
                          ((1024, 3), 3), ((65536, 3), 3), ((1048576, 3), 3), ((1048576, 3), 8)):
        torch.manual_seed(0)
        phi = shwd.losses.Norm_Flow_structure(flow_name="Planar", n_flow_layer=layers).to(dev)
        x = torch.randn(*shape, device=dev)
        if len(shape) == 3:  # (uncentred clouds with small weights: the per-layer amplification N u^ w stays moderate, see planar.cu)
            with torch.no_grad():
                for f in phi.net:
                    f.w.mul_(1.0 / shape[1])
        x.requires_grad_(True)
        w = torch.randn(*shape, device=dev)

        def step(f):
            y = f(x)
            torch.autograd.grad((y * w).sum(), [x] + list(phi.parameters()))

        def fwd_only():
            with torch.no_grad():
                phi(x)

        te, tf, t0 = timed(lambda: step(phi.forward_eager)), timed(lambda: step(phi)), timed(fwd_only)
        ge, gf = graphed(lambda: step(phi.forward_eager)), graphed(lambda: step(phi))
        err = (phi(x) - phi.forward_eager(x)).abs().max().item()
        print("| %s | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1e |" % ("x".join(map(str, shape)), layers, te, tf, ge, gf, t0, err))


if __name__ == "__main__":
    main()
