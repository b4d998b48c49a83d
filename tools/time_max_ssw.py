"""max-SSW wrapper (max_spherical_sliced_w.py:498-536) on one B200: the reference's call order (one SSW call per pair) against the
one-fused-call evaluation with per-pair frames (shwd_*_pp); one 'test' evaluation and one training call (max_iter ascent steps).
    gpurun -- python tools/time_max_ssw.py > gpurun_out/<tag>_max_ssw.md
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shwd  # noqa: E402

L = shwd.losses


def per_pair_ssw(a, b, P, device, p=2):  # a distinct callable: the wrapper then loops over the pairs like the reference
    return L.sliced_wasserstein_sphere(a, b, P, device, p=p)


def wall(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    dev = torch.device("cuda:0")
    print("| B | N | slices | p | per-pair calls: eval ms | one fused call: eval ms | per-pair: train call ms | fused: train call ms |")
    print("|---|---|---|---|---|---|---|---|")
    for B, N, P, p in ((32, 1024, 100, 2), (32, 1024, 100, 1), (8, 4096, 512, 2), (32, 128, 100, 2)):
        torch.manual_seed(0)
        first = torch.randn(B, N, 3, device=dev)
        second = (first * 0.9 + 0.1 * torch.randn(B, N, 3, device=dev)).requires_grad_(True)
        row = []
        for ssw in (per_pair_ssw, L.sliced_wasserstein_sphere):
            phi = L.transform_to_sphere().to(dev)
            op = torch.optim.SGD(phi.parameters(), lr=1e-3)
            crit = L.max_spherical_wassersten_distance(P, phi, ssw, op, p=p, max_iter=3, device=dev, verbose=False)

            def evaluate():
                v, _, _ = crit(first, second, "test")
                v.backward()

            def train():
                v, _, _ = crit(first, second, "train")
                v.backward()

            row.append((wall(evaluate), wall(train, 3)))
        print("| %d | %d | %d | %d | %.2f | %.2f | %.2f | %.2f |" % (B, N, P, p, row[0][0], row[1][0], row[0][1], row[1][1]))


if __name__ == "__main__":
    main()
