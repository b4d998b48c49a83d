"""cfg3 (BASELINE.json configs[2]): spherical sliced W, N = 4096, 512 slices, batch sharded over the GPUs of one box.

    python tools/bench_cfg3.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/bench_cfg3.py

Weak scaling like bench.py: every rank owns PAIRS_PER_GPU pairs of the global batch (shwd_b200.dist.sharded_pair_loss: no
collective on the data path, one 2-vector all-reduce for the global mean).  A step = loss forward + backward w.r.t. both clouds.
Timed with CUDA events, barrier + synchronize on both sides, max over ranks; one JSON line per p from rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd  # noqa: E402
from shwd_b200 import dist as sdist  # noqa: E402

PAIRS_PER_GPU, N, P, STEPS, WARMUP = 8, 4096, 512, 20, 3
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
g = torch.Generator().manual_seed(1234)
Bg = PAIRS_PER_GPU * world
x = F.normalize(torch.randn(Bg, N, 3, generator=g), dim=-1).to(dev)          # the same global batch on every rank;
y = F.normalize(torch.randn(Bg, N, 3, generator=g) + 0.2, dim=-1).to(dev)    # each rank evaluates its own shard
U, _ = torch.linalg.qr(torch.randn(P, 3, 2, generator=g))
U = U.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for p in (1.0, 2.0):
    fn = (lambda a, b: shwd.ops.spherical_sliced_w1(a, b, U)) if p == 1.0 else (lambda a, b: shwd.ops.spherical_sliced_wp(a, b, U, p))

    def step():
        xs, ys = x.detach().requires_grad_(True), y.detach().requires_grad_(True)
        loss = sdist.sharded_pair_loss(fn, xs, ys, rank, world)
        loss.backward()
        return loss

    import time
    t_end = time.perf_counter() + 0.5  # at least WARMUP steps AND half a second: the clocks of an idle GPU ramp up first
    k = 0
    while k < WARMUP or time.perf_counter() < t_end:
        step()
        k += 1
        if k % 16 == 0:
            torch.cuda.synchronize()
    barrier()
    evs = []
    for _ in range(STEPS):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = t.item() / STEPS
        print(json.dumps({"metric": "spherical sliced W loss fwd+bwd pairs/s (N=4096, 512 slices)", "p": p, "value": Bg / (ms * 1e-3),
                          "unit": "pairs/s", "n_gpus": world, "ms_per_step": ms, "steps": STEPS, "warmup": WARMUP, "scaling": "weak",
                          "config": {"workload": "cfg3", "pairs_per_gpu": PAIRS_PER_GPU, "points": N, "slices": P,
                                     "l2": "flushed between timed steps"}, "loss": float(loss.item())}))
if world > 1:
    dist.destroy_process_group()
