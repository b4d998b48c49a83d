#!/usr/bin/env python
"""Benchmark of the sphere-homeomorphic Wasserstein loss path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

metric   SHWD loss fwd+bwd pairs/s: one *step* = sphere map (centre + normalise) -> geodesic-cost entropic OT loss
         (p=2, eps=0.01, L=100 fixed iterations, mean over the batch) -> gradients w.r.t. both raw clouds, on a batch
         of B=32 synthetic registration pairs of N=1024 points (BASELINE config 2).
value    whole-job pairs/s with the inputs resident in HBM (CUDA events, max over ranks).
e2e      the same through the drop-in loss object (Geodesic_distance_W) from pinned HOST buffers: H2D of both clouds,
         loss + backward, D2H of the loss and both gradients, all inside the timed region.
scaling  weak: every rank owns its own B=32 batch (pairs are independent units); the only collective is
         the scalar loss all-reduce of the DDP step.
roofline FP32-issue roofline of the dominant kernel (the backward sweep kernel): algorithmic lane-ops of SURVEY.md 8(d)
         (33 per element-eval backward, 21 forward; E = (2L+1) N M element-evals per pair and direction) divided by the
         kernel's CUDA-event duration, against the FFMA issue rate measured live by shwd_peak_fp32.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, N_PTS, ITERS, EPS, P_COST = 32, 1024, 100, 0.01, 2.0
FWD_OPS, BWD_OPS = 21.0, 33.0  # algorithmic FP32 lane-ops per element-eval
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the two sweep kernels at this exact workload, from the
# `ncu --set full` capture summarised in profiles/ (bench.py cannot run under ncu while it is timing)
NCU_DRAM_BYTES_FWD, NCU_DRAM_BYTES_BWD = 11.97e6, 53.95e6
NCU_SOURCE = "profiles/r01e_ncu_sinkhorn_full_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
METRIC = "SHWD loss fwd+bwd pairs/s (B=32,N=1024)"
CONFIG = {"workload": "cfg2: synthetic registration pairs B=32/GPU N=M=1024, geodesic cost p=2, eps=0.01, L=100, sphere map "
                      "(centre+normalise) + loss + grads w.r.t. both clouds",
          "global_batch": None, "points": N_PTS, "sinkhorn_iters": ITERS, "eps": EPS,
          "parallelism": None, "l2": "flushed between timed steps (256 MiB write outside the timed events)"}


def registration_pairs(B, N, seed, device=None):
    """Synthetic ModelNet-shaped registration pairs (data_utils/Data_set_maker.py:154-171,
    train_W_COS.py:291-295): template = random surface cloud in the unit ball, source = rigid transform (Euler angles
    U(-45,45) deg, unit-norm translation) of it + N(0, 0.02^2) noise."""
    import torch
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(B, N, 3, generator=g)
    t = t / t.norm(dim=-1, keepdim=True) * (0.6 + 0.4 * torch.rand(B, N, 1, generator=g))  # bumpy star-shaped surface
    t = t * torch.tensor([1.0, 0.7, 0.5])
    ang = (torch.rand(B, 3, generator=g) * 2 - 1) * (math.pi / 4)
    cx, sx, cy, sy, cz, sz = ang[:, 0].cos(), ang[:, 0].sin(), ang[:, 1].cos(), ang[:, 1].sin(), ang[:, 2].cos(), ang[:, 2].sin()
    R = torch.zeros(B, 3, 3)
    R[:, 0, 0], R[:, 0, 1], R[:, 0, 2] = cy * cz, sx * sy * cz - cx * sz, cx * sy * cz + sx * sz
    R[:, 1, 0], R[:, 1, 1], R[:, 1, 2] = cy * sz, sx * sy * sz + cx * cz, cx * sy * sz - sx * cz
    R[:, 2, 0], R[:, 2, 1], R[:, 2, 2] = -sy, sx * cy, cx * cy
    tr = torch.randn(B, 1, 3, generator=g)
    tr = tr / tr.norm(dim=-1, keepdim=True)
    perm = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])
    s = torch.gather(t, 1, perm.unsqueeze(-1).expand(B, N, 3)) @ R.transpose(1, 2) + tr + 0.02 * torch.randn(B, N, 3, generator=g)
    if device is not None:
        t, s = t.to(device), s.to(device)
    return t.contiguous(), s.contiguous()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def cpu_reference_pairs_per_s(pairs, threads=None, repeats=1):
    """The reference's CPU torch loss path (oracle port of losses/Sinkhorn.py:25-60 on the geodesic cost matrix of
    s2_wasserstein.py:119-122, autograd backward) on `pairs` pairs of the benchmark workload."""
    import torch
    import oracle
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    t, s = registration_pairs(pairs, N_PTS, 1234)
    best = None
    for _ in range(repeats):
        x = t.clone().requires_grad_(True)
        y = s.clone().requires_grad_(True)
        t0 = time.perf_counter()
        xc = oracle.sphere_map(x, True, False)
        yc = oracle.sphere_map(y, True, False)
        loss = oracle.log_sinkhorn(xc, yc, "geodesic", P_COST, EPS, ITERS, None, "mean")
        loss.backward()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return pairs / best, threads, float(loss.item())


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; /root/reference cannot travel)."""
    if rank != 0:
        return
    sample_pairs = 2
    vals = []
    for i in range(args.warmup + args.steps):
        v, threads, _ = cpu_reference_pairs_per_s(sample_pairs)
        if i >= args.warmup:
            vals.append(v)
    value = len(vals) * sample_pairs / sum(sample_pairs / v for v in vals)
    cfg = dict(CONFIG, global_batch=B_PER_GPU * world, parallelism="cpu-threads=%d" % threads)
    sample = "%d pairs/step of the cfg2 workload (N=1024, L=100), fwd+bwd by torch autograd, float32" % sample_pairs
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sample_pairs / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import shwd
    from shwd_b200 import _lib
    lib = _lib.lib()

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, N = B_PER_GPU, N_PTS
    tmpl, src = registration_pairs(B, N, 1234 + rank, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    crit = shwd.losses.Geodesic_distance_W(device=dev, p=int(P_COST), eps=EPS, max_iter=ITERS)

    def step_resident():
        x = tmpl.detach().requires_grad_(True)
        y = src.detach().requires_grad_(True)
        res = shwd.entropic_ot(x, y, "geodesic", P_COST, EPS, ITERS, center=True)
        loss = res.cost.mean()
        loss.backward()
        return loss, x.grad, y.grad

    # ---- pinned host buffers for the end-to-end leg
    h_t, h_s = tmpl.cpu().pin_memory(), src.cpu().pin_memory()
    h_gx, h_gy = torch.empty_like(h_t).pin_memory(), torch.empty_like(h_s).pin_memory()
    h_loss = torch.empty(1).pin_memory()

    def step_e2e():
        x = h_t.to(dev, non_blocking=True)
        y = h_s.to(dev, non_blocking=True)
        x = x - x.mean(dim=1, keepdim=True)  # centring in the training loop, train_W_COS.py:167-168
        y = y - y.mean(dim=1, keepdim=True)
        x.requires_grad_(True)
        y.requires_grad_(True)
        loss = crit(x, y)
        loss.backward()
        h_loss.copy_(loss.detach().reshape(1), non_blocking=True)
        h_gx.copy_(x.grad, non_blocking=True)
        h_gy.copy_(y.grad, non_blocking=True)
        return loss

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        total = 0.0
        evs = []
        for _ in range(steps):
            flush.fill_(1)  # L2 flush, outside the timed events
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        total = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([total], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), out

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, out = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    loss_val = out[0].detach().clone()
    if world > 1:
        dist.all_reduce(loss_val, op=dist.ReduceOp.SUM)  # the DDP loss all-reduce: mean of equal-sized local means
        loss_val /= world
    ms_e2e, _ = timed(step_e2e, args.steps, args.warmup)
    status_ok = True

    # ---- per-kernel durations + live FP32 / MUFU peaks (rank 0)
    roof = None
    extra = {}
    if rank == 0:
        f32 = dict(device=dev, dtype=torch.float32)
        x4, y4 = torch.empty(B, N, 4, **f32), torch.empty(B, N, 4, **f32)
        HL = ITERS + 1
        alpha, beta = torch.empty(2, B, HL, N, **f32), torch.empty(2, B, HL, N, **f32)
        row_pc, col_pc, cost = torch.empty(B, N, **f32), torch.empty(B, N, **f32), torch.empty(B, **f32)
        it_run = torch.empty(1, device=dev, dtype=torch.int32)
        wsb = lib.shwd_sinkhorn_workspace_bytes(B, N, N, ITERS)
        ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
        gcost = torch.full((B,), 1.0 / B, **f32)
        g4x, g4y = torch.empty(B, N, 4, **f32), torch.empty(B, N, 4, **f32)
        s = torch.cuda.current_stream().cuda_stream
        p = lambda t: t.data_ptr()
        _lib.check(lib.shwd_sphere_map_fwd(p(tmpl), p(x4), None, B, N, 3, s), "map")
        _lib.check(lib.shwd_sphere_map_fwd(p(src), p(y4), None, B, N, 3, s), "map")

        def k_fwd():
            _lib.check(lib.shwd_sinkhorn_fwd(p(x4), p(y4), B, N, N, 0, P_COST, 1.0, EPS, ITERS, 0.0, HL, p(alpha), p(beta), p(row_pc),
                                             p(col_pc), p(cost), p(it_run), p(ws), wsb, s), "fwd")

        def k_bwd():
            _lib.check(lib.shwd_sinkhorn_bwd(p(x4), p(y4), B, N, N, 0, P_COST, 1.0, EPS, ITERS, p(alpha), p(beta), p(row_pc), p(col_pc),
                                             p(it_run), p(gcost), p(g4x), p(g4y), p(ws), wsb, s), "bwd")

        def ktime(fn, reps=5):
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return sum(ts) / len(ts)

        t_fwd, t_bwd = ktime(k_fwd), ktime(k_bwd)
        status_ok = int(ws[:4].view(torch.int32).item()) == 0
        scratch = torch.empty(148 * 4 * 512 * 2, **f32)
        ops = __import__("ctypes").c_double(0.0)

        def peak(fn):
            best = 0.0
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(fn(p(scratch), 2000, __import__("ctypes").byref(ops), s), "peak")
                e1.record()
                torch.cuda.synchronize()
                best = max(best, ops.value / (e0.elapsed_time(e1) * 1e-3))
            return best

        fp32_peak, mufu_peak = peak(lib.shwd_peak_fp32), peak(lib.shwd_peak_mufu)
        E = (2 * ITERS + 1) * N * N * B
        ach_b, ach_f = BWD_OPS * E / (t_bwd * 1e-3), FWD_OPS * E / (t_fwd * 1e-3)
        nominal = 148 * 128 * 1.965e9
        roof = {"bound": "fp32", "kernel": "sinkhorn_bwd_kernel<FAST_GEO2>", "achieved": ach_b / 1e12, "peak": fp32_peak / 1e12,
                "unit": "Tlane-op/s (FFMA = 1 lane-op)", "frac": ach_b / fp32_peak, "traffic": NCU_DRAM_BYTES_BWD,
                "traffic_source": NCU_SOURCE,
                "peak_source": "measured live: shwd_peak_fp32 FFMA chains (nominal 148x128x1.965 GHz = %.2f)" % (nominal / 1e12),
                "ms_per_launch": t_bwd, "algorithmic_ops_per_launch": BWD_OPS * E}
        extra = {"roofline_fwd": {"bound": "fp32", "kernel": "sinkhorn_fwd_kernel<FAST_GEO2>", "achieved": ach_f / 1e12,
                                  "peak": fp32_peak / 1e12, "unit": "Tlane-op/s (FFMA = 1 lane-op)", "frac": ach_f / fp32_peak,
                                  "ms_per_launch": t_fwd, "algorithmic_ops_per_launch": FWD_OPS * E,
                                  "traffic": NCU_DRAM_BYTES_FWD},
                 "roofline_sfu": {"note": "SURVEY.md 8(d) MUFU figure: 2 (fwd) + 3 (bwd) MUFU per element-eval against the MUFU "
                                          "issue rate measured live by shwd_peak_mufu; the backward executes 4 (sqrt, rsqrt, 2 ex2)",
                                  "frac_fwd": 2.0 * E / (t_fwd * 1e-3) / mufu_peak, "frac_bwd": 3.0 * E / (t_bwd * 1e-3) / mufu_peak,
                                  "frac_bwd_executed": 4.0 * E / (t_bwd * 1e-3) / mufu_peak},
                 "roofline_contract": {"note": "SURVEY.md 8(d): 54 lane-ops x E per pair (fwd+bwd) against the measured FP32 peak",
                                       "frac": (FWD_OPS + BWD_OPS) * E / ((t_fwd + t_bwd) * 1e-3) / fp32_peak,
                                       "frac_of_nominal": (FWD_OPS + BWD_OPS) * E / ((t_fwd + t_bwd) * 1e-3) / nominal},
                 "mufu_peak_gops": mufu_peak / 1e9, "fp32_peak_tlaneops": fp32_peak / 1e12}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # reported at N=1 only (torchrun pins OMP_NUM_THREADS=1)
        v, threads, _ = cpu_reference_pairs_per_s(4)
        cpu = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
               "sample": "4 pairs of the cfg2 workload (N=1024, L=100), one fwd+bwd by torch autograd on the host cores, float32"}

    if world > 1:
        dist.barrier()
    if rank == 0:
        pairs = B * world * args.steps
        value = pairs / (ms_total * 1e-3)
        e2e_v = pairs / (ms_e2e * 1e-3)
        cfg = dict(CONFIG, global_batch=B * world, parallelism="dp%d (batch-sharded, no data-path collective)" % world)
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": {"value": e2e_v, "unit": "pairs/s", "h2d_bytes_per_step": 2 * B * N * 12, "d2h_bytes_per_step": 2 * B * N * 12 + 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": 6 * args.steps, "roofline": roof, "cpu_baseline": cpu, "loss": float(loss_val.item()),
            "status_ok": status_ok,
        }
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
