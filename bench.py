#!/usr/bin/env python
"""Benchmark of the sphere-homeomorphic Wasserstein loss path (BASELINE.json metric and configs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg1..cfg5] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Default (what the driver runs): --config cfg2 --scaling weak.

cfg2     the headline metric, "SHWD loss fwd+bwd pairs/s (B=32,N=1024)": one *step* = sphere map (centre + normalise) ->
         geodesic-cost entropic OT loss (p=2, eps=0.01, L=100 fixed iterations, mean over the batch) -> gradients w.r.t.
         both raw clouds, on a batch of B=32 synthetic registration pairs of N=1024 points.
         --scaling weak: every rank owns its own B=32 batch.  --scaling strong: ONE global batch of 32 pairs split over
         the ranks (32/N pairs per GPU; the small-batch kernels of csrc/sinkhorn_lean.cu take over below 32).
cfg1     Comparison_.../main_rotation.py: per step, for a batch of 32 sphere-cloud pairs of N=1024, Chamfer
         (batch_reduction='sum') + log-Sinkhorn (eps=0.01, 100 iterations, 'sum', 'L2'), forward only as the script runs them.
cfg3     spherical sliced W (train_Pseudo_W_COS-style projected path): N=4096, 512 slices, p=2, 8 pairs per GPU, fwd+bwd.
cfg4     Wasserstein_flow_problem: one pair of N=16384 points, geodesic OT loss fwd+bwd + one Adam step on the evolving cloud.
cfg5     loss-kernel sweep N=256..16384 x B=1..256 (W_COS and Chamfer fwd+bwd); `value` is the B=32, N=256 point
         (the reference's own small runs), every point is listed under "sweep".
value    whole-job units/s with the inputs resident in HBM (CUDA events, max over ranks).
e2e      the same through the drop-in loss objects from pinned HOST buffers: H2D of both clouds, loss (+ backward), D2H of
         the loss and the gradients, all inside the timed region.
roofline of the dominant kernel: FP32-issue roofline for the OT sweeps (algorithmic lane-ops of SURVEY.md 8(d): 33 per
         element-eval backward, 21 forward; E = (2L+1) N M element-evals per pair and direction) against the FFMA issue
         rate measured live by shwd_peak_fp32; HBM roofline for the sliced path's segmented sort (12 B per key: key read,
         sorted value + int32 permutation written) against MEASURED_PEAKS.json's copy bandwidth.
"""
import argparse
import ctypes
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
SQE_FWD_OPS = 11.0             # squared-Euclidean cost (cfg1): 3 sub + 3 fma + affine + 4 online-LSE
CPU_CHUNK = 4                  # the CPU reference holds B=4 pairs of autograd tape at a time
NCU_BYTES_FILE = os.path.join(ROOT, "profiles", "ncu_dram_bytes.json")  # written by tools/ncu_summary.py --bytes


def ncu_bytes(cfg, kernel, default):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` (base name) at workload `cfg`, from the latest
    `ncu --set full` capture (bench.py cannot run under ncu while it is timing); written by tools/ncu_summary.py --bytes."""
    try:
        with open(NCU_BYTES_FILE) as fh:
            e = json.load(fh)[cfg][kernel]
        return float(e["bytes"]), "%s[%s][%s] (%s)" % (os.path.relpath(NCU_BYTES_FILE, ROOT), cfg, kernel, e.get("source", "?"))
    except Exception:
        return default, "profiles/r01e_ncu_sinkhorn_full_summary.txt (constant; no entry in profiles/ncu_dram_bytes.json)"


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "of fallback (B200_PROFILING.md: 6.65 TB/s)"


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


def sphere_pairs(B, N, seed, device=None, angle=2.0):
    """cfg1: unit-sphere clouds and their rotation about x (Comparison_.../Data_set_transformation.py:159)."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed)
    t = F.normalize(torch.randn(B, N, 3, generator=g), dim=-1)
    c, s_ = math.cos(angle), math.sin(angle)
    R = torch.tensor([[1.0, 0, 0], [0, c, -s_], [0, s_, c]])
    s = t @ R.T
    if device is not None:
        t, s = t.to(device), s.to(device)
    return t.contiguous(), s.contiguous()


def ellipsoid_pair(N, seed, device=None):
    """cfg4: ellipsoid (2,1,1) source and biased target (Flow_ellipsoid.ipynb:106-163)."""
    import torch
    g = torch.Generator().manual_seed(seed)

    def pts(theta, phi):
        return torch.stack([2 * torch.sin(theta) * torch.cos(phi), torch.sin(theta) * torch.sin(phi), torch.cos(theta)], -1)
    phi = torch.rand(N, generator=g) * 2 * math.pi
    theta = torch.acos(torch.rand(N, generator=g) * 2 - 1)
    src = pts(theta, phi)
    phi_t = torch.acos((torch.randn(N, generator=g) * 0.5).clamp(-1, 1)) * 2
    theta_t = torch.acos((torch.randn(N, generator=g) * 0.5).clamp(-1, 1))
    tgt = pts(theta_t, phi_t)
    if device is not None:
        src, tgt = src.to(device), tgt.to(device)
    return src.contiguous(), tgt.contiguous()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.mark_at = index, None, [], None

    def start(self):
        if os.environ.get("SHWD_BENCH_NO_CLOCKS") == "1":  # (diagnostic: is an outlier step the sampler's doing?)
            return
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

    def wait_first(self, timeout=5.0):
        """Block until the first sample has arrived: nvidia-smi's own start-up (driver attach, hundreds of ms) then lies before
        the warm-up instead of inside a 260 ms timed region -- one in four runs showed a 0.6 ms/step outlier on either leg."""
        t0 = time.time()
        while self.proc and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.02)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines[self.mark_at or 0:]:  # samples from the start of the first timed region on
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


# ---------------------------------------------------------------------------------------------------------------------
# CPU reference arm: the reference's torch loss path (oracle port) on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_step(config, units, threads=None):
    """One bounded sample of `config`'s step on the host: returns (seconds, units processed, loss)."""
    import torch
    import oracle
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    if config in ("cfg2", "cfg5"):
        N = N_PTS if config == "cfg2" else 256
        t, s = registration_pairs(units, N, 1234)
        x, y = t.clone().requires_grad_(True), s.clone().requires_grad_(True)
        t0 = time.perf_counter()
        loss = oracle.log_sinkhorn(oracle.sphere_map(x, True, False), oracle.sphere_map(y, True, False), "geodesic", P_COST, EPS, ITERS,
                                   None, "mean")
        loss.backward()
        return time.perf_counter() - t0, units, float(loss.item())
    if config == "cfg1":
        t, s = sphere_pairs(units, N_PTS, 1234)
        t0 = time.perf_counter()
        with torch.no_grad():
            cd, _ = oracle.chamfer_distance(s, t, "sum", "mean")
            sd = oracle.log_sinkhorn(t, s, "sqeuclid", 2, EPS, ITERS, 1e-9, "sum")
        return time.perf_counter() - t0, units, float(cd.item() + sd.item())
    if config == "cfg3":
        import torch.nn.functional as F
        g = torch.Generator().manual_seed(1234)
        x = F.normalize(torch.randn(units, 4096, 3, generator=g), dim=-1).requires_grad_(True)
        y = F.normalize(torch.randn(units, 4096, 3, generator=g) + 0.2, dim=-1).requires_grad_(True)
        U, _ = torch.linalg.qr(torch.randn(512, 3, 2, generator=g))
        t0 = time.perf_counter()
        loss = sum(oracle.sliced_wasserstein_sphere(x[i], y[i], U, p=2) for i in range(units)) / units
        loss.backward()
        return time.perf_counter() - t0, units, float(loss.item())
    if config == "cfg4":
        # the N x N tensors of the reference recurrence at N=16384 need ~1 GB each and L=100 levels of autograd tape:
        # a bounded sample runs the same step at N=4096 and is scaled by (16384/4096)^2 (the recurrence is O(N^2))
        src, tgt = ellipsoid_pair(4096, 1234)
        x = src.clone().requires_grad_(True)
        t0 = time.perf_counter()
        loss = oracle.log_sinkhorn(oracle.sphere_map(x[None], True, False), oracle.sphere_map(tgt[None], True, False), "geodesic", P_COST,
                                   EPS, 10, None, "mean")  # 10 of the 100 iterations
        loss.backward()
        dt = (time.perf_counter() - t0) * (ITERS / 10.0) * 16.0
        return dt, 1, float(loss.item())
    raise ValueError(config)


CPU_SAMPLE = {
    "cfg2": (CPU_CHUNK, "%d pairs per step = one of the 8 B=4 chunks BASELINE.md section 3 prescribes for the B=32 batch (the autograd "
                        "tape of all 32 would need ~34 GB); N=1024, L=100, fwd+bwd by torch autograd, float32" % CPU_CHUNK),
    "cfg1": (CPU_CHUNK, "%d pairs per step of the cfg1 batch: Chamfer + log-Sinkhorn (100 iterations) forward, float32" % CPU_CHUNK),
    "cfg3": (1, "1 pair per step of the 8-pair cfg3 batch: N=4096, 512 slices, p=2 bisection, fwd+bwd, float32"),
    "cfg4": (1, "one step at N=4096 with 10 of the 100 iterations, scaled by (100/10) x (16384/4096)^2 -- the full-size recurrence "
                "needs ~1 GB per N x N tensor and 200 of them on the tape"),
    "cfg5": (8, "8 pairs per step at the sweep's headline point B=32, N=256, L=100, fwd+bwd, float32"),
}


def run_reference(args, rank, world, meta):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; /root/reference cannot travel)."""
    if rank != 0:
        return
    units, sample = CPU_SAMPLE[args.config]
    tot_t, tot_u, threads = 0.0, 0, os.cpu_count() or 1
    for i in range(args.warmup + args.steps):
        dt, u, _ = cpu_reference_step(args.config, units, threads)
        if i >= args.warmup:
            tot_t += dt
            tot_u += u
    value = tot_u / tot_t
    print(json.dumps({
        "impl": "reference", "metric": meta["metric"], "value": value, "unit": meta["unit"], "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": meta["config"],
        "cpu_baseline": {"value": value, "unit": meta["unit"], "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": meta["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------------------------------
# workload descriptions (identical in both arms: the driver compares `config`)
# ---------------------------------------------------------------------------------------------------------------------
def workload_meta(config, scaling, world):
    l2 = "flushed between timed steps (256 MiB write outside the timed events)"
    if config == "cfg2":
        gb = B_PER_GPU * world if scaling == "weak" else B_PER_GPU
        return {"metric": "SHWD loss fwd+bwd pairs/s (B=32,N=1024)", "unit": "pairs/s",
                "config": {"workload": "cfg2: synthetic registration pairs B=32%s N=M=1024, geodesic cost p=2, eps=0.01, L=100, sphere "
                                       "map (centre+normalise) + loss + grads w.r.t. both clouds" % ("/GPU" if scaling == "weak" else " global"),
                           "global_batch": gb, "points": N_PTS, "sinkhorn_iters": ITERS, "eps": EPS,
                           "parallelism": "dp%d (batch-sharded, no data-path collective)" % world, "l2": l2}}
    if config == "cfg1":
        return {"metric": "main_rotation CD + log-Sinkhorn forward pairs/s (B=32,N=1024)", "unit": "pairs/s",
                "config": {"workload": "cfg1: Comparison_Wasserstein_with_Chamfer_distance/main_rotation.py, unit-sphere clouds rotated "
                                       "about x, B=32/GPU N=1024: chamfer_distance(batch_reduction='sum') + "
                                       "log_Sinkhorn_Distance_Loss(eps=0.01, max_iter=100, 'sum', 'L2'), forward only as the script does",
                           "global_batch": B_PER_GPU * world, "points": N_PTS, "sinkhorn_iters": ITERS, "eps": EPS,
                           "parallelism": "dp%d" % world, "l2": l2}}
    if config == "cfg3":
        return {"metric": "spherical sliced W loss fwd+bwd pairs/s (N=4096, 512 slices, p=2)", "unit": "pairs/s",
                "config": {"workload": "cfg3: sliced/projected path, N=M=4096, 512 great-circle slices, circular W_2^2 by bisection, "
                                       "8 pairs/GPU, loss + grads w.r.t. both clouds; resident leg = one CUDA-graph replay of fwd+bwd "
                                       "(shwd.graphed_loss), e2e leg eager", "global_batch": 8 * world, "points": 4096,
                           "slices": 512, "p": 2, "parallelism": "dp%d (batch-sharded)" % world, "l2": l2}}
    if config == "cfg4":
        return {"metric": "geodesic OT gradient-flow steps/s (one pair, N=16384)", "unit": "steps/s",
                "config": {"workload": "cfg4: Wasserstein_flow_problem ellipsoid flow, one pair N=M=16384, geodesic OT loss (p=2, eps=0.01, "
                                       "L=100) fwd+bwd + Adam step on the evolving cloud; no N x N tensor exists", "global_batch": world,
                           "points": 16384, "sinkhorn_iters": ITERS, "eps": EPS, "parallelism": "replicas x%d" % world, "l2": l2}}
    if config == "cfg5":
        return {"metric": "loss-kernel sweep: W_COS fwd+bwd pairs/s at B=32, N=256 (all points under 'sweep')", "unit": "pairs/s",
                "config": {"workload": "cfg5: W_COS (geodesic OT, L=100) and Chamfer fwd+bwd over N=256..16384 x B=1..256",
                           "global_batch": 32 * world, "points": 256, "sinkhorn_iters": ITERS, "eps": EPS,
                           "parallelism": "dp%d" % world, "l2": l2}}
    raise ValueError(config)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.scaling == "strong" and args.config != "cfg2":
        raise SystemExit("--scaling strong is defined for cfg2 (one global batch of 32 pairs)")
    meta = workload_meta(args.config, args.scaling, world)
    if args.impl == "reference":
        run_reference(args, rank, world, meta)
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

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    f32 = dict(device=dev, dtype=torch.float32)
    p = lambda t: t.data_ptr()
    cur = lambda: torch.cuda.current_stream().cuda_stream
    L = shwd.losses

    def pinned(t):
        return t.detach().cpu().pin_memory()

    # ---- per-config steps ----------------------------------------------------------------------------------------------
    cfg = args.config
    launches = 0
    sweep = None
    if cfg in ("cfg2", "cfg5"):
        if cfg == "cfg2":
            if args.scaling == "strong":
                if B_PER_GPU % world:
                    raise SystemExit("strong scaling splits 32 pairs: --gpus must divide 32")
                B = B_PER_GPU // world
                t_all, s_all = registration_pairs(B_PER_GPU, N_PTS, 1234)  # ONE global batch; this rank's contiguous shard
                tmpl, src = t_all[rank * B:(rank + 1) * B].to(dev), s_all[rank * B:(rank + 1) * B].to(dev)
            else:
                B = B_PER_GPU
                tmpl, src = registration_pairs(B, N_PTS, 1234 + rank, dev)
            N = N_PTS
        else:
            B, N = 32, 256
            tmpl, src = registration_pairs(B, N, 1234 + rank, dev)
        units_local = B
        crit = L.Geodesic_distance_W(device=dev, p=int(P_COST), eps=EPS, max_iter=ITERS)
        # end-to-end staging: both clouds travel as ONE pinned buffer / one H2D copy, both gradients and the loss come back as ONE
        # D2H copy (five separate DMA operations per step left the leg exposed to per-copy latency on a busy box)
        h_in = pinned(torch.stack([tmpl, src]))
        h_back = torch.empty(2 * B * N * 3 + 1).pin_memory()
        launches = 6  # 2 sphere-map fwd, OT fwd, OT bwd, 2 sphere-map bwd
        h2d, d2h = 2 * B * N * 12, 2 * B * N * 12 + 4

        def step_resident():
            x = tmpl.detach().requires_grad_(True)
            y = src.detach().requires_grad_(True)
            res = shwd.entropic_ot(x, y, "geodesic", P_COST, EPS, ITERS, center=True)
            loss = res.cost.mean()
            loss.backward()
            return loss

        def step_e2e():
            xy = h_in.to(dev, non_blocking=True)
            x, y = xy[0], xy[1]
            x = x - x.mean(dim=1, keepdim=True)  # centring in the training loop, train_W_COS.py:167-168
            y = y - y.mean(dim=1, keepdim=True)
            x.requires_grad_(True)
            y.requires_grad_(True)
            loss = crit(x, y)
            loss.backward()
            h_back.copy_(torch.cat([x.grad.reshape(-1), y.grad.reshape(-1), loss.detach().reshape(1)]), non_blocking=True)
            return loss
    elif cfg == "cfg1":
        B, N = B_PER_GPU, N_PTS
        tmpl, src = sphere_pairs(B, N, 1234 + rank, dev)
        units_local = B
        sk = L.log_Sinkhorn_Distance_Loss(eps=EPS, max_iter=ITERS, batch_reduction="sum", type_of_cost_norm="L2", dense_outputs=False)
        h_t, h_s, h_out = pinned(tmpl), pinned(src), torch.empty(2).pin_memory()
        launches = 5  # chamfer fwd + its reduction, 2 sphere-map (packing), OT fwd
        h2d, d2h = 2 * B * N * 12, 8

        def eval_pair(t, s):
            with torch.no_grad():
                cd = L.chamfer_distance(s, t, batch_reduction="sum")[0]   # main_rotation.py:203
                sd, _, _ = sk(t, s, dev)                                    # main_rotation.py:207-210
            return cd, sd

        def step_resident():
            cd, sd = eval_pair(tmpl, src)
            return cd + sd

        def step_e2e():
            t = h_t.to(dev, non_blocking=True)
            s = h_s.to(dev, non_blocking=True)
            cd, sd = eval_pair(t, s)
            h_out.copy_(torch.stack([cd, sd]), non_blocking=True)
            return cd + sd
    elif cfg == "cfg3":
        import torch.nn.functional as F
        B, N, P = 8, 4096, 512
        g = torch.Generator().manual_seed(1234 + rank)
        xs = F.normalize(torch.randn(B, N, 3, generator=g), dim=-1).to(dev)
        ys = F.normalize(torch.randn(B, N, 3, generator=g) + 0.2, dim=-1).to(dev)
        U0, _ = torch.linalg.qr(torch.randn(P, 3, 2, generator=g))
        U0 = U0.to(dev)
        units_local = B
        h_x, h_y = pinned(xs), pinned(ys)
        h_gx, h_gy, h_loss = torch.empty_like(h_x).pin_memory(), torch.empty_like(h_y).pin_memory(), torch.empty(1).pin_memory()
        launches = 5  # 2 project+sort (the sort CTAs compute their own keys), circular_wp, 2 project-bwd
        h2d, d2h = 2 * B * N * 12, 2 * B * N * 12 + 4

        # The step is ~0.9 ms of kernels behind ~30 Python-issued launches: eager, its time is the host's on a busy box (one
        # 1-GPU run measured 2.9 ms for the same kernels).  The resident leg therefore replays the whole forward + backward
        # as ONE CUDA graph through the package's public shwd.graphed_loss (graphs.py); the end-to-end leg below stays eager
        # (it draws its frames like the reference, randn + QR, inside the call).
        cfg3_graph = shwd.graphed_loss(lambda x, y: L.sliced_cost(x, y, U0, p=2).mean())
        xs_g, ys_g = xs.detach().requires_grad_(True), ys.detach().requires_grad_(True)

        def step_resident():
            loss, _grads = cfg3_graph.value_and_grad(xs_g, ys_g)
            return loss

        def step_e2e():
            x = h_x.to(dev, non_blocking=True).requires_grad_(True)
            y = h_y.to(dev, non_blocking=True).requires_grad_(True)
            loss = L.sliced_wasserstein_sphere(x, y, P, dev, p=2).mean()  # draws its frames like the reference (randn + QR)
            loss.backward()
            h_loss.copy_(loss.detach().reshape(1), non_blocking=True)
            h_gx.copy_(x.grad, non_blocking=True)
            h_gy.copy_(y.grad, non_blocking=True)
            return loss
    elif cfg == "cfg4":
        N = 16384
        src, tgt = ellipsoid_pair(N, 1234 + rank, dev)
        units_local = 1
        crit = L.Geodesic_distance_W(device=dev, p=int(P_COST), eps=EPS, max_iter=ITERS)
        evolving = src.clone().requires_grad_(True)
        opt = torch.optim.Adam([evolving], lr=0.01)
        h_src, h_tgt = pinned(src), pinned(tgt)
        h_out, h_loss = torch.empty_like(h_src).pin_memory(), torch.empty(1).pin_memory()
        launches = 7  # 2 sphere-map fwd, OT fwd, OT bwd, 2 sphere-map bwd, fused Adam
        h2d, d2h = 2 * N * 12, N * 12 + 4

        def flow_step(x, y, optim):
            optim.zero_grad(set_to_none=True)
            loss = crit(x, y)           # un-batched (N,3) inputs -> the x.dim()==2 branch, s2_wasserstein.py:31-32,46-48
            loss.backward()
            optim.step()
            return loss

        def step_resident():
            return flow_step(evolving, tgt, opt)

        def step_e2e():
            x = h_src.to(dev, non_blocking=True).requires_grad_(True)
            y = h_tgt.to(dev, non_blocking=True)
            o = torch.optim.Adam([x], lr=0.01)
            loss = flow_step(x, y, o)
            h_loss.copy_(loss.detach().reshape(1), non_blocking=True)
            h_out.copy_(x.detach(), non_blocking=True)
            return loss

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        if rank == 0 and sampler.mark_at is None:
            sampler.mark_at = len(sampler.lines)
        evs = []
        for _ in range(steps):
            flush.fill_(1)  # L2 flush, outside the timed events
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), out

    # clocks: started BEFORE the warm-up and killed AFTER both timed legs (B200_PROFILING.md: "start before, kill after") -- the
    # sampler's start-up and exit are process events on the same GPU and must not fall into a timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    ms_total, out = timed(step_resident, args.steps, args.warmup)
    loss_val = out.detach().clone().reshape(())
    if world > 1:
        dist.all_reduce(loss_val, op=dist.ReduceOp.SUM)  # the DDP loss all-reduce (after the timed region)
        loss_val /= world
    ms_e2e, _ = timed(step_e2e, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    status_ok = True

    # ---- roofline of the dominant kernel + live peaks (rank 0) -------------------------------------------------------
    roof, extra = None, {}

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

    def live_peaks():
        scratch = torch.empty(148 * 4 * 512 * 2, **f32)
        ops = ctypes.c_double(0.0)

        def peak(fn):
            best = 0.0
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(fn(p(scratch), 2000, ctypes.byref(ops), cur()), "peak")
                e1.record()
                torch.cuda.synchronize()
                best = max(best, ops.value / (e0.elapsed_time(e1) * 1e-3))
            return best
        return peak(lib.shwd_peak_fp32), peak(lib.shwd_peak_mufu)

    def ot_kernel_times(xa, ya, B, N, kind, want_bwd=True, thresh=0.0, center=True):
        """CUDA-event durations of the two OT launches alone (through the C ABI on preallocated buffers)."""
        x4, y4 = torch.empty(B, N, 4, **f32), torch.empty(B, N, 4, **f32)
        HL = ITERS + 1
        alpha, beta = torch.empty(2, B, HL, N, **f32), torch.empty(2, B, HL, N, **f32)
        row_pc, col_pc, cost = torch.empty(B, N, **f32), torch.empty(B, N, **f32), torch.empty(B, **f32)
        it_run = torch.empty(1, device=dev, dtype=torch.int32)
        wsb = lib.shwd_sinkhorn_workspace_bytes(B, N, N, ITERS)
        ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
        gcost = torch.full((B,), 1.0 / B, **f32)
        g4x, g4y = torch.empty(B, N, 4, **f32), torch.empty(B, N, 4, **f32)
        flags = (1 if center else 0) | (2 if kind == 0 else 0)
        _lib.check(lib.shwd_sphere_map_fwd(p(xa), p(x4), None, B, N, flags, cur()), "map")
        _lib.check(lib.shwd_sphere_map_fwd(p(ya), p(y4), None, B, N, flags, cur()), "map")

        def k_fwd():
            _lib.check(lib.shwd_sinkhorn_fwd(p(x4), p(y4), B, N, N, kind, P_COST, 1.0, EPS, ITERS, thresh, HL, p(alpha), p(beta), p(row_pc),
                                             p(col_pc), p(cost), p(it_run), p(ws), wsb, cur()), "fwd")

        def k_bwd():
            _lib.check(lib.shwd_sinkhorn_bwd(p(x4), p(y4), B, N, N, kind, P_COST, 1.0, EPS, ITERS, p(alpha), p(beta), p(row_pc), p(col_pc),
                                             p(it_run), p(gcost), p(g4x), p(g4y), p(ws), wsb, cur()), "bwd")
        t_f = ktime(k_fwd)
        t_b = ktime(k_bwd) if want_bwd else None
        ok = int(ws[:4].view(torch.int32).item()) == 0
        return t_f, t_b, ok

    if rank == 0:
        nominal = 148 * 128 * 1.965e9
        if cfg in ("cfg2", "cfg4", "cfg5", "cfg1"):
            fp32_peak, mufu_peak = live_peaks()
            psrc = "measured live: shwd_peak_fp32 FFMA chains (nominal 148x128x1.965 GHz = %.2f)" % (nominal / 1e12)
        if cfg in ("cfg2", "cfg5", "cfg4"):
            Bk, Nk = (1, 16384) if cfg == "cfg4" else (B, N)
            xa, ya = (src[None].contiguous(), tgt[None].contiguous()) if cfg == "cfg4" else (tmpl, src)
            t_fwd, t_bwd, status_ok = ot_kernel_times(xa, ya, Bk, Nk, 0)
            lean = bool(lib.shwd_sinkhorn_lean_regime(Bk, Nk, Nk))
            kname = "sinkhorn_bwd_lean_kernel<FAST_GEO2>" if lean else "sinkhorn_bwd_kernel<FAST_GEO2>"
            E = (2 * ITERS + 1) * Nk * Nk * Bk
            ach_b, ach_f = BWD_OPS * E / (t_bwd * 1e-3), FWD_OPS * E / (t_fwd * 1e-3)
            kbase = "sinkhorn_bwd_lean_kernel" if lean else "sinkhorn_bwd_kernel"
            tb_bytes, tb_src = ncu_bytes(cfg, kbase, 53.95e6 if cfg == "cfg2" and not lean else None)
            tf_bytes, _ = ncu_bytes(cfg, kbase.replace("bwd", "fwd"), 11.97e6 if cfg == "cfg2" and not lean else None)
            roof = {"bound": "fp32", "kernel": kname, "achieved": ach_b / 1e12, "peak": fp32_peak / 1e12,
                    "unit": "Tlane-op/s (FFMA = 1 lane-op)", "frac": ach_b / fp32_peak, "traffic": tb_bytes, "traffic_source": tb_src,
                    "peak_source": psrc, "ms_per_launch": t_bwd, "algorithmic_ops_per_launch": BWD_OPS * E}
            extra = {"roofline_fwd": {"bound": "fp32", "kernel": kname.replace("bwd", "fwd"), "achieved": ach_f / 1e12,
                                      "peak": fp32_peak / 1e12, "unit": "Tlane-op/s (FFMA = 1 lane-op)", "frac": ach_f / fp32_peak,
                                      "ms_per_launch": t_fwd, "algorithmic_ops_per_launch": FWD_OPS * E, "traffic": tf_bytes},
                     "roofline_sfu": {"note": "SURVEY.md 8(d) MUFU figure: 2 (fwd) + 3 (bwd) MUFU per element-eval against the MUFU "
                                              "issue rate measured live by shwd_peak_mufu; the backward executes 4 (sqrt, rsqrt, 2 ex2)",
                                      "frac_fwd": 2.0 * E / (t_fwd * 1e-3) / mufu_peak, "frac_bwd": 3.0 * E / (t_bwd * 1e-3) / mufu_peak,
                                      "frac_bwd_executed": 4.0 * E / (t_bwd * 1e-3) / mufu_peak},
                     "roofline_contract": {"note": "SURVEY.md 8(d): 54 lane-ops x E per pair (fwd+bwd) against the measured FP32 peak",
                                           "frac": (FWD_OPS + BWD_OPS) * E / ((t_fwd + t_bwd) * 1e-3) / fp32_peak,
                                           "frac_of_nominal": (FWD_OPS + BWD_OPS) * E / ((t_fwd + t_bwd) * 1e-3) / nominal},
                     "mufu_peak_gops": mufu_peak / 1e9, "fp32_peak_tlaneops": fp32_peak / 1e12}
        elif cfg == "cfg1":
            t_fwd, _, status_ok = ot_kernel_times(tmpl, src, B, N, 1, want_bwd=False, thresh=1e-9, center=False)
            E = (2 * ITERS + 1) * N * N * B
            ach = SQE_FWD_OPS * E / (t_fwd * 1e-3)
            roof = {"bound": "fp32", "kernel": "sinkhorn_fwd_kernel<FAST_SQE2>", "achieved": ach / 1e12, "peak": fp32_peak / 1e12,
                    "unit": "Tlane-op/s (FFMA = 1 lane-op)", "frac": ach / fp32_peak, "traffic": None, "peak_source": psrc,
                    "ms_per_launch": t_fwd, "algorithmic_ops_per_launch": SQE_FWD_OPS * E,
                    "note": "squared-Euclidean cost: 11 algorithmic lane-ops + 1 MUFU per element-eval; the kernel is XU (ex2) bound: "
                            "frac_mufu = %.3f of the measured MUFU issue rate" % (1.0 * E / (t_fwd * 1e-3) / mufu_peak)}
        elif cfg == "cfg3":
            S, n = 2 * B * P, N  # both clouds' slices in the step; one sort launch per cloud
            keys = torch.rand(B * P, n, **f32)
            outk = torch.empty_like(keys)
            perm = torch.empty(B * P, n, device=dev, dtype=torch.int32)
            wsb = lib.shwd_segmented_sort_workspace_bytes(B * P, n)
            ws = torch.empty(max(wsb, 8), device=dev, dtype=torch.uint8)

            def k_sort():
                _lib.check(lib.shwd_segmented_sort_i32(p(keys), B * P, n, p(outk), p(perm), p(ws), wsb, cur()), "sort")
            t_sort = ktime(k_sort)
            hbm, hsrc = measured_hbm_peak()
            bytes_alg = 12.0 * B * P * n
            tr_bytes, tr_src = ncu_bytes("cfg3", "segmented_sort_trim_kernel", None)
            roof = {"bound": "hbm", "kernel": "segmented_sort_trim_kernel (one launch per cloud: 4096 slices x 4096 keys; in the loss the "
                                              "sort CTAs also compute the keys, shwd_sort_projected)",
                    "achieved": bytes_alg / (t_sort * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": bytes_alg / (t_sort * 1e-3) / 1e9 / hbm,
                    "traffic": tr_bytes, "traffic_source": tr_src, "peak_source": hsrc, "ms_per_launch": t_sort,
                    "algorithmic_bytes_per_launch": bytes_alg, "gkeys_per_s": B * P * n / (t_sort * 1e-3) / 1e9,
                    "note": "SURVEY.md 8(d) unfused contract: 16 P (n+m) B per pair over the whole step = %.0f GB/s of %.0f" % (
                        16.0 * P * 2 * N * B / (ms_total / args.steps * 1e-3) / 1e9, hbm)}

    # ---- cfg5: the sweep itself (rank 0, after the timed headline point) ---------------------------------------------
    if cfg == "cfg5" and rank == 0:
        sweep = []
        for Ns in (256, 1024, 4096, 16384):
            for Bs in (1, 4, 32, 256):
                work = (2 * ITERS + 1) * Ns * Ns * Bs
                if work > 6.0e11:
                    continue
                xa, ya = registration_pairs(Bs, Ns, Ns + Bs, dev)
                xa.requires_grad_(True)
                ya.requires_grad_(True)

                def w_step():
                    xa.grad = ya.grad = None
                    shwd.entropic_ot(xa, ya, "geodesic", P_COST, EPS, ITERS, center=True).cost.mean().backward()

                def c_step():
                    xa.grad = ya.grad = None
                    L.chamfer_distance(xa, ya)[0].backward()
                ms_w, ms_c = ktime(w_step, 3), ktime(c_step, 3)
                sweep.append({"B": Bs, "N": Ns, "wcos_ms": ms_w, "wcos_pairs_per_s": Bs / (ms_w * 1e-3),
                              "wcos_frac_fp32_contract": 54.0 * work / (ms_w * 1e-3) / fp32_peak,
                              "lean_kernels": bool(lib.shwd_sinkhorn_lean_regime(Bs, Ns, Ns)),
                              "chamfer_ms": ms_c, "chamfer_pairs_per_s": Bs / (ms_c * 1e-3)})

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # reported at N=1 only (torchrun pins OMP_NUM_THREADS=1)
        units, sample = CPU_SAMPLE[cfg]
        dt, u, _ = cpu_reference_step(cfg, units)
        cpu = {"value": u / dt, "unit": meta["unit"], "cores": os.cpu_count() or 1, "kind": "port", "sample": sample}

    if world > 1:
        dist.barrier()
    if rank == 0:
        units = units_local * world * args.steps
        line = {
            "metric": meta["metric"], "value": units / (ms_total * 1e-3), "unit": meta["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": meta["config"], "clocks": clocks,
            "e2e": {"value": units / (ms_e2e * 1e-3), "unit": meta["unit"], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches * args.steps, "roofline": roof, "cpu_baseline": cpu, "loss": float(loss_val.item()),
            "status_ok": status_ok,
        }
        line.update(extra)
        if sweep is not None:
            line["sweep"] = sweep
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
