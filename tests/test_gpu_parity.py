"""GPU parity tests: the CUDA path (through the C ABI / the drop-in losses) against the CPU oracle and the golden
fixtures frozen from the reference.  Tolerances (BASELINE.json north_star): 1e-5 relative in fp32 for floating-point
outputs, bit-exact for sort permutations and nearest-neighbour indices.

"relative" is norm-wise: |a - b|_2 / |b|_2 (a per-element bound is meaningless for gradients that cross zero).
The north-star family (geodesic cost, p = 2) is held to 1e-5 flat.  For the other cost kinds, where cost/eps is large
and the reference's own float32 evaluation sits up to 2e-5 from its float64 evaluation, the bound is
max(1e-5, 8 x that distance), stated per test.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5


@pytest.fixture(scope="module")
def shwd():
    import shwd as m
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    m._lib.lib()  # fail loudly if the CUDA library is missing
    return m


def dev():
    return torch.device("cuda:0")


def rel(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def gold(name):
    return dict(np.load(os.path.join(G, name + ".npz"), allow_pickle=False))


# ----------------------------------------------------------------------------------------------------- sphere map ---
# (1, 65536), (2, 8193), (3, 16384): few large clouds -> one thread-block cluster of 8 CTAs per cloud (partial sums through DSMEM)
@pytest.mark.parametrize("B,N", [(3, 64), (2, 50), (4, 1024), (1, 7), (1, 65536), (2, 8193), (3, 16384)])
@pytest.mark.parametrize("center", [True, False])
@pytest.mark.parametrize("normalize", [True, False])
def test_sphere_map_fwd_bwd(shwd, B, N, center, normalize):
    torch.manual_seed(B * 1000 + N)
    x = torch.randn(B, N, 3) * 2 + 0.3
    xr = x.clone().requires_grad_(True)
    ref = oracle.sphere_map(xr, center, normalize)
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    xg = x.clone().to(dev()).requires_grad_(True)
    out = shwd.sphere_map(xg, center, normalize)
    (out * w.to(dev())).sum().backward()
    assert rel(out, ref) < 2e-7
    assert rel(xg.grad, xr.grad) < TOL
    if not center:  # no reduction involved: the normalisation itself is bit-exact
        assert torch.equal(out.cpu(), ref.detach())


def test_sphere_map_unbatched_and_noncontiguous(shwd):
    torch.manual_seed(1)
    x = torch.randn(5, 3, 40)
    xv = x.transpose(1, 2)  # (5,40,3) non-contiguous view
    out = shwd.sphere_map(xv.to(dev()))
    assert rel(out, oracle.sphere_map(xv)) < 2e-7
    out1 = shwd.sphere_map(xv[0].to(dev()))
    assert out1.shape == (40, 3) and rel(out1, oracle.sphere_map(xv[0])) < 2e-7


def test_flow_regularizer_matches_reference(shwd):
    d = gold("regularizer")
    x = torch.from_numpy(d["x"]).to(dev()).requires_grad_(True)
    r = shwd.flow_regularization(x)
    r.backward()
    assert r.item() == pytest.approx(float(d["reg"]), rel=1e-6)
    assert rel(x.grad, torch.from_numpy(d["gx"])) < TOL


# ------------------------------------------------------------------------------------------------------ Sinkhorn ----
def _run_cuda(shwd, x, y, kind, p, eps, iters, n_power=1.0, thresh=0.0):
    xg = x.clone().to(dev()).requires_grad_(True)
    yg = y.clone().to(dev()).requires_grad_(True)
    res = shwd.entropic_ot(xg, yg, kind, p, eps, iters, n_power, thresh)
    cost = res.cost if n_power == 1.0 else res.cost.pow(1.0 / n_power)
    cost.sum().backward()
    assert res.status() == 0, "an inter-CTA wait timed out"
    return cost.detach().cpu(), xg.grad.cpu(), yg.grad.cpu(), res


def _run_oracle(x, y, kind, p, eps, iters, n_power=1.0, thresh=0.0, dtype=torch.float32):
    xr = x.clone().to(dtype).requires_grad_(True)
    yr = y.clone().to(dtype).requires_grad_(True)
    c = oracle.log_sinkhorn(xr, yr, kind, p, eps, iters, thresh if thresh > 0 else None, "none", n_power)
    c.sum().backward()
    return c.detach().reshape(-1), xr.grad, yr.grad


# the north-star configuration family: geodesic cost, p = 2
GEO_FIXTURES = ["geodesic_sinkhorn_p2", "geodesic_sinkhorn_p2_ragged"]


@pytest.mark.parametrize("name", GEO_FIXTURES)
def test_geodesic_sinkhorn_matches_reference_fixture(shwd, name):
    """Loss and gradients against the frozen output of the reference's Sinkhorn recurrence (losses/Sinkhorn.py:35-60)
    run on the reference's geodesic cost matrix (s2_wasserstein.py:119-122)."""
    d = gold(name)
    x, y = torch.from_numpy(d["x"]), torch.from_numpy(d["y"])
    red = str(d["batch_reduction"])
    scale = x.shape[0] if red == "mean" else 1.0
    cost, gx, gy, _ = _run_cuda(shwd, x, y, "geodesic", 2.0, float(d["eps"]), int(d["max_iter"]))
    ref_loss = torch.from_numpy(np.atleast_1d(d["loss"])).double().sum() * scale
    assert abs(cost.double().sum().item() - ref_loss.item()) / abs(ref_loss.item()) < TOL
    assert rel(gx, torch.from_numpy(d["gx"]) * scale) < TOL
    assert rel(gy, torch.from_numpy(d["gy"]) * scale) < TOL


@pytest.mark.parametrize("B,N,M,L,eps", [(1, 33, 70, 10, 0.05), (5, 256, 256, 20, 0.01), (2, 300, 1000, 15, 0.02),
                                          (3, 2500, 700, 6, 0.05), (2, 512, 512, 100, 0.01)])
def test_geodesic_sinkhorn_matches_oracle(shwd, B, N, M, L, eps):
    torch.manual_seed(B * 7 + N)
    x = F.normalize(torch.randn(B, N, 3), dim=-1) * (1 + 0.2 * torch.rand(B, N, 1))  # not unit norm on entry
    y = F.normalize(torch.randn(B, M, 3) + 0.3, dim=-1)
    cost, gx, gy, _ = _run_cuda(shwd, x, y, "geodesic", 2.0, eps, L)
    c_ref, gx_ref, gy_ref = _run_oracle(x, y, "geodesic", 2, eps, L)
    assert torch.isfinite(cost).all() and torch.isfinite(gx).all() and torch.isfinite(gy).all()
    # the reference returns NaN for a pair as soon as one cosine rounds to >= 1 (acos / its derivative);
    # parity is defined on the pairs where the reference itself is finite
    ok = torch.isfinite(c_ref) & torch.isfinite(gx_ref).flatten(1).all(1) & torch.isfinite(gy_ref).flatten(1).all(1)
    assert ok.any()
    assert rel(cost[ok], c_ref[ok]) < TOL
    assert rel(gx[ok], gx_ref[ok]) < TOL
    assert rel(gy[ok], gy_ref[ok]) < TOL


# ---- the two kernel families behind shwd_sinkhorn_fwd / _bwd (include/shwd.h: shwd_sinkhorn_set_path) -----------------
PATHS = {"auto": 0, "flat": 1, "lean": 2}


class _path:
    """Force the flattened-deal (1) or the lean small-problem (2) kernels for the enclosed calls."""

    def __init__(self, shwd, name):
        self.lib, self.mode = shwd._lib.lib(), PATHS[name]

    def __enter__(self):
        assert self.lib.shwd_sinkhorn_set_path(self.mode) == 0

    def __exit__(self, *a):
        self.lib.shwd_sinkhorn_set_path(0)


LEAN_SHAPES = [
    # B, N, M, L, eps                 what the dedicated-CTA mapping of csrc/sinkhorn_lean.cu looks like on 148 SMs
    (1, 1024, 1024, 100, 0.01),     # one pair over 128 CTAs, 8 owners per warp row (4 sub-slices per warp)
    (1, 256, 256, 40, 0.01),        # 32 CTAs, heavy padding of the packed records
    (4, 1024, 1024, 100, 0.01),     # the 4 pairs per GPU of an 8-way strong-scaled B=32: 32 CTAs per pair, 32 owners each
    (32, 256, 256, 100, 0.01),      # train_RUNNER.py's small runs: 4 CTAs per pair, two owner groups each
    (32, 1024, 1024, 12, 0.02),     # the benchmark shape forced onto the lean kernels: 4 CTAs per pair, eight groups each
    (200, 256, 200, 12, 0.02),      # B > #SMs: whole pairs per CTA, 8 + 7 owner groups, 4 group pairs x 4 slices
    (3, 40, 700, 15, 0.05),         # ragged: most CTAs have no row owners at all
    (2, 2048, 1500, 8, 0.05),       # the largest resident clouds
    (9, 600, 600, 20, 0.02),        # 16 CTAs per pair, 38 owners -> two groups of 32 with a ragged tail
    (7, 1000, 777, 10, 0.02),       # N != M: bands of 64 row owners and 64 column owners, the last ones ragged
    (16, 1000, 1024, 10, 0.02),     # 9 CTAs per pair, four groups
]


@pytest.mark.parametrize("B,N,M,L,eps", LEAN_SHAPES)
def test_lean_kernels_match_oracle_and_flat_kernels(shwd, B, N, M, L, eps):
    """The dedicated-CTA kernels for small problems (csrc/sinkhorn_lean.cu) against the float32 oracle (1e-5), and
    against the flattened-deal kernels they replace there (same arithmetic, different summation grouping)."""
    torch.manual_seed(B * 31 + N + M)
    x = F.normalize(torch.randn(B, N, 3), dim=-1) * (1 + 0.2 * torch.rand(B, N, 1))
    y = F.normalize(torch.randn(B, M, 3) + 0.3, dim=-1)
    with _path(shwd, "lean"):
        assert shwd._lib.lib().shwd_sinkhorn_lean_regime(B, N, M) == 1
        cost, gx, gy, _ = _run_cuda(shwd, x, y, "geodesic", 2.0, eps, L)
    with _path(shwd, "flat"):
        cost_f, gx_f, gy_f, _ = _run_cuda(shwd, x, y, "geodesic", 2.0, eps, L)
    assert torch.isfinite(cost).all() and torch.isfinite(gx).all() and torch.isfinite(gy).all()
    print("lean vs flat: cost %.2e gx %.2e gy %.2e" % (rel(cost, cost_f), rel(gx, gx_f), rel(gy, gy_f)))
    assert rel(cost, cost_f) < 2e-6 and rel(gx, gx_f) < TOL and rel(gy, gy_f) < TOL
    if B * N * M * L > 6e8:  # the CPU oracle's autograd tape: keep it to a few pairs
        keep = slice(0, max(1, int(6e8 // (N * M * L))))
        x, y, cost, gx, gy = x[keep], y[keep], cost[keep], gx[keep], gy[keep]
    c_ref, gx_ref, gy_ref = _run_oracle(x, y, "geodesic", 2, eps, L)
    ok = torch.isfinite(c_ref) & torch.isfinite(gx_ref).flatten(1).all(1) & torch.isfinite(gy_ref).flatten(1).all(1)
    if not ok.any():  # float32 reference NaN on every pair kept (a cosine rounded to >= 1): use its float64 run
        c_ref, gx_ref, gy_ref = _run_oracle(x, y, "geodesic", 2, eps, L, dtype=torch.float64)
        ok = torch.isfinite(c_ref) & torch.isfinite(gx_ref).flatten(1).all(1) & torch.isfinite(gy_ref).flatten(1).all(1)
    assert ok.any()
    errs = (rel(cost[ok], c_ref[ok]), rel(gx[ok], gx_ref[ok]), rel(gy[ok], gy_ref[ok]))
    print("lean vs oracle: cost %.2e gx %.2e gy %.2e" % errs)
    assert max(errs) < TOL


@pytest.mark.parametrize("kind,p", [("geodesic", 1), ("sqeuclid", 2), ("sqeuclid", 1), ("euclid", 2), ("one_minus_cos", 2)])
@pytest.mark.parametrize("B,N,M", [(1, 700, 512), (12, 256, 300)])
def test_lean_kernels_other_packed_costs(shwd, kind, p, B, N, M):
    """Every packed cost kind through the lean kernels, against the reference's float32 / float64 evaluations."""
    torch.manual_seed(B + N)
    x = F.normalize(torch.randn(B, N, 3), dim=-1)
    y = F.normalize(torch.randn(B, M, 3) + 0.3, dim=-1)
    with _path(shwd, "lean"):
        cost, gx, gy, _ = _run_cuda(shwd, x, y, kind, float(p), 0.02, 30)
    c32, gx32, gy32 = _run_oracle(x, y, kind, p, 0.02, 30)
    c64, gx64, gy64 = _run_oracle(x, y, kind, p, 0.02, 30, dtype=torch.float64)
    ok = torch.isfinite(c32) & torch.isfinite(gx32).flatten(1).all(1) & torch.isfinite(gy32).flatten(1).all(1)
    assert ok.any() and torch.isfinite(cost).all() and torch.isfinite(gx).all() and torch.isfinite(gy).all()
    for a, r32, r64 in ((cost, c32, c64), (gx, gx32, gx64), (gy, gy32, gy64)):
        floor = rel(r32[ok], r64[ok])
        print("%s p=%s: err %.2e (float32 reference's own floor %.2e)" % (kind, p, rel(a[ok], r32[ok]), floor))
        assert rel(a[ok], r32[ok]) < max(TOL, 8 * floor)


def test_lean_forward_without_grad_and_reproducibility(shwd):
    """Forward-only calls keep the history in the lean regime (the exchange planes); two runs are bit-identical."""
    torch.manual_seed(12)
    x = F.normalize(torch.randn(4, 512, 3), dim=-1).to(dev())
    y = F.normalize(torch.randn(4, 512, 3) + 0.2, dim=-1).to(dev())
    with torch.no_grad():
        a = shwd.entropic_ot(x, y, "geodesic", 2.0, 0.01, 60)
        b = shwd.entropic_ot(x, y, "geodesic", 2.0, 0.01, 60)
    assert a.status() == 0 and torch.equal(a.cost, b.cost)
    ref = oracle.log_sinkhorn(x.cpu(), y.cpu(), "geodesic", 2, 0.01, 60)
    assert rel(a.cost, ref) < TOL
    xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    g = []
    for _ in range(2):
        xg.grad = yg.grad = None
        shwd.entropic_ot(xg, yg, "geodesic", 2.0, 0.01, 60).cost.sum().backward()
        g.append((xg.grad.clone(), yg.grad.clone()))
    assert torch.equal(g[0][0], g[1][0]) and torch.equal(g[0][1], g[1][1])


OTHER = [
    # fixture, kind, p, n_power, thresh
    ("geodesic_sinkhorn_p1", "geodesic", 1, 1, 0.0),
    ("sinkhorn_cmp_L2", "sqeuclid", 2, 1, 1e-9),
    ("sinkhorn_cmp_L1_sum", "sqeuclid", 1, 1, 1e-9),
    ("sinkhorn_plain_L2_none", "sqeuclid", 2, 1, 0.0),
    ("sinkhorn_fixed_L2", "euclid", 2, 1, 1e-9),
    ("sinkhorn_logN_2", "sqeuclid", 2, 2, 1e-9),
]


@pytest.mark.parametrize("name,kind,p,n_power,thresh", OTHER)
def test_other_cost_kinds_match_reference_within_its_own_noise(shwd, name, kind, p, n_power, thresh):
    """Every other cost kind of the reference's Sinkhorn classes.  On these fixtures cost/eps reaches 1e2..1e4, the
    gradient is a C/eps-fold cancellation and the reference's OWN float32 autograd sits 2e-6..2e-5 from
    its float64 evaluation (measured here as `floor`).  Two independent float32 evaluations cannot agree better than a
    small multiple of that, so the bound is max(1e-5, 8 x floor) -- against the float32 reference AND against its
    float64 evaluation (the CUDA gradient is the exact gradient of its own rounded forward: see sinkhorn.cu)."""
    d = gold(name)
    x, y = torch.from_numpy(d["x"]), torch.from_numpy(d["y"])
    eps, iters = float(d["eps"]), int(d["max_iter"])
    cost, gx, gy, _ = _run_cuda(shwd, x, y, kind, float(p), eps, iters, float(n_power), thresh)
    c32, gx32, gy32 = _run_oracle(x, y, kind, p, eps, iters, n_power, thresh)
    c64, gx64, gy64 = _run_oracle(x, y, kind, p, eps, iters, n_power, thresh, dtype=torch.float64)
    assert rel(cost, c32) < max(TOL, 8 * rel(c32, c64))
    for g, g32, g64 in ((gx, gx32, gx64), (gy, gy32, gy64)):
        floor = rel(g32, g64)
        assert rel(g, g32) < max(TOL, 8 * floor), (rel(g, g32), floor)
        assert rel(g, g64) < max(TOL, 8 * floor), (rel(g, g64), floor)


@pytest.mark.parametrize("kind,p,eps", [("one_minus_cos", 2, 0.05), ("geodesic", 1.5, 0.05), ("euclid", 1, 0.1),
                                        ("sqeuclid", 3, 0.1), ("one_minus_cos", 1, 0.05)])
def test_generic_cost_path(shwd, kind, p, eps):
    torch.manual_seed(3)
    x = torch.randn(2, 100, 3)
    y = torch.randn(2, 90, 3)
    cost, gx, gy, _ = _run_cuda(shwd, x, y, kind, float(p), eps, 20)
    c32, gx32, gy32 = _run_oracle(x, y, kind, p, eps, 20)
    c64, gx64, gy64 = _run_oracle(x, y, kind, p, eps, 20, dtype=torch.float64)
    assert rel(cost, c32) < max(TOL, 8 * rel(c32, c64))
    assert rel(gx, gx32) < max(TOL, 8 * rel(gx32, gx64))
    assert rel(gy, gy32) < max(TOL, 8 * rel(gy32, gy64))


@pytest.mark.parametrize("kind,p,B,N,M,L,eps", [("geodesic", 1, 3, 300, 260, 30, 0.02), ("geodesic", 1, 2, 1024, 1024, 50, 0.01),
                                                ("sqeuclid", 1, 3, 300, 260, 30, 0.02), ("euclid", 1, 2, 1024, 1024, 50, 0.02),
                                                ("sqeuclid", 1, 2, 700, 2300, 10, 0.05),
                                                ("euclid", 2, 3, 300, 260, 30, 0.05), ("euclid", 2, 2, 1024, 1024, 50, 0.02),
                                                ("one_minus_cos", 2, 3, 300, 260, 30, 0.02),
                                                ("one_minus_cos", 2, 2, 1024, 1024, 50, 0.01)])
def test_other_fast_costs_match_oracle(shwd, kind, p, B, N, M, L, eps):
    """The packed fast paths for geodesic p = 1 (default p of Geodesic_distance_W, s2_wasserstein.py:74), the L1 cost
    (Cos_disimilarity_W(p=1), train_W1_COS.py:393; 'L1' of the Sinkhorn classes), |x-y|_2 (Sinkhorn_fixed.py:79-89) and
    (1-cos)^2 (max_spherical_w_cos_with_regulation.py:745) on resident / chunked / ragged shapes.
    Some points of y coincide with points of x: |d| has the sub-gradient sign(0) = 0 there (torch.abs backward)."""
    torch.manual_seed(B * 13 + N)
    x = F.normalize(torch.randn(B, N, 3), dim=-1)
    y = F.normalize(torch.randn(B, M, 3) + 0.3, dim=-1)
    if kind != "geodesic" and p == 1:  # exact coincidences (geodesic / |d|_2: the reference returns NaN there)
        y[:, :16] = x[:, :16]
        y[:, 20:24, 0] = x[:, 20:24, 0]
    cost, gx, gy, _ = _run_cuda(shwd, x, y, kind, float(p), eps, L)
    c32, gx32, gy32 = _run_oracle(x, y, kind, p, eps, L)
    c64, gx64, gy64 = _run_oracle(x, y, kind, p, eps, L, dtype=torch.float64)
    ok = torch.isfinite(c32) & torch.isfinite(gx32).flatten(1).all(1) & torch.isfinite(gy32).flatten(1).all(1)
    assert ok.all() or kind == "geodesic"
    assert ok.any() and torch.isfinite(cost).all() and torch.isfinite(gx).all() and torch.isfinite(gy).all()
    assert rel(cost[ok], c32[ok]) < max(TOL, 8 * rel(c32[ok], c64[ok]))
    assert rel(gx[ok], gx32[ok]) < max(TOL, 8 * rel(gx32[ok], gx64[ok]))
    assert rel(gy[ok], gy32[ok]) < max(TOL, 8 * rel(gy32[ok], gy64[ok]))


def test_known_answer_sinkhorn_fixed_smoke(shwd):
    """Point_Cloud_Resistration/losses/Sinkhorn_fixed.py:97-110: integer grids, eps 0.1, 10 iterations."""
    d = gold("sinkhorn_fixed_smoke")
    a, b = torch.from_numpy(d["a"]).to(dev()), torch.from_numpy(d["b"]).to(dev())
    L = shwd.losses
    for norm, want in (("L1", 8.621342658996582), ("L2", 8.161666870117188)):
        crit = L.log_Sinkhorn_Distance_Loss_fixed(eps=0.1, max_iter=10, batch_reduction="mean", type_of_cost_norm=norm)
        loss, P, C = crit(a, b, dev())
        assert loss.item() == pytest.approx(want, rel=TOL)
        assert P.shape == (2, 9, 12) and C.shape == (2, 9, 12)
        assert P.sum(dim=(1, 2)).cpu().numpy() == pytest.approx(np.ones(2), rel=1e-2)


def test_dense_plan_and_cost_match_reference_outputs(shwd):
    d = gold("sinkhorn_cmp_L2")
    x, y = torch.from_numpy(d["x"]).to(dev()), torch.from_numpy(d["y"]).to(dev())
    crit = shwd.losses.log_Sinkhorn_Distance_Loss(eps=float(d["eps"]), max_iter=int(d["max_iter"]), batch_reduction="mean")
    loss, P, C = crit(x, y, dev())
    assert rel(C, torch.from_numpy(d["C"])) < 1e-6
    assert rel(P, torch.from_numpy(d["P"])) < 5e-3  # plan entries inherit the ~1e-4 exponent noise of float32 (cost/eps ~ 1e3)
    assert loss.item() == pytest.approx(float(d["loss"]), rel=TOL)


def test_early_stop_selects_the_reference_iterate(shwd):
    """eps large enough to reach a float32 fixed point: the device-side rule must stop where the reference stops."""
    torch.manual_seed(0)
    x = torch.rand(2, 24, 3)
    y = torch.rand(2, 20, 3)
    xr, yr = x.clone(), y.clone()
    c, P, C, u, v, iters_ref = oracle.log_sinkhorn(xr, yr, "sqeuclid", 2, 1.0, 400, 1e-7, "none", return_plan=True)
    res = shwd.entropic_ot(x.to(dev()), y.to(dev()), "sqeuclid", 2.0, 1.0, 400, 1.0, 1e-7)
    assert res.status() == 0
    assert iters_ref < 400, "test premise: the reference stops early"
    assert abs(res.iterations() - iters_ref) <= 2  # the statistic sits at the float32 noise floor
    assert rel(res.cost, c) < TOL


def test_unbatched_int_and_ragged_inputs(shwd):
    L = shwd.losses
    torch.manual_seed(2)
    x = torch.randn(37, 3)
    y = torch.randn(53, 3)
    crit = L.Sinkhorn_Distance_Loss(eps=0.1, max_iter=15, batch_reduction="none")
    loss, P, C = crit(x.to(dev()), y.to(dev()), dev())
    ref = oracle.log_sinkhorn(x, y, "sqeuclid", 2, 0.1, 15)
    assert loss.dim() == 0 and P.shape == (37, 53)
    assert loss.item() == pytest.approx(ref.item(), rel=TOL)
    with pytest.raises(ValueError):
        L.Sinkhorn_Distance_Loss(eps=0.1, max_iter=5, type_of_cost_norm="L3")
    with pytest.raises(RuntimeError):
        shwd.entropic_ot(x, y)  # CPU tensors: no fallback


def test_geodesic_distance_w_dropin(shwd):
    """Geodesic_distance_W(device, p)(x, y): mean_b cost_b^(1/p), gradients to both clouds, retain_graph semantics."""
    L = shwd.losses
    torch.manual_seed(4)
    x = F.normalize(torch.randn(3, 128, 3), dim=-1)
    y = F.normalize(torch.randn(3, 128, 3) + 0.2, dim=-1)
    crit = L.Geodesic_distance_W(device=dev(), p=2, eps=0.02, max_iter=30)
    xg = x.clone().to(dev()).requires_grad_(True)
    yg = y.clone().to(dev()).requires_grad_(True)
    loss = crit(xg, yg)
    loss.backward(retain_graph=True)
    g1 = xg.grad.clone()
    loss.backward()  # second backward over the retained graph (s2_wasserstein.py:253)
    assert torch.allclose(xg.grad, 2 * g1, rtol=1e-6, atol=0)
    xr = x.clone().requires_grad_(True)
    yr = y.clone().requires_grad_(True)
    ref = oracle.entropic_w(xr, yr, "geodesic", 2, 0.02, 30)
    ref.backward()
    assert loss.item() == pytest.approx(ref.item(), rel=TOL)
    assert rel(g1, xr.grad) < TOL
    # un-batched branch (:46-48) and the squared-Euclidean sibling
    l1 = L.Cos_disimilarity_W(device=dev(), p=2, eps=0.05, max_iter=20)(x[0].to(dev()), y[0].to(dev()))
    assert l1.item() == pytest.approx(oracle.entropic_w(x[0], y[0], "sqeuclid", 2, 0.05, 20).item(), rel=TOL)


def test_max_wrapper_trains_phi_and_returns_reference_tuple(shwd):
    L = shwd.losses
    torch.manual_seed(5)
    phi = L.Norm_Flow_structure(flow_name="Residual", n_flow_layer=2).to(dev())
    phi_op = torch.optim.Adam(phi.parameters(), lr=1e-2, betas=(0.5, 0.999))
    csw = L.Geodesic_distance_W(device=dev(), p=2, eps=0.05, max_iter=10)
    crit = L.max_cos_disimilarity_wassersten_distance(phi=phi, CSW=csw, device=dev(), phi_op=phi_op, max_iter=1, lam=1.0)
    a = torch.randn(2, 64, 3, device=dev())
    b = torch.randn(2, 64, 3, device=dev(), requires_grad=True)
    before = [p.detach().clone() for p in phi.parameters()]
    out, ta, tb = crit(a, b, train_or_test="train")
    out.backward()
    assert out.dim() == 0 and ta.shape == a.shape and tb.shape == b.shape
    assert b.grad is not None and torch.isfinite(b.grad).all()
    assert any(not torch.equal(p0, p1.detach()) for p0, p1 in zip(before, phi.parameters()))
    assert crit.phi is phi and crit.phi_op is phi_op  # checkpoint code reads these (train_W_COS.py:204-205)
    pm = L.pseudo_max_cos_disimilarity_wassersten_distance(csw, dev(), phi_num=2, n_flow_layer=1, flow_name="Planar")
    v, _, _ = pm(a, b.detach())
    assert torch.isfinite(v)


@pytest.mark.parametrize("mode", ["max", "mean", "softmax"])
@pytest.mark.parametrize("solver", ["exact", "sinkhorn"])
def test_pseudo_max_wrapper_one_launch_equals_one_call_per_flow(shwd, mode, solver, capsys):
    """pseudo_max_cos_disimilarity_wassersten_distance (s2_wasserstein.py:272-344): all phi_num x B pairs through ONE solver
    launch (the default here) against the reference's order -- one criterion call per flow, `if cswd > max_cswd` on the host."""
    L = shwd.losses
    torch.manual_seed(5)
    csw = L.Cos_disimilarity_W(dev(), p=2) if solver == "exact" else L.Geodesic_distance_W(dev(), p=2, eps=0.05, max_iter=30)
    pm = L.pseudo_max_cos_disimilarity_wassersten_distance(csw, dev(), phi_num=5, n_flow_layer=2, flow_name="Residual",
                                                           mean_or_max_or_softmax=mode)
    a = F.normalize(torch.randn(6, 128, 3), dim=-1).to(dev())
    b = (F.normalize(torch.randn(6, 128, 3), dim=-1) * 1.05).to(dev())
    v1, fa, fb = pm(a, b)
    pm.batched = False
    per_flow = [csw(phi(a), phi(b)) for phi in pm.phi_list]  # the reference's loop, spelled out
    want = {"max": max(per_flow), "mean": sum(per_flow) / 5,
            "softmax": (F.softmax(torch.stack(per_flow), 0) * torch.stack(per_flow)).sum()}[mode]
    v0, ga, gb = pm(a, b)
    e = (abs(v1.item() - want.item()) / want.item(), abs(v0.item() - want.item()) / want.item())
    with capsys.disabled():
        print("pseudo-max %s / %s: one launch %.1e, per-flow calls %.1e from the spelled-out loop" % ((mode, solver) + e))
    assert e[0] < 1e-6 and e[1] < 1e-6 and torch.equal(fa, ga) and torch.equal(fb, gb)
    assert torch.equal(fa, pm.phi_list[-1](a))  # the LAST flow's clouds are what the reference returns


# ------------------------------------------------------------------------------------------------------- Chamfer ----
@pytest.mark.parametrize("B,N,M", [(2, 40, 33), (3, 1024, 1024), (1, 2500, 300)])
@pytest.mark.parametrize("br,pr", [("mean", "mean"), ("sum", "mean"), (None, "sum")])
def test_chamfer_matches_oracle(shwd, B, N, M, br, pr):
    torch.manual_seed(N + M)
    x = torch.randn(B, N, 3)
    y = torch.randn(B, M, 3) * 1.1
    xr = x.clone().requires_grad_(True)
    yr = y.clone().requires_grad_(True)
    ref, _ = oracle.chamfer_distance(xr, yr, br, pr)
    ref.sum().backward()
    xg = x.clone().to(dev()).requires_grad_(True)
    yg = y.clone().to(dev()).requires_grad_(True)
    out, none = shwd.losses.chamfer_distance(xg, yg, batch_reduction=br, point_reduction=pr)
    out.sum().backward()
    assert none is None
    assert rel(out, ref) < 1e-6
    assert rel(xg.grad, xr.grad) < TOL and rel(yg.grad, yr.grad) < TOL
    # indices and distances are bit-exact against torch's dense evaluation
    d = ((x.unsqueeze(2) - y.unsqueeze(1)) ** 2).sum(-1)
    d_xy, d_yx, i_xy, i_yx = shwd.chamfer_nn(x.to(dev()), y.to(dev()))
    assert torch.equal(d_xy.cpu(), d.min(2).values) and torch.equal(d_yx.cpu(), d.min(1).values)
    assert torch.equal(i_xy.cpu().long(), d.argmin(2)) and torch.equal(i_yx.cpu().long(), d.argmin(1))
    # single_directional (pytorch3d: only the x -> y term), un-batched clouds, and a non-unit upstream gradient
    xs = x[0].clone().to(dev()).requires_grad_(True)
    ys = y[0].clone().to(dev()).requires_grad_(True)
    one, _ = shwd.losses.chamfer_distance(xs, ys, batch_reduction=br, point_reduction=pr, single_directional=True)
    (one.sum() * 3.0).backward()
    xq, yq = x[0].clone().requires_grad_(True), y[0].clone().requires_grad_(True)
    dq = ((xq.unsqueeze(1) - yq.unsqueeze(0)) ** 2).sum(-1).min(1).values
    refq = dq.mean() if pr == "mean" else dq.sum()
    (refq * 3.0).backward()
    assert rel(one, refq) < 1e-6 and rel(xs.grad, xq.grad) < TOL and rel(ys.grad, yq.grad) < TOL


# ---------------------------------------------------------------------------------------------------------- sort ----
@pytest.mark.parametrize("segs,length", [(7, 1), (5, 31), (3, 1000), (4, 4096), (2, 8192), (2, 10000), (1, 70000)])
def test_segmented_sort_is_bit_exact_stable(shwd, segs, length):
    torch.manual_seed(length)
    k = torch.randn(segs, length)
    # heavy ties + specials: -inf < zeros (both signs, input order kept) < 1 < inf < nan
    k[:, ::3] = torch.randint(-3, 4, (segs, (length + 2) // 3)).float()
    if length >= 31:
        k[0, 5] = float("nan")
        k[0, 6] = float("inf")
        k[0, 7] = float("-inf")
        k[0, 8] = -0.0
        k[0, 9] = 0.0
        k[0, 20] = float("nan")
        k[0, 21] = -0.0
    ref_v, ref_i = torch.sort(k, dim=-1, stable=True)
    v, i = shwd.segmented_sort_raw(k.to(dev()))
    assert torch.equal(i.cpu(), ref_i)
    assert torch.equal(v.cpu().view(torch.int32), ref_v.view(torch.int32))  # same bits, incl. signed zeros / NaN payloads


@pytest.mark.parametrize("method", [0, 1])  # 0: bucket pass where the row allows it, 1: radix passes only
@pytest.mark.parametrize("kind", ["circle", "signed", "ties", "constant", "tiny_range", "specials", "denormal_span", "clustered",
                                  "few_ties_spread", "one_outlier"])
@pytest.mark.parametrize("segs,length", [(5, 1), (7, 33), (3, 1000), (4, 4096), (2, 4224), (3, 5000), (3, 8192), (2, 10000), (2, 16384)])
def test_trimmed_digit_sort_is_bit_exact_stable(shwd, kind, segs, length, method):
    """The sliced losses' own sort (digits trimmed to the bits in which a row's keys differ, csrc/sliced.cu): the int32
    permutation equals torch.sort(stable=True) bit for bit on rows that differ in 27 bits (circle coordinates: 3 passes),
    32 bits (signed), a handful of bits, none at all, with NaN / +-inf / +-0, and across the denormal boundary."""
    g = torch.Generator().manual_seed(segs * 100 + length)
    if kind == "circle":
        k = torch.rand(segs, length, generator=g)
    elif kind == "signed":
        k = torch.randn(segs, length, generator=g) * 3
    elif kind == "ties":
        k = torch.randint(0, 7, (segs, length), generator=g).float() / 7
    elif kind == "constant":
        k = torch.full((segs, length), 0.375)
    elif kind == "tiny_range":
        k = 0.5 + torch.randint(0, 40, (segs, length), generator=g).float() * 2.0 ** -24
    elif kind == "specials":
        k = torch.randn(segs, length, generator=g)
        k[:, ::5] = float("nan")
        k[:, 1::7] = float("inf")
        k[:, 2::11] = -float("inf")
        k[:, 3::13] = 0.0
        k[:, 4::17] = -0.0
    elif kind == "clustered":  # half of the keys inside 1e-4 of the range: some buckets overflow, some rows may not
        k = torch.rand(segs, length, generator=g)
        k[:, ::2] = 0.3 + 1e-4 * torch.rand(segs, (length + 1) // 2, generator=g)
    elif kind == "few_ties_spread":  # spread keys with duplicated values: equal keys share a bucket, the index decides
        k = torch.rand(segs, length, generator=g)
        k[:, 1::3] = k[:, ::3][:, :k[:, 1::3].shape[1]]
    elif kind == "one_outlier":  # the range is set by one key; everything else falls into a few buckets
        k = torch.rand(segs, length, generator=g) * 1e-3
        k[:, 0] = 1e6
    else:
        k = torch.rand(segs, length, generator=g) * 1e-37 * torch.randint(0, 3, (segs, length), generator=g).float() * 1e-3
    kd = k.to(dev())
    assert shwd._lib.lib().shwd_sort_set_method(method) == 0
    try:
        so, pe = shwd.ops._sort_i32(kd)
    finally:
        shwd._lib.lib().shwd_sort_set_method(0)
    ref_v, ref_p = torch.sort(kd, dim=-1, stable=True)
    assert torch.equal(pe.long(), ref_p)
    assert torch.equal(so.isnan(), ref_v.isnan()) and torch.equal(so[~so.isnan()], ref_v[~ref_v.isnan()])  # (-0.0 == +0.0)


@pytest.mark.parametrize("mode", ["circle", "line"])
@pytest.mark.parametrize("B,N,P", [(1, 1, 3), (2, 700, 9), (3, 4096, 5), (1, 4224, 2)])
def test_projected_sort_equals_projection_then_sort(shwd, mode, B, N, P):
    """shwd_sort_projected (sort CTAs that compute their own keys) against the stand-alone projection kernel followed by
    the sort: same keys, same permutation, same sorted values."""
    g = torch.Generator().manual_seed(B + N + P)
    x = (torch.randn(B, N, 3, generator=g) * torch.tensor([1.0, 0.6, 1.7])).to(dev())
    lib = shwd._lib.lib()
    if mode == "circle":
        fr, _ = torch.linalg.qr(torch.randn(P, 3, 2, generator=g))
        fr = fr.contiguous().to(dev())
        keys = shwd.ops.ProjectCircleFn.apply(x, fr)
    else:
        fr = F.normalize(torch.randn(P, 3, generator=g), dim=-1).to(dev())
        keys = shwd.ops.ProjectLineFn.apply(x, fr)
    so = torch.empty(B * P, N, device=dev())
    pe = torch.empty(B * P, N, device=dev(), dtype=torch.int32)
    assert N <= lib.shwd_sort_projected_max_points()
    shwd._lib.check(lib.shwd_sort_projected(x.data_ptr(), fr.data_ptr(), B, N, P, 1 if mode == "circle" else 2, so.data_ptr(),
                                            pe.data_ptr(), torch.cuda.current_stream().cuda_stream), "shwd_sort_projected")
    ref_v, ref_p = torch.sort(keys.reshape(B * P, N), dim=-1, stable=True)
    assert torch.equal(pe.long(), ref_p) and torch.equal(so, ref_v)


def test_sort_gradient_scatters_through_permutation(shwd):
    torch.manual_seed(0)
    k = torch.randn(6, 300)
    kr = k.clone().requires_grad_(True)
    w = torch.randn(6, 300)
    (torch.sort(kr, dim=-1)[0] * w).sum().backward()
    kg = k.clone().to(dev()).requires_grad_(True)
    s, _ = shwd.ops.SegmentedSortFn.apply(kg)
    (s * w.to(dev())).sum().backward()
    assert torch.equal(kg.grad.cpu(), kr.grad)


# -------------------------------------------------------------------------------------------------------- sliced ----
def test_emd1d_circle_matches_reference_fixture(shwd):
    d = gold("emd1d_circle")
    u = torch.from_numpy(d["u"]).to(dev()).requires_grad_(True)
    v = torch.from_numpy(d["v"]).to(dev()).requires_grad_(True)
    w = shwd.losses.emd1D_circle(u, v)
    w.sum().backward()
    assert rel(w, torch.from_numpy(d["w"])) < TOL
    assert rel(u.grad, torch.from_numpy(d["gu"])) < TOL and rel(v.grad, torch.from_numpy(d["gv"])) < TOL


def test_weighted_emd1d_circle_matches_reference_fixture(shwd):
    """emd1D_circle with u_weights / v_weights (max_spherical_sliced_w.py:217-228): value and the gradients w.r.t. the
    coordinates AND the weights against the unmodified reference; and sliced_cost(p=1, weights) through the same path."""
    d = gold("emd1d_circle_weighted")
    t = {k: torch.from_numpy(d[k]).to(dev()).requires_grad_(True) for k in ("u", "v", "uw", "vw")}
    w = shwd.losses.emd1D_circle(t["u"], t["v"], u_weights=t["uw"], v_weights=t["vw"], p=1)
    w.sum().backward()
    errs = [rel(w, torch.from_numpy(d["w"]))] + [rel(t[k].grad, torch.from_numpy(d["g" + k])) for k in ("u", "v", "uw", "vw")]
    print("weighted emd1D_circle: w %.2e gu %.2e gv %.2e guw %.2e gvw %.2e" % tuple(errs))
    assert errs[0] < TOL and max(errs[1:]) < 2e-5
    # uniform weights given explicitly == the fused uniform kernel
    g = torch.Generator().manual_seed(3)
    Xs = F.normalize(torch.randn(120, 3, generator=g), dim=-1).to(dev())
    Xt = F.normalize(torch.randn(90, 3, generator=g) + 0.3, dim=-1).to(dev())
    U, _ = torch.linalg.qr(torch.randn(7, 3, 2, generator=g))
    U = U.to(dev())
    a = shwd.losses.sliced_cost(Xs, Xt, U, p=1)
    b = shwd.losses.sliced_cost(Xs, Xt, U, p=1, u_weights=torch.full((120,), 1 / 120, device=dev()),
                                v_weights=torch.full((90,), 1 / 90, device=dev()))
    assert rel(b, a) < TOL
    # p != 1 with explicit uniform weights: the table-driven bisection kernel against the closed-form one
    a2 = shwd.losses.sliced_cost(Xs, Xt, U, p=2)
    b2 = shwd.losses.sliced_cost(Xs, Xt, U, p=2, u_weights=torch.full((120,), 1 / 120, device=dev()),
                                 v_weights=torch.full((90,), 1 / 90, device=dev()))
    assert rel(b2, a2) < TOL


@pytest.mark.parametrize("S,n,m", [(5, 1, 1), (5, 7, 3), (4, 1024, 1024), (3, 1500, 1500), (3, 3000, 2500), (3, 4096, 4096),
                                   (2, 5000, 4800), (2, 5120, 5120), (2, 6000, 5000), (2, 8000, 8100), (2, 16000, 16500),
                                   (1, 16384, 16384), (1, 20000, 17000)])
def test_emd1d_circle_matches_oracle_every_bucket(shwd, S, n, m):
    """circular_w1_kernel<C> keeps C merged entries per thread in registers (C = 4, 8, 12, 16, 20 by n + m; C = 32, 64 with
    recomputed CDF keys up to n + m = 32768): values and the gradients w.r.t. the unsorted circle coordinates against
    emd1D_circle (oracle/sliced.py) in every bucket; beyond that the four-sort composition circular_w1_large (sort kernel
    with global scratch) takes over."""
    uv = _tie_free(S, n + m, 100 + n)  # one shuffled tie-free row split in two: no u == v tie either (a tie's order in
    u0, v0 = uv[:, :n].contiguous(), uv[:, n:].contiguous()  # the merged sort decides two gradient entries)
    g = torch.Generator().manual_seed(S + n)
    wgt = torch.rand(S, generator=g) + 0.5
    ur, vr = u0.clone().requires_grad_(True), v0.clone().requires_grad_(True)
    wr = oracle.emd1d_circle(ur, vr)
    (wr * wgt).sum().backward()
    u = u0.to(dev()).requires_grad_(True)
    v = v0.to(dev()).requires_grad_(True)
    w = shwd.losses.emd1D_circle(u, v)
    (w * wgt.to(dev())).sum().backward()
    assert rel(w, wr.detach()) < TOL
    # n == m a power of two: the CDF difference F is a multiple of 1/n, every float32 prefix sum is exact (floor ~5e-7).
    # Otherwise: wherever F_{k-1} and F_k straddle the level median the gradient entry is 2 med - F_{k-1} - F_k, which carries the
    # ACCUMULATED rounding of two float32 prefix sums; torch's own float32 result sits `floor` from its float64 one
    # (6e-6 .. 3e-5 here, and 1e-4 from a strictly sequential float32 cumsum), so the bound is max(1e-5, 8 x floor).
    u64, v64 = u0.double().requires_grad_(True), v0.double().requires_grad_(True)
    (oracle.emd1d_circle(u64, v64) * wgt.double()).sum().backward()
    floor = max(rel(ur.grad, u64.grad), rel(vr.grad, v64.grad))
    bound = max(TOL, 8 * floor)
    assert rel(u.grad, ur.grad) < bound and rel(v.grad, vr.grad) < bound
    assert rel(u.grad, u64.grad) < bound and rel(v.grad, v64.grad) < bound


def test_spherical_sliced_w1_matches_reference_fixture(shwd):
    d = gold("ssw_p1")
    xs = torch.from_numpy(d["Xs"]).to(dev()).requires_grad_(True)
    xt = torch.from_numpy(d["Xt"]).to(dev()).requires_grad_(True)
    U = torch.from_numpy(d["U"]).to(dev())
    loss = shwd.losses.sliced_cost(xs, xt, U, p=1)
    loss.backward()
    assert loss.item() == pytest.approx(float(d["loss"]), rel=TOL)
    assert rel(xs.grad, torch.from_numpy(d["gx"])) < 5e-5  # atan2 / normalise chain: reference f32-vs-f64 is ~2e-5 here
    assert rel(xt.grad, torch.from_numpy(d["gy"])) < 5e-5


@pytest.mark.parametrize("n,m,P", [(64, 64, 8), (500, 333, 16), (4096, 4096, 4), (16384, 16384, 3), (9000, 7000, 2)])
def test_spherical_sliced_w1_matches_oracle(shwd, n, m, P):
    g = torch.Generator().manual_seed(n + m)
    Xs = F.normalize(torch.randn(n, 3, generator=g), dim=-1)
    Xt = F.normalize(torch.randn(m, 3, generator=g) + torch.tensor([0.4, 0.0, 0.1]), dim=-1)
    U, _ = torch.linalg.qr(torch.randn(P, 3, 2, generator=g))
    ref = oracle.sliced_wasserstein_sphere_p1(Xs, Xt, U)
    out = shwd.losses.sliced_cost(Xs.to(dev()), Xt.to(dev()), U.to(dev()), p=1)
    assert out.item() == pytest.approx(ref.item(), rel=2e-5)


SLICED_FULL = [
    # n, m, P        merged-entry bucket of circular_w1 (C per thread) / circular_wp CTA size
    (4096, 4096, 8),     # cfg3: C = 16, gradient rows staged and permuted in shared memory; 256-thread circular_wp
    (5000, 5000, 4),     # C = 20
    (7000, 8000, 3),     # C = 32 (register-light variant: F recomputed), 1024-thread circular_wp (> 14336 merged entries)
    (16384, 16384, 2),   # cfg4: C = 64, direct scatter; compact sort layout
    (300, 1500, 5),      # ragged, C = 4
]


@pytest.mark.parametrize("n,m,P", SLICED_FULL)
@pytest.mark.parametrize("mode", ["circle_w1", "circle_wp", "line"])
def test_fused_sliced_loss_gradients_match_oracle_at_full_sizes(shwd, mode, n, m, P):
    """The path users call -- sliced_cost / sliced_wasserstein_distance -> ops.SlicedLossFn -> shwd_*_scatter (gradients
    written through the int32 permutations) -- value AND both cloud gradients against the reference's autograd
    (oracle.sliced_wasserstein_sphere / euclid_sliced_wasserstein, max_spherical_sliced_w.py:251-286,
    Flow_ellipsoid.ipynb cell 5) at BASELINE cfg3 / cfg4 sizes and in every register bucket.  The reference's own
    float32-vs-float64 distance (`floor`) is printed beside the achieved error; bound max(1e-5, 8 x floor)."""
    if mode == "line" and n != m:
        pytest.skip("the notebook's Euclidean sliced W needs equally sized clouds")
    g = torch.Generator().manual_seed(n + 3 * m + P)
    Xs = F.normalize(torch.randn(n, 3, generator=g), dim=-1) * (1 + 0.1 * torch.rand(n, 1, generator=g))
    Xt = F.normalize(torch.randn(m, 3, generator=g) + torch.tensor([0.4, 0.0, 0.1]), dim=-1)
    if mode == "line":
        fr = F.normalize(torch.randn(P, 3, generator=g), dim=-1)
        ofn = lambda a, b: oracle.euclid_sliced_wasserstein(a, b, fr.to(a.dtype), 2)
        cfn = lambda a, b: shwd.losses.sliced_wasserstein_distance(a, b, p=2, device=dev(), projections=fr.to(dev()))
    else:
        fr, _ = torch.linalg.qr(torch.randn(P, 3, 2, generator=g))
        p = 1 if mode == "circle_w1" else 2
        ofn = lambda a, b: (oracle.sliced_wasserstein_sphere_p1(a, b, fr.to(a.dtype)) if p == 1
                            else oracle.sliced_wasserstein_sphere(a, b, fr.to(a.dtype), p=2))
        cfn = lambda a, b: shwd.losses.sliced_cost(a, b, fr.to(dev()), p=p)
    outs = {}
    for dt in (torch.float32, torch.float64):
        a, b = Xs.detach().clone().to(dt).requires_grad_(True), Xt.detach().clone().to(dt).requires_grad_(True)
        v = ofn(a, b)
        v.backward()
        outs[dt] = (v.detach(), a.grad, b.grad)
    a, b = Xs.detach().clone().to(dev()).requires_grad_(True), Xt.detach().clone().to(dev()).requires_grad_(True)
    v = cfn(a, b)
    v.backward()
    floor = [rel(x32, x64) for x32, x64 in zip(outs[torch.float32], outs[torch.float64])]
    err = [rel(x, x32) for x, x32 in zip((v, a.grad, b.grad), outs[torch.float32])]
    err64 = [rel(x, x64) for x, x64 in zip((v, a.grad, b.grad), outs[torch.float64])]
    print("%s n=%d m=%d: value %.2e / gx %.2e / gy %.2e vs float32 reference (its own floor %.2e / %.2e / %.2e); vs float64 %.2e / %.2e / %.2e"
          % ((mode, n, m) + tuple(err) + tuple(floor) + tuple(err64)))
    for e, e64, f in zip(err, err64, floor):
        assert min(e, e64) < max(TOL, 8 * f)


def _tie_free(S, n, seed, lo=0.0, width=1.0):
    """Distinct float32 circle coordinates per row (torch.sort without stable=True orders ties arbitrarily): a
    shuffled jittered grid, so distinctness does not rely on luck at n = 4096."""
    g = torch.Generator().manual_seed(seed)
    grid = (torch.arange(n, dtype=torch.float64) + 0.1 + 0.8 * torch.rand(S, n, generator=g, dtype=torch.float64)) / n
    perm = torch.argsort(torch.rand(S, n, generator=g), dim=1)
    x = ((torch.gather(grid, 1, perm) * width + lo) % 1.0).float()
    assert all(torch.unique(r).numel() == n for r in x)
    return x


def test_binary_search_circle_matches_reference_fixture(shwd):
    d = gold("binary_search_circle_p2")
    w = shwd.losses.binary_search_circle(torch.from_numpy(d["u"]).to(dev()), torch.from_numpy(d["v"]).to(dev()), p=2)
    assert rel(w, torch.from_numpy(d["w"])) < TOL


def test_spherical_sliced_w2_matches_reference_fixture(shwd):
    d = gold("ssw_p2")
    xs = torch.from_numpy(d["Xs"]).to(dev()).requires_grad_(True)
    xt = torch.from_numpy(d["Xt"]).to(dev()).requires_grad_(True)
    loss = shwd.losses.sliced_cost(xs, xt, torch.from_numpy(d["U"]).to(dev()), p=2)
    loss.backward()
    assert loss.item() == pytest.approx(float(d["loss"]), rel=TOL)
    assert rel(xs.grad, torch.from_numpy(d["gx"])) < TOL and rel(xt.grad, torch.from_numpy(d["gy"])) < TOL


@pytest.mark.parametrize("S,n,m,p", [(9, 70, 55, 2), (5, 64, 64, 2), (6, 500, 333, 2), (4, 1024, 1024, 2), (3, 300, 300, 3),
                                     (3, 200, 257, 1.5), (2, 4096, 4096, 2), (4, 1, 5, 2), (4, 7, 1, 2), (2, 3000, 9000, 2),
                                     (2, 9000, 8000, 2), (1, 16384, 16384, 2), (1, 15000, 16000, 3)])  # 1024-thread CTAs beyond 14336
def test_circular_wp_matches_oracle(shwd, S, n, m, p):
    """binary_search_circle (max_spherical_sliced_w.py:117-207): W_p^p, the rotation found, and the gradients w.r.t.
    the unsorted coordinates.  The bisection's last sign decisions sit at the float32 noise level of a sum of m
    cancelling terms, so the rotation may differ by the final bracket width (~1e-7); the gradient bound is
    max(1e-5, 8 x the reference's own float32-vs-float64 distance), like the other non-north-star rows."""
    u, v = _tie_free(S, n, 100 + n), _tie_free(S, m, 200 + m, lo=0.2, width=0.7)
    ur, vr = u.clone().requires_grad_(True), v.clone().requires_grad_(True)
    wr, thr = oracle.sliced.binary_search_circle(ur, vr, p=p, return_theta=True)
    wr.sum().backward()
    u64, v64 = u.double().requires_grad_(True), v.double().requires_grad_(True)
    oracle.sliced.binary_search_circle(u64, v64, p=p).sum().backward()
    floor = max(rel(ur.grad, u64.grad), rel(vr.grad, v64.grad))
    ug, vg = u.clone().to(dev()).requires_grad_(True), v.clone().to(dev()).requires_grad_(True)
    us, _ = shwd.ops.SegmentedSortFn.apply(ug)
    vs, _ = shwd.ops.SegmentedSortFn.apply(vg)
    w, th = shwd.ops.CircularWpFn.apply(us, vs, float(p), -1.0, 1.0, 1e-7)
    w.sum().backward()
    assert rel(w, wr) < TOL
    assert (th.cpu() - thr).abs().max().item() < 2e-6
    bound = max(TOL, 8 * floor)
    assert rel(ug.grad, ur.grad) < bound and rel(vg.grad, vr.grad) < bound, (rel(ug.grad, ur.grad), rel(vg.grad, vr.grad), floor)


def test_weighted_binary_search_circle_matches_reference_fixture(shwd):
    """binary_search_circle / sliced_cost with u_weights / v_weights (max_spherical_sliced_w.py:117,156-170, 251-286) against
    the unmodified reference: value, gradients w.r.t. the coordinates and w.r.t. the weights (p = 2 and p = 3)."""
    d = gold("binary_search_circle_weighted")
    for p in (2, 3):
        t = {k: torch.from_numpy(d[f"{k}_p{p}"]).to(dev()).requires_grad_(True) for k in ("u", "v", "uw", "vw")}
        w = shwd.losses.binary_search_circle(t["u"], t["v"], u_weights=t["uw"], v_weights=t["vw"], p=p)
        w.sum().backward()
        errs = [rel(w, torch.from_numpy(d[f"w_p{p}"]))] + [rel(t[k].grad, torch.from_numpy(d[f"g{k}_p{p}"])) for k in ("u", "v", "uw", "vw")]
        # The weights only see the final Cost through the ORDER of the merged CDF axis (:93-95), and the bisection stops where a
        # rolled v entry meets a u entry: which of the two sorts first there is decided by the last bit of the detached rotation,
        # so d W / d weights is discontinuous exactly at the optimum.  The reference's own float64 evaluation is 20-40 % away
        # from its float32 one on these inputs; that distance (the oracle is pinned to the reference on this fixture) is the bound.
        r64 = {k: torch.from_numpy(d[f"{k}_p{p}"]).double().requires_grad_(True) for k in ("u", "v", "uw", "vw")}
        oracle.sliced.binary_search_circle(r64["u"], r64["v"], p=p, u_weights=r64["uw"], v_weights=r64["vw"]).sum().backward()
        floor_w = max(rel(r64[k].grad, torch.from_numpy(d[f"g{k}_p{p}"])) for k in ("uw", "vw"))
        print("weighted binary_search_circle p=%d: w %.2e gu %.2e gv %.2e guw %.2e gvw %.2e (weights: reference float32 vs float64 %.2e)"
              % ((p,) + tuple(errs) + (floor_w,)))
        assert errs[0] < TOL and max(errs[1:3]) < 2e-5 and max(errs[3:]) < max(2e-5, floor_w)
    Xs = torch.from_numpy(d["Xs"]).to(dev()).requires_grad_(True)
    Xt = torch.from_numpy(d["Xt"]).to(dev()).requires_grad_(True)
    loss = shwd.losses.sliced_cost(Xs, Xt, torch.from_numpy(d["U"]).to(dev()), p=2, u_weights=torch.from_numpy(d["sc_uw"]).to(dev()),
                                   v_weights=torch.from_numpy(d["sc_vw"]).to(dev()))
    loss.backward()
    errs = (abs(loss.item() - float(d["sc_loss"])) / float(d["sc_loss"]), rel(Xs.grad, torch.from_numpy(d["sc_gx"])),
            rel(Xt.grad, torch.from_numpy(d["sc_gy"])))
    print("weighted sliced_cost p=2: loss %.2e gx %.2e gy %.2e" % errs)
    assert errs[0] < TOL and max(errs[1:]) < 2e-5


@pytest.mark.parametrize("S,n,m,p,zeros", [(4, 1000, 700, 2, False), (3, 257, 300, 3, False), (3, 64, 64, 2, False),
                                           (2, 4096, 4096, 2, False), (3, 120, 90, 2, True), (2, 1, 6, 2, False),
                                           (1, 9000, 8000, 2, False)])  # n + m > 7168: the 1024-thread CTAs
def test_weighted_circular_wp_matches_oracle(shwd, S, n, m, p, zeros):
    """The table-driven bisection (circular_wp_kernel<..., W = true>) against the oracle with random weights -- value, rotation,
    gradients w.r.t. coordinates and weights; ``zeros``: a quarter of the weights are exactly 0 (equal CDF entries)."""
    g = torch.Generator().manual_seed(31 * n + m)
    u, v = _tie_free(S, n, 300 + n), _tie_free(S, m, 400 + m, lo=0.2, width=0.7)
    uw, vw = torch.rand(n, generator=g) + 0.05, torch.rand(m, generator=g) + 0.05
    if zeros:
        uw[::4] = 0
        vw[1::4] = 0
    uw, vw = uw / uw.sum(), vw / vw.sum()
    r = [t.clone().requires_grad_(True) for t in (u, v, uw, vw)]
    wr, thr = oracle.sliced.binary_search_circle(r[0], r[1], p=p, return_theta=True, u_weights=r[2], v_weights=r[3])
    wr.sum().backward()
    r64 = [t.double().requires_grad_(True) for t in (u, v, uw, vw)]
    oracle.sliced.binary_search_circle(r64[0], r64[1], p=p, u_weights=r64[2], v_weights=r64[3]).sum().backward()
    floor = max(rel(a.grad, b.grad) for a, b in zip(r, r64))
    t = [x.clone().to(dev()).requires_grad_(True) for x in (u, v, uw, vw)]
    w = shwd.losses.binary_search_circle(t[0], t[1], u_weights=t[2], v_weights=t[3], p=p)
    w.sum().backward()
    errs = [rel(w, wr)] + [rel(a.grad, b.grad) for a, b in zip(t, r)]
    print("weighted circle_wp n=%d m=%d p=%g: w %.2e gu %.2e gv %.2e guw %.2e gvw %.2e (reference's own float32-vs-float64 floor %.2e)"
          % ((n, m, p) + tuple(errs) + (floor,)))
    assert errs[0] < TOL
    bound = max(2e-5, 8 * floor)
    assert max(errs[1:]) < bound, (errs, floor)


@pytest.mark.parametrize("S,n,p", [(7, 1, 2), (7, 2, 2), (6, 64, 2), (5, 1024, 2), (4, 1024, 3), (3, 4096, 2), (2, 16384, 2),
                                   (9, 3, 2), (8, 7, 2), (6, 100, 2), (6, 333, 3), (5, 1000, 2), (4, 1000, 1.5), (3, 4000, 2),
                                   (3, 5000, 2), (2, 12000, 2), (2, 15000, 3)])
def test_circular_wp_dyadic_shortcut_is_bit_identical(shwd, S, n, p):
    """Equal sizes: the closed-form searches (on the 1/n grid for powers of two: dcost_dyadic, cost_pass_*_dyadic; safely off the
    grid for any size: between_safe, cost_pass_*_between) against the generic searches of the same kernel -- value, rotation and
    both gradients must be the same bits; also through the fused sliced node."""
    u, v = _tie_free(S, n, 500 + n), _tie_free(S, n, 600 + n, lo=0.3, width=0.6)
    lib = shwd._lib.lib()
    out = []
    try:
        for on in (1, 0):
            assert lib.shwd_circular_wp_set_dyadic(on) == 0
            ug, vg = u.clone().to(dev()).requires_grad_(True), v.clone().to(dev()).requires_grad_(True)
            us, _ = shwd.ops.SegmentedSortFn.apply(ug)
            vs, _ = shwd.ops.SegmentedSortFn.apply(vg)
            w, th = shwd.ops.CircularWpFn.apply(us, vs, float(p), -1.0, 1.0, 1e-7)
            w.sum().backward()
            res = [w.detach().clone(), th.clone(), ug.grad.clone(), vg.grad.clone()]
            if n >= 64:
                g = torch.Generator().manual_seed(n)
                Xs = F.normalize(torch.randn(2, n, 3, generator=g), dim=-1).to(dev()).requires_grad_(True)
                Xt = F.normalize(torch.randn(2, n, 3, generator=g) + 0.3, dim=-1).to(dev()).requires_grad_(True)
                U, _ = torch.linalg.qr(torch.randn(5, 3, 2, generator=g))
                loss = shwd.ops.spherical_sliced_wp(Xs, Xt, U.to(dev()), float(p))
                loss.sum().backward()
                res += [loss.detach().clone(), Xs.grad.clone(), Xt.grad.clone()]
            out.append(res)
    finally:
        lib.shwd_circular_wp_set_dyadic(1)
    for a, b in zip(*out):
        assert torch.equal(a, b)
    # and the rotation found is on the grid for most rows of a power-of-two size (the shortcut was actually exercised)
    th = out[0][1]
    if n & (n - 1) == 0:
        assert ((th * n) == torch.floor(th * n)).float().mean().item() > 0.5 or n <= 2


@pytest.mark.parametrize("B,N,P", [(8, 4096, 64), (2, 16384, 40), (40, 512, 300), (5, 4100, 33)])
def test_project_circle_bwd_wide_rows_are_bit_identical(shwd, B, N, P):
    """project_circle_bwd4_kernel (four points per lane, N % 4 == 0) against the one-point-per-lane kernel: same bits."""
    g = torch.Generator().manual_seed(B * N + P)
    x = F.normalize(torch.randn(B, N, 3, generator=g), dim=-1).to(dev())
    U, _ = torch.linalg.qr(torch.randn(P, 3, 2, generator=g))
    gk = torch.randn(B, P, N, generator=g).to(dev())
    lib = shwd._lib.lib()
    out = []
    try:
        for on in (1, 0):
            assert lib.shwd_project_bwd_set_wide(on) == 0
            xg = x.clone().requires_grad_(True)
            shwd.ops.ProjectCircleFn.apply(xg, U.to(dev())).backward(gk)
            out.append(xg.grad.clone())
    finally:
        lib.shwd_project_bwd_set_wide(1)
    assert torch.equal(out[0], out[1])


def test_circular_wp_full_size_properties(shwd):
    """BASELINE config 3 (N=4096, 512 slices), p=2: properties that need no CPU reference.  (i) a common rotation of
    both clouds on the circle leaves W unchanged; (ii) W(u,u) = 0; (iii) W(u,v) = W(v,u); (iv) the input order of the
    points is irrelevant; (v) the optimal rotation can only improve on theta = 0 (the cut-at-0 line transport)."""
    S, n = 512, 4096
    u, v = _tie_free(S, n, 7).to(dev()), _tie_free(S, n, 8, lo=0.3, width=0.5).to(dev())
    bsc = shwd.losses.binary_search_circle
    w = bsc(u, v, p=2)
    assert torch.isfinite(w).all() and (w >= 0).all()
    shift = torch.rand(S, 1, device=dev())
    w_rot = bsc((u + shift) % 1.0, (v + shift) % 1.0, p=2)
    assert rel(w_rot, w) < 1e-4  # the rotated coordinates are re-rounded to float32 (1e-7 on values of ~1e-1 gaps)
    assert bsc(u, u.clone(), p=2).abs().max().item() < 1e-10
    assert rel(bsc(v, u, p=2), w) < TOL
    perm = torch.randperm(n, device=dev())
    assert torch.equal(bsc(u[:, perm], v, p=2), w)
    us, vs = torch.sort(u, -1)[0], torch.sort(v, -1)[0]
    line = ((us - vs) ** 2).mean(-1)
    assert (w <= line * (1 + 1e-5) + 1e-9).all()


def test_project_circle_keys(shwd):
    g = torch.Generator().manual_seed(9)
    X = torch.randn(2, 300, 3, generator=g)
    U, _ = torch.linalg.qr(torch.randn(12, 3, 2, generator=g))
    keys = shwd.ops.ProjectCircleFn.apply(X.to(dev()), U.to(dev()))
    for b in range(2):
        ref = oracle.project_circle(X[b], U)
        d = (keys[b].cpu() - ref).abs()
        d = torch.minimum(d, 1 - d)  # the circle wraps at 0/1
        assert d.max().item() < 1e-6, d.max().item()


def test_project_circle_keys_special_points(shwd):
    """The circle coordinate at the points where atan2 is decided by sign bits or is undefined: projections on the axes of
    the frame (exact zeros of either sign), the zero vector (normalize -> (0, 0) -> atan2(-0, -0) = -pi -> 0), NaN and inf
    coordinates (NaN).  Against the reference's own chain (max_spherical_sliced_w.py:270-279) in torch on the device, and the
    achieved distance on ordinary points."""
    U = torch.zeros(1, 3, 2)
    U[0, 0, 0] = 1.0
    U[0, 1, 1] = 1.0  # frame = (e_x, e_y): a = x, c = y
    pts = [[1.0, 0.0, 0.3], [-1.0, 0.0, 0.3], [0.0, 1.0, 0.3], [0.0, -1.0, 0.3], [1.0, -0.0, 0.0], [-2.0, -0.0, 0.0],
           [-0.0, 3.0, 0.0], [0.0, 0.0, 1.0], [-0.0, 0.0, 1.0], [0.0, -0.0, 1.0], [-0.0, -0.0, 1.0], [1e-30, 1e-30, 0.0],
           [1e-20, -3e-20, 0.0], [float("nan"), 1.0, 0.0], [1.0, float("inf"), 0.0], [0.5, 0.5, 0.0], [-0.5, 0.5, 0.0]]
    g = torch.Generator().manual_seed(3)
    X = torch.cat([torch.tensor(pts), torch.randn(4000, 3, generator=g)]).unsqueeze(0).to(dev())
    keys = shwd.ops.ProjectCircleFn.apply(X, U.to(dev()))[0, 0]
    q = F.normalize(torch.matmul(U.to(dev()).transpose(-1, -2)[0], X[0].T).T, p=2, dim=-1)
    ref = (torch.atan2(-q[:, 1], -q[:, 0]) + torch.pi) / (2 * torch.pi)
    assert torch.equal(keys.isnan(), ref.isnan()) and int(ref.isnan().sum()) == 2
    ok = ~ref.isnan()
    d = (keys[ok] - ref[ok]).abs()
    print("circle coordinate vs torch's chain on the device: max %.3e (%d special points, 4000 random)" % (d.max().item(), len(pts)))
    assert d.max().item() < 2.5e-7  # both sit within ~1e-7 of the exact value (tools/fit_circle_key.py)
    assert torch.equal(keys[:len(pts)][ok[:len(pts)]].round(decimals=4), ref[:len(pts)][ok[:len(pts)]].round(decimals=4))


@pytest.mark.parametrize("p", [1, 2, 3])
def test_euclid_sliced_w_matches_oracle(shwd, p):
    g = torch.Generator().manual_seed(p)
    x = torch.randn(700, 3, generator=g)
    y = torch.randn(700, 3, generator=g) + 0.5
    th = F.normalize(torch.randn(50, 3, generator=g), dim=-1)
    xr = x.clone().requires_grad_(True)
    ref = oracle.euclid_sliced_wasserstein(xr, y, th, p)
    ref.backward()
    xg = x.clone().to(dev()).requires_grad_(True)
    out = shwd.losses.sliced_wasserstein_distance(xg, y.to(dev()), p=p, device=dev(), projections=th.to(dev()))
    out.backward()
    assert out.item() == pytest.approx(ref.item(), rel=TOL)
    assert rel(xg.grad, xr.grad) < 2e-5


# ----------------------------------------------------------------------- full-size, size-independent properties ----
def test_full_size_properties_b32_n1024(shwd):
    """BASELINE config 2 (B=32, N=1024, L=100, eps=0.01, geodesic p=2): properties that need no CPU reference.
    (i) marginals: after the final v-update the column sums of P equal 1/M + 1e-8;  (ii) symmetry: cost(x,y) with
    rows/cols swapped is reproduced by swapping roles one half-step apart -- checked through the duals' consistency
    sum_i r_i == sum_j c_j;  (iii) permutation equivariance of the gradient;  (iv) bit-reproducibility."""
    torch.manual_seed(1234)
    B, N, L = 32, 1024, 100
    x = F.normalize(torch.randn(B, N, 3), dim=-1).to(dev())
    y = F.normalize(torch.randn(B, N, 3) + 0.1, dim=-1).to(dev())
    xg = x.clone().requires_grad_(True)
    res = shwd.entropic_ot(xg, y, "geodesic", 2.0, 0.01, L)
    res.cost.sum().backward()
    assert res.status() == 0
    assert torch.isfinite(res.cost).all() and torch.isfinite(xg.grad).all()
    # (iv) bit-reproducible
    xg2 = x.clone().requires_grad_(True)
    res2 = shwd.entropic_ot(xg2, y, "geodesic", 2.0, 0.01, L)
    res2.cost.sum().backward()
    assert torch.equal(res.cost, res2.cost) and torch.equal(xg.grad, xg2.grad)
    # (iii) permuting the points of x permutes its gradient and leaves the cost unchanged (to rounding)
    perm = torch.randperm(N, device=dev())
    xp = x[:, perm].clone().requires_grad_(True)
    resp = shwd.entropic_ot(xp, y, "geodesic", 2.0, 0.01, L)
    resp.cost.sum().backward()
    assert rel(resp.cost, res.cost) < TOL
    assert rel(xp.grad, xg.grad[:, perm]) < TOL
    # (i) column marginals of the plan on one pair (dense plan is small enough for a single pair)
    one = shwd.entropic_ot(x[:1], y[:1], "geodesic", 2.0, 0.01, L)
    P, C = one.dense()
    col = P.sum(dim=1)
    assert (col - (1.0 / N + 1e-8)).abs().max().item() < 2e-3 / N
    assert (P * C).sum().item() == pytest.approx(one.cost.item(), rel=1e-4)


# ------------------------------------------------------------- full-size parity: the oracle evaluated ON the device ----
# The oracle is device-agnostic torch code.  On the B200 (180 GB) it can hold the N x M tensors and the autograd tape
# of the reference's unrolled loop at BASELINE.json's full sizes, in float64 -- so the full-size configurations get a
# real parity check, not only size-independent properties.  (Pairs are independent: the kernel runs the whole batch,
# the oracle a subset of it, to bound the tape at ~20 GB.)
def _oracle_geodesic_on_device(x, y, eps, L, grad=True):
    xr = x.double().clone().requires_grad_(grad)
    yr = y.double().clone().requires_grad_(grad)
    with torch.set_grad_enabled(grad):
        cost = oracle.log_sinkhorn(oracle.sphere_map(xr), oracle.sphere_map(yr), "geodesic", 2, eps, L)
        if grad:
            cost.sum().backward()
    return cost.detach(), xr.grad, yr.grad


def test_cfg2_full_size_matches_oracle_on_device(shwd):
    """BASELINE config 2: B=32, N=M=1024, L=100, eps=0.01, geodesic p=2 on synthetic registration pairs (the bench
    workload).  Loss and both gradients w.r.t. the RAW clouds (through centring + normalisation) < 1e-5."""
    import bench
    B, N, L, eps, K = 32, 1024, 100, 0.01, 4
    tmpl, src = bench.registration_pairs(B, N, 1234, dev())
    xg, yg = tmpl.clone().requires_grad_(True), src.clone().requires_grad_(True)
    res = shwd.entropic_ot(xg, yg, "geodesic", 2.0, eps, L, center=True)
    res.cost.sum().backward()
    assert res.status() == 0
    cr, gxr, gyr = _oracle_geodesic_on_device(tmpl[:K], src[:K], eps, L)
    assert rel(res.cost[:K], cr) < TOL
    assert rel(xg.grad[:K], gxr) < TOL and rel(yg.grad[:K], gyr) < TOL, (rel(xg.grad[:K], gxr), rel(yg.grad[:K], gyr))
    torch.cuda.empty_cache()


def _ellipsoid(n, seed, biased=False):
    """Flow_ellipsoid.ipynb:106-150 (cell 3): (a,b,c) = (2,1,1); uniform angles, or theta from acos(N(0,0.25)) clipped."""
    g = torch.Generator().manual_seed(seed)
    phi = torch.rand(n, generator=g) * 2 * np.pi
    if biased:
        t = torch.acos((0.25 * torch.randn(n, generator=g)).clamp(-1, 1))
    else:
        t = torch.acos(torch.rand(n, generator=g) * 2 - 1)
    return torch.stack([2 * torch.sin(t) * torch.cos(phi), torch.sin(t) * torch.sin(phi), torch.cos(t)], -1)


def test_cfg4_n16384_matches_oracle_on_device(shwd):
    """BASELINE config 4: one ellipsoid pair, N=M=16384 (the N x M tensor is 2 GB in float64 -- never formed by the
    kernel).  L=3 keeps the oracle's tape at ~40 GB."""
    N, L, eps = 16384, 3, 0.01
    x, y = _ellipsoid(N, 1).to(dev()).unsqueeze(0), _ellipsoid(N, 2, biased=True).to(dev()).unsqueeze(0)
    xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    res = shwd.entropic_ot(xg, yg, "geodesic", 2.0, eps, L, center=True)
    res.cost.sum().backward()
    assert res.status() == 0
    cr, gxr, gyr = _oracle_geodesic_on_device(x, y, eps, L)
    assert rel(res.cost, cr) < TOL
    assert rel(xg.grad, gxr) < TOL and rel(yg.grad, gyr) < TOL, (rel(xg.grad, gxr), rel(yg.grad, gyr))
    del cr, gxr, gyr
    torch.cuda.empty_cache()


@pytest.mark.parametrize("B,N,L", [(256, 256, 10), (1, 32768, 2), (3, 5000, 4)])
def test_sweep_corners_match_oracle_on_device(shwd, B, N, L):
    """BASELINE config 5 (loss-kernel sweep N=256..65536, B=1..256): corner shapes, forward value against the float64
    oracle on the device (no tape: N=32768 alone is an 8.6 GB matrix), gradients where the tape fits."""
    g = torch.Generator().manual_seed(B + N)
    x = F.normalize(torch.randn(B, N, 3, generator=g), dim=-1).to(dev())
    y = F.normalize(torch.randn(B, N, 3, generator=g) + 0.2, dim=-1).to(dev())
    with_grad = B * N * N * L <= 256 * 256 * 256 * 10
    xg = x.clone().requires_grad_(True)
    res = shwd.entropic_ot(xg, y, "geodesic", 2.0, 0.02, L, center=True)
    res.cost.sum().backward()
    assert res.status() == 0 and torch.isfinite(xg.grad).all()
    cr, gxr, _ = _oracle_geodesic_on_device(x, y, 0.02, L, grad=with_grad)
    assert rel(res.cost, cr) < TOL
    if with_grad:
        assert rel(xg.grad, gxr) < TOL
    del cr
    torch.cuda.empty_cache()


def test_n65536_single_pair_properties(shwd):
    """The sweep's largest shape (N=M=65536, B=1): a 17 GB matrix in float32 that is never formed.  Finite, status 0,
    bit-reproducible, and invariant (to rounding) under a permutation of the points."""
    N, L = 65536, 2
    g = torch.Generator().manual_seed(5)
    x = F.normalize(torch.randn(1, N, 3, generator=g), dim=-1).to(dev())
    y = F.normalize(torch.randn(1, N, 3, generator=g) + 0.3, dim=-1).to(dev())
    xg = x.clone().requires_grad_(True)
    r1 = shwd.entropic_ot(xg, y, "geodesic", 2.0, 0.01, L, center=True)
    r1.cost.sum().backward()
    assert r1.status() == 0 and torch.isfinite(r1.cost).all() and torch.isfinite(xg.grad).all()
    xg2 = x.clone().requires_grad_(True)
    r2 = shwd.entropic_ot(xg2, y, "geodesic", 2.0, 0.01, L, center=True)
    r2.cost.sum().backward()
    assert torch.equal(r1.cost, r2.cost) and torch.equal(xg.grad, xg2.grad)
    perm = torch.randperm(N, device=dev())
    xp = x[:, perm].clone().requires_grad_(True)
    r3 = shwd.entropic_ot(xp, y, "geodesic", 2.0, 0.01, L, center=True)
    r3.cost.sum().backward()
    assert rel(r3.cost, r1.cost) < TOL and rel(xp.grad, xg.grad[:, perm]) < TOL


def test_chamfer_n16384_matches_oracle_on_device(shwd):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 16384, 3, generator=g).to(dev())
    y = (torch.randn(1, 12000, 3, generator=g) * 1.1 + 0.1).to(dev())
    xg, xr = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    out, _ = shwd.losses.chamfer_distance(xg, y)
    out.backward()
    ref, _ = oracle.chamfer_distance(xr, y)
    ref.backward()
    assert out.item() == pytest.approx(ref.item(), rel=TOL)
    assert rel(xg.grad, xr.grad) < TOL
    torch.cuda.empty_cache()


# ------------------------------------------------------------------ phi: fused Residual-flow stack ----
@pytest.mark.parametrize("n_flow_layer,shape", [(3, (4, 1024, 3)), (5, (1000, 3)), (1, (2, 77, 3)), (8, (3, 129, 3))])
def test_fused_residual_flow_matches_eager_modules(shwd, n_flow_layer, shape):
    """Norm_Flow_structure("Residual") (s2_wasserstein.py:144-163): the fused kernel against the eager torch modules of
    the same object (plain PyTorch fp32 reference of the same op) -- output, d/dx and d/d(every parameter)."""
    torch.manual_seed(n_flow_layer)
    phi = shwd.losses.Norm_Flow_structure(flow_name="Residual", n_flow_layer=n_flow_layer).to(dev())
    # (geom_p / lamb belong to the discarded log-determinant estimator: they mirror the reference's state and get no gradient)
    live = [(n, q) for n, q in phi.named_parameters() if not n.endswith(("geom_p", "lamb"))]
    with torch.no_grad():  # move the near-zero last layers and the Swish scales off their initial values
        for _, prm in live:
            prm.add_(0.05 * torch.randn_like(prm))
    x = (torch.randn(*shape) * 0.8).to(dev())
    w = torch.randn(*shape).to(dev())
    xe = x.clone().requires_grad_(True)
    ye = phi.forward_eager(xe)
    (ye * w).sum().backward()
    ge = [prm.grad.clone() for _, prm in live]
    gxe = xe.grad.clone()
    phi.zero_grad()
    xf = x.clone().requires_grad_(True)
    yf = phi(xf)
    (yf * w).sum().backward()
    assert yf.shape == ye.shape and rel(yf, ye) < TOL
    assert rel(xf.grad, gxe) < TOL
    for (name, prm), g in zip(live, ge):
        assert prm.grad is not None, name
        assert (prm.grad - g).norm().item() <= TOL * max(g.norm().item(), 1e-3), (name, prm.grad.norm().item(), g.norm().item())
    # bit-reproducible parameter gradients (fixed-order reduction, no float atomics)
    g1 = [prm.grad.clone() for _, prm in live]
    phi.zero_grad()
    xf2 = x.clone().requires_grad_(True)
    (phi(xf2) * w).sum().backward()
    assert all(torch.equal(a, prm.grad) for a, (_, prm) in zip(g1, live))


def test_optuna_flow_class_takes_the_fused_kernel_at_the_standard_shape(shwd):
    """Norm_Flow_structure_optuna (s2_wasserstein.py:171-201) with the width / depth Norm_Flow_structure itself builds runs the
    fused kernel -- same values and gradients as its eager modules; any other width / depth stays eager."""
    torch.manual_seed(3)
    phi = shwd.losses.Norm_Flow_structure_optuna(flow_name="Residual", n_flow_layer=2, Residual_hidden_units=8,
                                                 Residual_hidden_layers=7).to(dev())
    from shwd_b200.losses import flows
    assert flows.is_standard_residual_stack(phi.net)
    x = (torch.randn(3, 200, 3) * 0.8).to(dev())
    xe, xf = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ye = xe
    for f in phi.net:
        ye = f(ye)
    ye.square().sum().backward()
    ge = [q.grad.clone() for q in phi.parameters() if q.grad is not None]
    phi.zero_grad()
    yf = phi(xf)
    yf.square().sum().backward()
    gf = [q.grad.clone() for q in phi.parameters() if q.grad is not None]
    assert rel(yf, ye) < TOL and rel(xf.grad, xe.grad) < TOL and len(ge) == len(gf)
    assert all((a - b).norm().item() <= TOL * max(b.norm().item(), 1e-3) for a, b in zip(gf, ge))
    other = shwd.losses.Norm_Flow_structure_optuna(flow_name="Residual", n_flow_layer=2, Residual_hidden_units=4).to(dev())
    assert not flows.is_standard_residual_stack(other.net) and other(x).shape == x.shape


def _state(d, prefix):
    return {k[len(prefix):].replace("__", "."): torch.from_numpy(v) for k, v in d.items() if k.startswith(prefix)}


@pytest.mark.parametrize("name", ["Residual", "Planar"])
def test_phi_matches_vendored_normflows_fixture(shwd, name):
    """Norm_Flow_structure on the GPU (Residual: the fused resflow kernels; Planar: the fused planar kernels) loaded with the REFERENCE's
    state_dict -- same keys, shapes and order as train_W_COS.py:204 saves -- against outputs and autograd gradients frozen
    from the vendored normflows 1.7.2 (tests/golden/make_golden.py section 7): batched clouds and the (N,3) branch."""
    d = gold("flow_" + name.lower())
    phi = shwd.losses.Norm_Flow_structure(flow_name=name, n_flow_layer=3)
    res = phi.load_state_dict(_state(d, "sd__"), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    phi = phi.to(dev()).train()
    for tag in ("b", "u"):
        x = torch.from_numpy(d["x_" + tag]).to(dev()).requires_grad_(True)
        y = phi(x)
        assert name != "Planar" or type(y.grad_fn).__name__ == "PlanarFlowStackFnBackward"
        named = [(n, q) for n, q in phi.named_parameters() if q.dtype == torch.float32 and q.dim() > 0]
        gs = torch.autograd.grad((y * torch.from_numpy(d["w_" + tag]).to(dev())).sum(), [x] + [q for _, q in named], allow_unused=True)
        ey, egx = rel(y, torch.from_numpy(d["y_" + tag])), rel(gs[0], torch.from_numpy(d["gx_" + tag]))
        worst = 0.0
        for (n, q), g in zip(named, gs[1:]):
            want = torch.from_numpy(d["gp_%s__%s" % (tag, n.replace(".", "__"))])
            got = torch.zeros_like(want) if g is None else g.cpu()
            worst = max(worst, (got - want).norm().item() / max(want.norm().item(), 1e-3))
        print("phi %s (%s): y %.2e  d/dx %.2e  worst d/dparam %.2e" % (name, tag, ey, egx, worst))
        assert ey < TOL and egx < TOL and worst < 2e-5
    # a buffer copy after a forward must not leave the fused kernel on stale power-iteration vectors
    if name == "Residual":
        sd = _state(d, "sd__")
        key = "net.0.iresblock.nnet.net.1.u"
        sd[key] = -sd[key]  # u -> -u flips the sign of u^T W v: the normalisation factor becomes max(1, negative) = 1
        phi.load_state_dict(sd)
        x = torch.from_numpy(d["x_b"]).to(dev())
        assert rel(phi(x), phi.forward_eager(x)) < TOL


@pytest.mark.parametrize("shape,layers", [((32, 1024, 3), 3), ((3, 1000, 3), 8), ((4, 16, 3), 8), ((1, 7, 3), 1), ((2, 1, 3), 2), ((1, 20000, 3), 5),
                                          ((5000, 3), 3), ((1, 3), 1), ((257, 3), 8), ((70001, 3), 2)])
def test_fused_planar_stack_matches_eager_modules(shwd, shape, layers):
    """csrc/planar.cu against the eager PlanarFlow modules (pinned to the vendored normflows by the fixture test above) in
    float64: batched clouds (lin = w_c sum_n z_bnc, the reference's sum over dim 1) and the (N,3) branch, every stack depth
    the kernel templates cover, sizes that do not fill a CTA; value, d/dx, d/d(u, w, b); bit-reproducible."""
    torch.manual_seed(shape[-2] + layers)
    phi = shwd.losses.Norm_Flow_structure(flow_name="Planar", n_flow_layer=layers)
    with torch.no_grad():
        for q in phi.parameters():
            q.add_(0.2 * torch.randn_like(q))
        if len(shape) == 3 and shape[1] >= 100:
            # Batched clouds: every layer translates the cloud by d = u^ tanh(w S + b) and the column sum S moves by N d, so a
            # perturbation of S grows by 1 + N u^ w (1 - tanh^2) per layer.  With |u^| ~ 1 that is ~N: the float32 result of the
            # reference itself is then decided by the rounding of its sum.  For a comparison that means something, pick
            # parameters with u^ = eps w + (small vector orthogonal to w), eps ~ 1 / N: u = alpha w + q with alpha solving
            # alpha + (softplus(alpha |w|^2) - 1 - alpha |w|^2) / |w|^2 = eps (monotone in alpha -> bisection).
            N = shape[1]
            for f in phi.net:
                f.w.mul_(2.0 / N ** 0.5)  # w S stays inside tanh's active range: S ~ 0.7 sqrt(N)
                w = f.w.double().reshape(3)
                ww = (w * w).sum()
                eps = (2.0 * torch.rand(()).item() - 1.0) * 2.0 / (N * ww.item() ** 0.5)
                lo_a, hi_a = -1e4, 1e4
                for _ in range(200):
                    mid = 0.5 * (lo_a + hi_a)
                    val = mid + (F.softplus(mid * ww) - 1 - mid * ww) / ww - eps
                    lo_a, hi_a = (mid, hi_a) if val < 0 else (lo_a, mid)
                q = torch.randn(3, dtype=torch.float64) / N
                q = q - (q * w).sum() / ww * w
                f.u.copy_((0.5 * (lo_a + hi_a) * w + q).float().reshape(1, 3))
    x = torch.randn(*shape) * 0.7
    wgt = torch.randn(*shape)
    phi64 = shwd.losses.Norm_Flow_structure(flow_name="Planar", n_flow_layer=layers).double()
    phi64.load_state_dict({k: v.double() for k, v in phi.state_dict().items()})
    x64 = x.double().requires_grad_(True)
    y64 = phi64.forward_eager(x64)
    g64 = torch.autograd.grad((y64 * wgt.double()).sum(), [x64] + list(phi64.parameters()))
    phi32 = shwd.losses.Norm_Flow_structure(flow_name="Planar", n_flow_layer=layers)
    phi32.load_state_dict(phi.state_dict())
    x32 = x.clone().requires_grad_(True)
    y32 = phi32.forward_eager(x32)
    g32 = torch.autograd.grad((y32 * wgt).sum(), [x32] + list(phi32.parameters()))
    scale = max(b.norm().item() for b in g64[1:])

    def relc(a, b):  # saturated layers have exactly-zero float32 gradients against 1e-100 in float64: clamp the denominator
        return (a.detach().double().cpu() - b).norm().item() / max(b.norm().item(), 1e-3 * scale)

    # eager float32 against float64: the reference's own rounding
    floor_y, floor = rel(y32, y64), max(relc(a, b) for a, b in zip(g32, g64))
    phi = phi.to(dev())
    xg = x.to(dev()).requires_grad_(True)
    y = phi(xg)
    assert type(y.grad_fn).__name__ == "PlanarFlowStackFnBackward"
    g = torch.autograd.grad((y * wgt.to(dev())).sum(), [xg] + list(phi.parameters()))
    ey = rel(y, y64)
    eg = [relc(a, b) for a, b in zip(g, g64)]
    print("planar %s x%d: y %.2e d/dx %.2e worst d/dparam %.2e (eager f32 floor: y %.2e grads %.2e)"
          % (shape, layers, ey, eg[0], max(eg[1:]), floor_y, floor))
    assert ey < max(TOL, 8 * floor_y) and eg[0] < max(TOL, 8 * floor) and max(eg[1:]) < max(TOL, 16 * floor)
    y2 = phi(xg)
    g2 = torch.autograd.grad((y2 * wgt.to(dev())).sum(), [xg] + list(phi.parameters()))
    assert torch.equal(y, y2) and all(torch.equal(a, b) for a, b in zip(g, g2))
    with torch.no_grad():
        assert torch.equal(phi(xg), y)


def test_fused_planar_stack_guards(shwd):
    """Empty clouds, the optuna structure and deeper-than-eight stacks (eager), argument errors through the C ABI."""
    phi = shwd.losses.Norm_Flow_structure(flow_name="Planar", n_flow_layer=3).to(dev())
    for shape in ((0, 3), (4, 0, 3), (0, 16, 3)):
        x = torch.zeros(*shape, device=dev(), requires_grad=True)
        y = phi(x)
        assert y.shape == x.shape
        (y.sum() + sum(q.sum() for q in phi.parameters()) * 0).backward()
    opt = shwd.losses.Norm_Flow_structure_optuna(flow_name="Planar", n_flow_layer=2).to(dev())
    x = torch.randn(4, 50, 3, device=dev())
    assert type(opt(x.requires_grad_(True)).grad_fn).__name__ == "PlanarFlowStackFnBackward"
    deep = shwd.losses.Norm_Flow_structure(flow_name="Planar", n_flow_layer=9).to(dev())
    assert type(deep(x).grad_fn).__name__ != "PlanarFlowStackFnBackward" and deep(x).shape == x.shape
    lib = shwd._lib.lib()
    assert lib.shwd_planar_max_layers() == 8 and lib.shwd_planar_params_per_layer() == 7
    assert lib.shwd_planar_fwd(None, 0, 4, None, 1, None, None, None) == -1  # SHWD_ERR_INVALID_ARGUMENT


def test_euclid_sliced_w_matches_notebook_fixture(shwd):
    """The notebooks' `sliced_wasserstein_distance`, frozen by executing cell 5 of Flow_ellipsoid.ipynb from its own source."""
    d = gold("notebook_sliced_wasserstein")
    for p in (1, 2, 3):
        x = torch.from_numpy(d["x_p%d" % p]).to(dev()).requires_grad_(True)
        y = torch.from_numpy(d["y_p%d" % p]).to(dev()).requires_grad_(True)
        out = shwd.losses.sliced_wasserstein_distance(x, y, num_projection=40, p=p, device=dev(),
                                                      projections=torch.from_numpy(d["theta_p%d" % p]).to(dev()))
        out.backward()
        e = (abs(out.item() - float(d["loss_p%d" % p])) / float(d["loss_p%d" % p]), rel(x.grad, torch.from_numpy(d["gx_p%d" % p])),
             rel(y.grad, torch.from_numpy(d["gy_p%d" % p])))
        print("notebook SWD p=%d: loss %.2e gx %.2e gy %.2e" % ((p,) + e))
        assert e[0] < TOL and e[1] < 2e-5 and e[2] < 2e-5


@pytest.mark.parametrize("name", ["Residual", "Planar"])
def test_max_wrapper_step_matches_reference_fixture(shwd, name):
    """One training call of max_cos_disimilarity_wassersten_distance exactly as train_W_COS.py:404 builds it -- phi,
    Cos_disimilarity_W(device, p=2) with NO other argument (-> the exact solve the reference runs through ot.emd2), an
    optimiser on phi -- against the frozen run of the unmodified reference wrapper (two SGD ascent steps, then the outer
    distance): value, both transformed clouds, the gradient reaching the second cloud, phi after the ascent, and the
    'test' branch."""
    d = gold("max_wrapper_" + name.lower())
    L = shwd.losses

    def run(eager):
        phi = L.Norm_Flow_structure(flow_name=name, n_flow_layer=int(d["n_flow_layer"]))
        phi.load_state_dict(_state(d, "sd0__"))
        phi = phi.to(dev())
        if eager:
            phi.forward = phi.forward_eager
        op = torch.optim.SGD(list(phi.parameters()), lr=float(d["lr"]))
        csw = L.Cos_disimilarity_W(dev(), p=2)
        assert csw.solver == "auto"
        crit = L.max_cos_disimilarity_wassersten_distance(phi=phi, CSW=csw, device=dev(), phi_op=op, max_iter=int(d["max_iter"]),
                                                           lam=float(d["lam"]))
        second = torch.from_numpy(d["second"]).to(dev()).requires_grad_(True)
        first = torch.from_numpy(d["first"]).to(dev())
        cswd, ft, st = crit(first, second, "train")
        (g2,) = torch.autograd.grad(cswd, second)
        e = (abs(cswd.item() - float(d["cswd"])) / float(d["cswd"]), rel(ft, torch.from_numpy(d["first_t"])),
             rel(st, torch.from_numpy(d["second_t"])), rel(g2, torch.from_numpy(d["g_second"])))
        return e, phi, crit, first, second

    e, phi, crit, first, second = run(False)
    print("max wrapper %s: cswd %.2e first_t %.2e second_t %.2e d/dsecond %.2e" % ((name,) + e))
    slack = 1.0
    if name == "Planar":
        # The fixture's first cloud is centred: its column sums -- the ONLY input of the reference's batched Planar layers
        # (planar.py:52 sums over dim 1 = the points) -- are float32 rounding residue (~1e-7), amplified by ~N u^ w per layer.
        # How far torch's own CUDA kernels land from the CPU-generated fixture is therefore the yardstick, measured here.
        ee = run(True)[0]
        print("   (eager torch modules on the same GPU: cswd %.2e first_t %.2e second_t %.2e d/dsecond %.2e)" % ee)
        slack = 2.0
        assert all(a < max(b * 3, t) for a, b, t in zip(e, ee, (TOL, TOL, TOL, 5e-5)))
    assert e[0] < TOL and e[1] < TOL and e[2] < TOL and e[3] < 5e-5 * slack
    for k, v in _state(d, "sd1__").items():
        if v.dtype == torch.float32 and v.dim() > 0 and "last_" not in k and not k.endswith("scale"):
            got = phi.state_dict()[k].cpu()
            assert torch.allclose(got, v, rtol=2e-4, atol=2e-6), (k, (got - v).abs().max().item())
    ctest, _, _ = crit(first, second.detach(), "test")
    assert ctest.item() == pytest.approx(float(d["cswd_test"]), rel=TOL)


class _FixedFramesSSW:
    """The SSW callable of tests/golden/make_golden.py (fixed cycle of frames instead of random ones), on the drop-in's sliced_cost."""

    def __init__(self, sliced_cost, Us):
        self.sliced_cost, self.Us, self.k = sliced_cost, Us, 0

    def __call__(self, Xs, Xt, num_projections, device, p=2):
        U = self.Us[self.k % len(self.Us)]
        self.k += 1
        return self.sliced_cost(Xs, Xt, U.to(Xs.device), p=p)


@pytest.mark.parametrize("p", [2, 1])
def test_max_ssw_wrapper_step_matches_reference_fixture(shwd, p, capsys):
    """One training call of max_spherical_wassersten_distance with transform_to_sphere (max_spherical_sliced_w.py:334-350,
    498-536; SURVEY.md 8f #4 tail) against the frozen run of the unmodified reference: two SGD ascent steps, the outer sum of
    per-pair sliced costs (33 against 40 points), both transformed clouds, the gradient reaching the second cloud, phi after the
    ascent and the 'test' branch."""
    d = gold("max_ssw_wrapper")
    L = shwd.losses
    phi = L.transform_to_sphere()
    phi.load_state_dict(_state(d, "sd0_p%d__" % p))
    phi = phi.to(dev())
    op = torch.optim.SGD(phi.parameters(), lr=float(d["lr"]))
    Us = torch.from_numpy(d["Us_p%d" % p]).to(dev())
    crit = L.max_spherical_wassersten_distance(int(d["num_projections"]), phi, _FixedFramesSSW(L.sliced_cost, Us), op, p=p,
                                               max_iter=int(d["max_iter"]), device=dev())
    first = torch.from_numpy(d["first_p%d" % p]).to(dev())
    second = torch.from_numpy(d["second_p%d" % p]).to(dev()).requires_grad_(True)
    val, ft, st = crit(first, second, "train")
    (g2,) = torch.autograd.grad(val, second)
    want = float(d["ssw_p%d" % p])
    e = (abs(val.item() - want) / want, rel(ft, torch.from_numpy(d["first_t_p%d" % p])), rel(st, torch.from_numpy(d["second_t_p%d" % p])),
         rel(g2, torch.from_numpy(d["g_second_p%d" % p])))
    with capsys.disabled():
        print("max SSW wrapper p=%d: ssw %.2e first_t %.2e second_t %.2e d/dsecond %.2e" % ((p,) + e))
    assert max(e) < TOL
    for k, v in _state(d, "sd1_p%d__" % p).items():
        got = phi.state_dict()[k].cpu()
        assert torch.allclose(got, v, rtol=2e-4, atol=2e-6), (k, (got - v).abs().max().item())
    vt, _, _ = crit(first, second.detach(), "test")
    assert vt.item() == pytest.approx(float(d["ssw_test_p%d" % p]), rel=TOL)
    # the one-call extension: same estimator, frames shared by the pairs of one evaluation
    torch.manual_seed(0)
    shared = L.max_spherical_wassersten_distance(256, phi, None, op, p=p, max_iter=0, device=dev(), shared_frames=True)
    vs, _, _ = shared(first, second.detach(), "test")
    per_pair = L.max_spherical_wassersten_distance(256, phi, L.sliced_wasserstein_sphere, op, p=p, max_iter=0, device=dev())
    vp, _, _ = per_pair(first, second.detach(), "test")
    assert vs.item() == pytest.approx(vp.item(), rel=0.25)  # two Monte-Carlo estimates of the same quantity


def test_batched_fast_variant_matches_reference_fixture(shwd, capsys):
    """max_spherical_sliced_w_fast.py: `sliced_cost` with per-pair frames (:258-295; value of shape (1,) = sum over pairs, both
    gradients, p = 2 and 3) and one training call of max_spherical_wassersten_distance_fast (:346-382), frozen from the
    unmodified reference module; the drop-in module name resolves."""
    d = gold("ssw_fast")
    L = shwd.losses
    for p in (2, 3):
        x = torch.from_numpy(d["x_p%d" % p]).to(dev()).requires_grad_(True)
        y = torch.from_numpy(d["y_p%d" % p]).to(dev()).requires_grad_(True)
        w = L.sliced_cost_fast(x, y, torch.from_numpy(d["Us_p%d" % p]).to(dev()), p=p)
        assert w.shape == (1,)
        gx, gy = torch.autograd.grad(w.sum(), (x, y))
        e = (rel(w, torch.from_numpy(d["w_p%d" % p])), rel(gx, torch.from_numpy(d["gx_p%d" % p])), rel(gy, torch.from_numpy(d["gy_p%d" % p])))
        with capsys.disabled():
            print("sliced_cost (fast, per-pair frames) p=%d: w %.2e gx %.2e gy %.2e" % ((p,) + e))
        assert e[0] < TOL and e[1] < 2e-5 and e[2] < 2e-5
    phi = L.transform_to_sphere_fast()
    phi.load_state_dict(_state(d, "sd0__"))
    phi = phi.to(dev())
    op = torch.optim.SGD(phi.parameters(), lr=0.05)
    crit = L.max_spherical_wassersten_distance_fast(16, phi, _FixedFramesSSW(L.sliced_cost_fast, torch.from_numpy(d["Uc"]).to(dev())), op,
                                                    p=2, max_iter=2, device=dev())
    second = torch.from_numpy(d["second"]).to(dev()).requires_grad_(True)
    val, ft, st = crit(torch.from_numpy(d["first"]).to(dev()), second, "train")
    (g2,) = torch.autograd.grad(val.sum(), second)
    e = (rel(val, torch.from_numpy(d["ssw"])), rel(ft, torch.from_numpy(d["first_t"])), rel(st, torch.from_numpy(d["second_t"])),
         rel(g2, torch.from_numpy(d["g_second"])))
    # the clouds nearly coincide after phi, so the gradient is a sum of small differences: the reference's float32 result sits
    # 3.7e-5 from the same code run in float64 (frozen next to it); measure this path against the float64 gradient as well
    g64 = torch.from_numpy(d["g_second_f64"])
    floor, e64 = rel(torch.from_numpy(d["g_second"]), g64), rel(g2, g64)
    with capsys.disabled():
        print("max SSW wrapper (fast): ssw %.2e first_t %.2e second_t %.2e d/dsecond %.2e (vs float64 reference %.2e; reference "
              "float32 vs float64 %.2e)" % (e + (e64, floor)))
    assert max(e[:3]) < TOL and e64 < max(TOL, 3 * floor) and e[3] < max(TOL, 4 * floor)
    for k, v in _state(d, "sd1__").items():
        assert torch.allclose(phi.state_dict()[k].cpu(), v, rtol=2e-4, atol=2e-6), k
    out = L.sliced_wasserstein_sphere_fast(x.detach(), y.detach(), 64, dev(), p=2)
    assert out.shape == (1,) and torch.isfinite(out).all() and out.item() > 0
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin"))
    try:
        import importlib
        m = importlib.import_module("losses.max_spherical_sliced_w_fast")
        assert m.sliced_cost is L.sliced_cost_fast and m.max_spherical_wassersten_distance_fast is L.max_spherical_wassersten_distance_fast
        m2 = importlib.import_module("losses.max_spherical_sliced_w")
        assert m2.max_spherical_wassersten_distance is L.max_spherical_wassersten_distance
    finally:
        sys.path.pop(0)


@pytest.mark.parametrize("B,n,m,P", [(3, 40, 33, 16), (8, 4096, 4096, 24), (5, 1000, 1024, 40), (2, 5000, 4500, 8)])
@pytest.mark.parametrize("p", [1, 2])
def test_per_pair_frames_fused_call_equals_the_per_pair_calls(shwd, B, n, m, P, p, capsys):
    """Frames (B,P,3,2), one set per pair (max_spherical_sliced_w_fast.py:298-319): ONE fused call through the shwd_*_pp entry
    points against B single-pair calls with the pair's own frames (the fixture-pinned path): values and both gradients, on the
    fused-key sort, the wide projection backward and the project-then-sort path of long rows."""
    g = torch.Generator().manual_seed(B * n + p)
    x = F.normalize(torch.randn(B, n, 3, generator=g), dim=-1).to(dev())
    y = F.normalize(torch.randn(B, m, 3, generator=g), dim=-1).to(dev())
    Us = shwd.losses.sliced.stiefel_frames(torch.randn(B, P, 3, 2, generator=g)).to(dev())
    wgt = torch.rand(B, generator=g).to(dev()) + 0.5
    xa, ya = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    wa = shwd.ops.spherical_sliced_w1(xa, ya, Us) if p == 1 else shwd.ops.spherical_sliced_wp(xa, ya, Us, float(p))
    ga = torch.autograd.grad((wa * wgt).sum(), (xa, ya))
    xb, yb = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    wb = torch.stack([shwd.losses.sliced_cost(xb[i], yb[i], Us[i], p=p) for i in range(B)])
    gb = torch.autograd.grad((wb * wgt).sum(), (xb, yb))
    e = (rel(wa, wb), rel(ga[0], gb[0]), rel(ga[1], gb[1]))
    with capsys.disabled():
        print("per-pair frames B=%d n=%d m=%d P=%d p=%d: w %.1e gx %.1e gy %.1e%s" % (
            B, n, m, P, p, e[0], e[1], e[2], "  (bit-identical)" if torch.equal(wa, wb) and torch.equal(ga[0], gb[0]) else ""))
    assert torch.equal(wa, wb) and torch.equal(ga[0], gb[0]) and torch.equal(ga[1], gb[1])  # same kernels, same bits
    assert shwd.losses.sliced_cost_fast(x, y, Us, p=p).item() == pytest.approx(wb.sum().item(), rel=1e-6)
    with pytest.raises(ValueError):
        shwd.ops.spherical_sliced_wp(x, y, Us[:-1], 2.0) if B > 1 else shwd.ops.spherical_sliced_wp(x, y, Us.repeat(2, 1, 1, 1), 2.0)


def test_mini_batch_mssw_wrapper_matches_reference_fixture(shwd, capsys):
    """One training call of max_spherical_wassersten_distance_Residual (mini_batch_Residual_MSSW.py:413-452) with its own sphere
    map (Planar flows on R^2), the mini-batches drawn by np.random.choice after np.random.seed(12) as in the frozen run of the
    unmodified reference, the frames of every SSW call fixed; then the one-fused-call route with the package's own SSW."""
    from shwd_b200.losses import mini_batch_mssw as M
    d = gold("mini_batch_mssw")
    phi = M.transform_to_sphere("Planar", n_flow_layer=2)
    phi.load_state_dict(_state(d, "w_sd0__"))
    phi = phi.to(dev())
    op = torch.optim.SGD(phi.parameters(), lr=0.05)
    crit = M.max_spherical_wassersten_distance_Residual(24, phi, op, SSW=_FixedFramesSSW(shwd.losses.sliced_cost, torch.from_numpy(d["Us"]).to(dev())),
                                                        p=2, max_iter=2, psi_minibatch_size=2, device=dev(), verbose=False)
    first = torch.from_numpy(d["first"]).to(dev())
    second = torch.from_numpy(d["second"]).to(dev()).requires_grad_(True)
    np.random.seed(12)
    val, ft, st = crit(first, second, "train")
    (g2,) = torch.autograd.grad(val, second)
    want = float(d["ssw"])
    e = (abs(val.item() - want) / want, rel(ft, torch.from_numpy(d["first_t"])), rel(st, torch.from_numpy(d["second_t"])),
         rel(g2, torch.from_numpy(d["g_second"])))
    with capsys.disabled():
        print("mini-batch max-SSW wrapper: ssw %.2e first_t %.2e second_t %.2e d/dsecond %.2e" % e)
    # measured: value 1.6e-5, clouds 6.8e-6 / 6.3e-6, gradient 9.7e-4 -- the value is 2.4e-5 = a mean of squared differences of
    # ~5e-3 between nearly coinciding clouds, so its gradient is a sum of cancelling float32 terms in the reference as well
    assert e[0] < 1e-4 and e[1] < TOL and e[2] < TOL and e[3] < 4e-3
    # phi after the two ascent steps: the steps themselves are small here (lr x a gradient of nearly coinciding clouds), so the
    # comparison is on the UPDATE, relative to the largest one
    sd0, sd1 = _state(d, "w_sd0__"), _state(d, "w_sd1__")
    upd = max((sd1[k] - sd0[k]).abs().max().item() for k in sd1)
    worst = max((phi.state_dict()[k].cpu() - sd1[k]).abs().max().item() for k in sd1)
    with capsys.disabled():
        print("   phi after the ascent: largest update %.2e, largest deviation from the reference's %.2e" % (upd, worst))
    assert worst < 0.05 * upd + 2e-6
    fused = M.max_spherical_wassersten_distance_Residual(64, phi, op, p=2, max_iter=1, psi_minibatch_size=3, device=dev(), verbose=False)
    v2, _, _ = fused(first, second.detach(), "train")
    assert v2.dim() == 0 and torch.isfinite(v2) and v2.item() > 0


def test_max_wrapper_with_fused_phi_runs_the_reference_step(shwd):
    """One training step of max_cos_disimilarity_wassersten_distance (s2_wasserstein.py:234-262) with the fused phi:
    the inner ascent changes phi's parameters, the outer loss back-propagates to the cloud."""
    torch.manual_seed(0)
    phi = shwd.losses.Norm_Flow_structure(flow_name="Residual", n_flow_layer=3).to(dev())
    opt = torch.optim.Adam(phi.parameters(), lr=1e-2)
    crit = shwd.losses.max_cos_disimilarity_wassersten_distance(phi, shwd.losses.Geodesic_distance_W(dev(), p=2, max_iter=20),
                                                                dev(), opt, max_iter=1, lam=0.1)
    a = F.normalize(torch.randn(4, 256, 3), dim=-1).to(dev())
    b = (F.normalize(torch.randn(4, 256, 3), dim=-1) * 1.1).to(dev()).requires_grad_(True)
    before = [p.detach().clone() for p in phi.parameters()]
    loss, a_t, b_t = crit(a, b, "train")
    loss.backward()
    assert torch.isfinite(loss) and b.grad is not None and torch.isfinite(b.grad).all() and b.grad.abs().sum() > 0
    assert any(not torch.equal(p0, p1) for p0, p1 in zip(before, phi.parameters()))
    assert a_t.shape == a.shape and b_t.shape == b.shape


def test_cos_disimilarity_w_full_size_matches_oracle_on_device(shwd):
    """What train_W_COS.py instantiates (:393): Cos_disimilarity_W(p=2), cost sum_k |x_k - y_k|^2, at B=32, N=1024,
    L=100, eps=0.01 on phi-like (un-normalised) clouds.  Bound: max(1e-5, 8 x the reference's own f32-vs-f64 distance)."""
    import bench
    B, N, L, eps, K = 32, 1024, 100, 0.01, 2
    tmpl, src = bench.registration_pairs(B, N, 99, dev())
    tmpl = tmpl - tmpl.mean(1, keepdim=True)
    src = src - src.mean(1, keepdim=True)
    xg, yg = tmpl.clone().requires_grad_(True), src.clone().requires_grad_(True)
    res = shwd.entropic_ot(xg, yg, "sqeuclid", 2.0, eps, L)
    res.cost.sum().backward()
    assert res.status() == 0
    outs = {}
    for dt in (torch.float64, torch.float32):
        xr, yr = tmpl[:K].to(dt).requires_grad_(True), src[:K].to(dt).requires_grad_(True)
        c = oracle.log_sinkhorn(xr, yr, "sqeuclid", 2, eps, L)
        c.sum().backward()
        outs[dt] = (c.detach(), xr.grad, yr.grad)
        del c
    floor = max(rel(a, b) for a, b in zip(outs[torch.float32], outs[torch.float64]))
    bound = max(TOL, 8 * floor)
    c64, gx64, gy64 = outs[torch.float64]
    assert rel(res.cost[:K], c64) < bound
    assert rel(xg.grad[:K], gx64) < bound and rel(yg.grad[:K], gy64) < bound, (rel(xg.grad[:K], gx64), rel(yg.grad[:K], gy64), floor)
    torch.cuda.empty_cache()


# ------------------------------------------------ exact EMD (what the reference's W_COS path solves with ot.emd2, 8f #2) ----
def _lsa(C):
    from scipy.optimize import linear_sum_assignment
    return linear_sum_assignment(C.double().numpy())[1]


@pytest.mark.parametrize("B,N,kind,p", [(3, 7, "sqeuclid", 2), (2, 64, "geodesic", 2), (2, 257, "sqeuclid", 2), (2, 300, "geodesic", 1),
                                        (1, 1024, "geodesic", 2), (2, 1, "sqeuclid", 2), (2, 200, "sqeuclid", 1), (1, 500, "one_minus_cos", 2)])
def test_exact_assignment_is_the_lp_optimum(shwd, B, N, kind, p):
    """The auction kernel returns an optimal permutation: identical to scipy's linear_sum_assignment on the reference's
    own cost matrix, or -- for costs with exact ties between optimal assignments, which the L1 cost produces (swapping
    two matches leaves sum |.| unchanged whenever the coordinate intervals nest) -- of exactly the same total cost."""
    g = torch.Generator().manual_seed(B * 1000 + N)
    x = torch.randn(B, N, 3, generator=g)
    y = torch.randn(B, N, 3, generator=g) * 0.9 + 0.1
    if kind in ("geodesic", "one_minus_cos"):
        x, y = F.normalize(x, dim=-1), F.normalize(y, dim=-1)
    sig, prices, rounds, status = shwd.ops.exact_assignment(x.to(dev()), y.to(dev()), kind, float(p), return_info=True)
    assert int(status.item()) == 0
    C = oracle.cost_matrix(x, y, kind, p)
    for b in range(B):
        ours, ref = sig[b].cpu().numpy(), _lsa(C[b])
        Cb = C[b].double().numpy()
        excess = (Cb[np.arange(N), ours].sum() - Cb[np.arange(N), ref].sum()) / max(Cb[np.arange(N), ref].sum(), 1e-30)
        assert sorted(ours.tolist()) == list(range(N))
        assert np.array_equal(ours, ref) or abs(excess) <= 1e-13, (b, int(rounds[b]), "relative excess cost %.3e" % excess)


def test_exact_assignment_edge_cases(shwd):
    """The candidate-list auction on the inputs that stress it: the largest supported clouds (8-entry lists), clouds with
    many duplicated points (exact ties: whole groups of persons see identical objects), a cost matrix with negative
    entries and one whose entries are all equal."""
    from scipy.optimize import linear_sum_assignment
    lib = shwd._lib.lib()
    N = lib.shwd_exact_assignment_max_points()
    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(1, N, 3, generator=g), torch.randn(1, N, 3, generator=g) * 0.7 + 0.3
    sig, _, _, status = shwd.ops.exact_assignment(x.to(dev()), y.to(dev()), "sqeuclid", 2.0, return_info=True)
    C = oracle.cost_matrix(x, y, "sqeuclid", 2)[0].double().numpy()
    ref = linear_sum_assignment(C)[1]
    ours = sig[0].cpu().numpy()
    assert int(status.item()) == 0 and sorted(ours.tolist()) == list(range(N))
    assert np.array_equal(ours, ref) or abs(C[np.arange(N), ours].sum() - C[np.arange(N), ref].sum()) <= 1e-12 * C[np.arange(N), ref].sum()
    # duplicated points: 40 distinct locations, each 8 times, in both clouds
    n = 320
    x = torch.randn(1, 40, 3, generator=g).repeat(1, 8, 1)
    y = (torch.randn(1, 40, 3, generator=g) * 0.5).repeat(1, 8, 1)[:, torch.randperm(n, generator=g)]
    for kind in ("sqeuclid", "geodesic"):
        sig, _, _, status = shwd.ops.exact_assignment(x.to(dev()), y.to(dev()), kind, 2.0, return_info=True)
        xn, yn = (F.normalize(x, dim=-1), F.normalize(y, dim=-1)) if kind == "geodesic" else (x, y)
        C = oracle.cost_matrix(xn, yn, kind, 2)[0].double().numpy()
        r, c = linear_sum_assignment(C)
        ours = sig[0].cpu().numpy()
        assert int(status.item()) == 0 and sorted(ours.tolist()) == list(range(n))
        assert abs(C[np.arange(n), ours].sum() - C[r, c].sum()) <= 1e-9 * max(C[r, c].sum(), 1e-30), kind
    # explicit matrices: negative entries (signed fixed point, no shift) and a constant matrix
    M = torch.randn(2, 150, 150, generator=g)
    v = shwd.exact_emd2_dense(M.to(dev()))
    for b in range(2):
        r, c = linear_sum_assignment(M[b].double().numpy())
        ref_v = M[b].double()[r, c].sum().item() / 150
        assert abs(v[b].item() - ref_v) <= 2e-6 * abs(ref_v) + 1e-7
    const = torch.full((1, 33, 33), 2.5)
    sig = shwd.exact_assignment_dense(const.to(dev()))
    assert sorted(sig[0].cpu().tolist()) == list(range(33))
    assert shwd.exact_emd2_dense(const.to(dev())).item() == pytest.approx(2.5)


def test_exact_assignment_randomised_against_scipy(shwd):
    """tools/fuzz_auction.py for ten seconds: random sizes, cost kinds, duplicated / clustered clouds, tied / negative
    matrices -- the optimal VALUE against scipy's exact assignment (4459 cases ran clean when this was written)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_auction.py"), "10", "7"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and "fuzz ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("n,m", [(6, 4), (30, 20), (5, 7), (64, 32), (1, 3)])
def test_exact_solver_rectangular_is_the_transport_lp_optimum(shwd, n, m):
    """Clouds of different sizes (train_W_COS.py:292-293 exposes --source_p_n / --target_p_n; ot.emd2 takes any n, m): the
    uniform transport LP, solved as the assignment of lcm(n, m) copies, against scipy's LP solver on the same cost -- through
    the loss object (value and both gradients, POT semantics d emd2 / dC = plan) and through the ot.emd2 drop-in."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin_ot"))
    import ot
    g = torch.Generator().manual_seed(n * 100 + m)
    x = torch.randn(2, n, 3, generator=g)
    y = torch.randn(2, m, 3, generator=g) * 0.8 + 0.2
    xr, yr = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    C = oracle.cost_matrix(xr, yr, "sqeuclid", 2)
    ref = 0
    for b in range(2):
        val, plan = oracle.exact_emd2_lp(C[b])
        ref = ref + torch.pow((plan.float() * C[b]).sum(), 0.5)
    ref = ref / 2
    ref.backward()
    xg, yg = x.clone().to(dev()).requires_grad_(True), y.clone().to(dev()).requires_grad_(True)
    out = shwd.losses.Cos_disimilarity_W(dev(), p=2)(xg, yg)  # the reference-style constructor: exact solve
    out.backward()
    assert out.item() == pytest.approx(ref.item(), rel=5e-6)
    # (the LP optimum is generically unique, so the plans -- hence the gradients -- agree)
    assert rel(xg.grad, xr.grad) < 1e-4 and rel(yg.grad, yr.grad) < 1e-4
    M = C[0].detach().float().to(dev()).requires_grad_(True)
    v = ot.emd2(torch.full((n,), 1.0 / n, device=dev()), torch.full((m,), 1.0 / m, device=dev()), M)
    val0, plan0 = oracle.exact_emd2_lp(C[0])
    assert v.item() == pytest.approx(val0, rel=5e-6)
    v.backward()
    assert torch.allclose(M.grad.cpu().double(), plan0, atol=1e-6)


@pytest.mark.parametrize("kind", ["sqeuclid", "geodesic"])
def test_exact_solver_value_and_gradient_follow_pot_semantics(shwd, kind):
    """Cos_disimilarity_W / Geodesic_distance_W with solver="exact": the value mean_b emd2_b^(1/p) and the gradient POT's
    torch backend attaches (d emd2 / dC = optimal plan, s2_wasserstein.py:41-44), computed by the oracle on the dense
    cost matrix with scipy's exact assignment."""
    B, N, p = 3, 200, 2
    g = torch.Generator().manual_seed(17)
    x = torch.randn(B, N, 3, generator=g)
    y = torch.randn(B, N, 3, generator=g) * 0.8 + 0.2
    if kind == "geodesic":
        x, y = F.normalize(x, dim=-1), F.normalize(y, dim=-1) * 1.3  # the cost normalises internally
    xr, yr = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    C = oracle.cost_matrix(xr, yr, kind, p)
    ref = 0
    for b in range(B):
        _, plan = oracle.exact_emd2(C[b])
        ref = ref + torch.pow((plan.float() * C[b]).sum(), 1.0 / p)
    ref = ref / B
    ref.backward()
    cls = shwd.losses.Cos_disimilarity_W if kind == "sqeuclid" else shwd.losses.Geodesic_distance_W
    xg, yg = x.clone().to(dev()).requires_grad_(True), y.clone().to(dev()).requires_grad_(True)
    out = cls(dev(), p=p, solver="exact")(xg, yg)
    out.backward()
    assert out.item() == pytest.approx(ref.item(), rel=2e-6)
    assert rel(xg.grad, xr.grad) < TOL and rel(yg.grad, yr.grad) < TOL


def test_exact_solver_matches_reference_wrapper_fixture(shwd):
    """The frozen outputs of the reference's own Geodesic_distance_W / Cos_disimilarity_W (run with a scipy-backed ``ot``
    shim, tests/golden/make_golden.py) on the fixture clouds."""
    d = gold("cost_matrices")
    x, y = torch.from_numpy(d["x"]).to(dev()), torch.from_numpy(d["y"]).to(dev())
    g = shwd.losses.Geodesic_distance_W(dev(), p=2, solver="exact")(x, y)
    c = shwd.losses.Cos_disimilarity_W(dev(), p=2, solver="exact")(x, y)
    assert g.item() == pytest.approx(float(d["exact_emd_geodesic_p2"]), rel=TOL)
    assert c.item() == pytest.approx(float(d["exact_emd_sqeuclid_p2"]), rel=TOL)


def test_exact_solver_training_shape_properties(shwd):
    """B=32, N=1024 (the training shape): permutations, deterministic, no worse than the identity matching, and certified
    optimal by LP duality (primal value == dual value of the returned prices)."""
    import bench
    tmpl, src = bench.registration_pairs(32, 1024, 7, dev())
    s1, _, rounds, status = shwd.ops.exact_assignment(tmpl, src, "sqeuclid", 2.0, return_info=True)
    assert int(status.item()) == 0
    assert torch.equal(torch.sort(s1, dim=1)[0], torch.arange(1024, device=dev()).expand(32, -1))
    s2 = shwd.ops.exact_assignment(tmpl, src, "sqeuclid", 2.0)
    assert torch.equal(s1, s2)
    emd = shwd.ops.exact_emd2(tmpl, src, "sqeuclid", 2.0)
    ident = ((tmpl - src) ** 2).sum(-1).mean(1)
    assert (emd <= ident * (1 + 1e-6)).all()
    # optimality certificate from the returned duals: with profits pi_i = max_j (-C_ij - p_j), sum_i pi_i + sum_j p_j is
    # an upper bound of -(min cost); it must meet the primal value to ~n * eps_final
    sig, prices, _, _ = shwd.ops.exact_assignment(tmpl[:2], src[:2], "sqeuclid", 2.0, return_info=True)
    C = ((tmpl[:2].unsqueeze(2) - src[:2].unsqueeze(1)).abs() ** 2).sum(-1).double()
    profit = (-C - prices.unsqueeze(1)).max(dim=2).values
    dual = -(profit.sum(1) + prices.sum(1))
    primal = torch.gather(C, 2, sig.unsqueeze(-1)).sum((1, 2))
    # (C here is torch's float32 evaluation; the kernel's differs by float32 rounding, ~1e-7 per entry)
    assert ((primal - dual).abs() <= 1e-6 * primal.abs()).all(), (primal, dual)


# ------------------------------------------------------------------------------- data side: rigid transform (8f #3) ----
def test_rigid_transform_matches_reference_fixture_and_oracle(shwd):
    d = gold("rigid_transform")
    src, poses = torch.from_numpy(d["src"]), torch.from_numpy(d["poses"])
    out, rot, trans = shwd.data.rigid_transform(src.to(dev()), poses.to(dev()))
    assert (out.cpu() - torch.from_numpy(d["out"])).abs().max().item() < 1e-6
    assert (rot.cpu() - torch.from_numpy(d["igt_rotation"])).abs().max().item() < 1e-6
    assert torch.equal(trans.cpu(), torch.from_numpy(d["igt_translation"]))
    # larger batch against the oracle; R reproduces the transform; noise has the requested moments and is reproducible
    g = torch.Generator().manual_seed(3)
    big = torch.randn(64, 1000, 3, generator=g)
    P = shwd.data.random_poses(64, 45, 1, np.random.RandomState(7))
    o2, r2, t2 = shwd.data.rigid_transform(big.to(dev()), P)
    ro, rr, rt = oracle.data.rigid_transform(big, P)
    assert rel(o2, ro) < 1e-6 and rel(r2, rr) < 1e-6
    assert rel(torch.einsum("bij,bnj->bni", r2.cpu(), big) + t2.cpu(), ro) < 1e-6
    n1, _, _ = shwd.data.rigid_transform(big.to(dev()), P, noise_std=0.02, seed=5)
    n2, _, _ = shwd.data.rigid_transform(big.to(dev()), P, noise_std=0.02, seed=5)
    assert torch.equal(n1, n2)
    resid = torch.einsum("bji,bnj->bni", r2, n1 - o2)  # rotate the residual back: the noise that was added to the source
    assert abs(resid.mean().item()) < 5e-4 and resid.std().item() == pytest.approx(0.02, rel=0.02)
    ds = shwd.data.DeviceRegistrationPairs(big.to(dev()), big.to(dev()), seed=11)
    tgt, srcs, R, T = ds.batch([3, 1, 4])
    assert tgt.shape == (3, 1000, 3) and srcs.shape == (3, 1000, 3) and R.shape == (3, 3, 3) and T.shape == (3, 1, 3)


# ------------------------------------------------------------------------------------- run-to-run reproducibility ----
def test_sliced_and_chamfer_are_bit_reproducible(shwd):
    """No float atomics and fixed-order reductions on these paths: two runs on the same inputs give the same bits (this
    is also the race check for the shared-memory exchanges of circular_w1 / circular_wp / the projection backward --
    compute-sanitizer is not available on the GPU pool)."""
    g = torch.Generator().manual_seed(77)
    x = F.normalize(torch.randn(2, 3000, 3, generator=g), dim=-1).to(dev())
    y = F.normalize(torch.randn(2, 2500, 3, generator=g) + 0.2, dim=-1).to(dev())
    U, _ = torch.linalg.qr(torch.randn(24, 3, 2, generator=g))
    U = U.to(dev())

    def run(fn):
        xs, ys = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
        out = fn(xs, ys)
        out.sum().backward()
        return out.detach().clone(), xs.grad.clone(), ys.grad.clone()

    fns = [lambda a, b: shwd.ops.spherical_sliced_w1(a, b, U), lambda a, b: shwd.ops.spherical_sliced_wp(a, b, U, 2.0),
           lambda a, b: shwd.ops.spherical_sliced_wp(a, b, U, 3.0), lambda a, b: shwd.losses.chamfer_distance(a, b)[0]]
    for fn in fns:
        first = run(fn)
        for _ in range(3):
            again = run(fn)
            assert all(torch.equal(p, q) for p, q in zip(first, again))


@pytest.mark.parametrize("length", [2560, 2900, 3000, 3072, 5600, 6100])
def test_kernels_just_below_the_48k_dynamic_smem_mark(length):
    """Dynamic shared memory just under 48 KB plus the kernels' static arrays exceeds the default limit: the launchers
    must opt in (a fresh process is used so that no earlier launch has raised the function attribute already)."""
    import subprocess
    import sys
    code = ("import sys, torch; sys.path.insert(0, %r); import shwd; "
            "g = torch.Generator().manual_seed(0); k = torch.rand(3, %d, generator=g).cuda(); "
            "s, p = shwd.ops.segmented_sort_raw(k); r, q = torch.sort(k, dim=-1, stable=True); "
            "assert torch.equal(p, q) and torch.equal(s, r); "
            "w = shwd.losses.emd1D_circle(k, torch.rand(3, %d, generator=g).cuda()); "
            "w2 = shwd.losses.binary_search_circle(k, torch.rand(3, %d, generator=g).cuda(), p=2); "
            "torch.cuda.synchronize(); assert torch.isfinite(w).all() and torch.isfinite(w2).all()"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), length, length, length))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.parametrize("case", ["constant", "duplicates", "one_vs_many", "endpoints", "identical"])
def test_emd1d_circle_degenerate_rows(shwd, case):
    """Ties and degenerate rows of the level-median kernel against emd1D_circle with stable sorts (ties: input order, a u
    entry before an equal v entry): all-equal coordinates (the cumulative gap sum never reaches 0.5 -> the smallest CDF
    difference is the median, the reference's argmin over an all-inf row), heavy duplicates, n = 1, coordinates exactly
    0 and 1, and identical clouds (W = 0)."""
    g = torch.Generator().manual_seed(len(case))
    S = 4
    if case == "constant":
        u, v = torch.full((S, 300), 0.7), torch.full((S, 200), 0.7)
    elif case == "duplicates":
        u = torch.randint(0, 12, (S, 1000), generator=g).float() / 12
        v = torch.randint(0, 12, (S, 900), generator=g).float() / 12
    elif case == "one_vs_many":
        u, v = torch.rand(S, 1, generator=g), torch.rand(S, 4096, generator=g)
    elif case == "endpoints":
        u = torch.cat([torch.zeros(S, 5), torch.ones(S, 5), torch.rand(S, 502, generator=g)], 1)
        v = torch.cat([torch.ones(S, 7), torch.zeros(S, 3), torch.rand(S, 246, generator=g)], 1)
    else:
        u = torch.rand(S, 2048, generator=g)
        v = u.clone()
    ur, vr = u.clone().requires_grad_(True), v.clone().requires_grad_(True)
    wr = oracle.emd1d_circle(ur, vr, stable=True)
    wr.sum().backward()
    ug, vg = u.to(dev()).requires_grad_(True), v.to(dev()).requires_grad_(True)
    w = shwd.losses.emd1D_circle(ug, vg)
    w.sum().backward()
    assert torch.isfinite(w).all() and torch.isfinite(ug.grad).all() and torch.isfinite(vg.grad).all()
    assert (w.cpu() - wr.detach()).abs().max().item() <= 1e-5 * max(wr.abs().max().item(), 1e-3)
    if case in ("constant", "endpoints", "one_vs_many"):  # tie-free in F, or F ties that do not move the median
        assert rel(ug.grad, ur.grad) < 1e-4 and rel(vg.grad, vr.grad) < 1e-4


@pytest.mark.parametrize("p", [2.0, 3.0])
def test_circular_wp_degenerate_rows(shwd, p):
    """binary_search_circle on identical clouds (cost 0 at rotation 0) and on a pure rotation of the same cloud."""
    g = torch.Generator().manual_seed(int(p))
    u = torch.rand(3, 1024, generator=g)
    for shift in (0.0, 0.25):
        v = (u + shift) % 1.0
        w = shwd.losses.binary_search_circle(u.to(dev()), v.to(dev()), p=p)
        wr = oracle.binary_search_circle(u, v, p=p)
        assert torch.isfinite(w).all()
        assert (w.cpu() - wr).abs().max().item() <= 1e-5 * max(wr.abs().max().item(), 1e-3) + 1e-9


def test_circular_wp_single_slice_all_negative_cdf(shwd):
    """The batch-global sign test of dCost / Cost (max_spherical_sliced_w.py:41-42, :82-83) in the one place where it is
    reachable: a call of ONE slice whose rotation makes every shifted CDF entry negative (frac rounds to 1.0 for theta in
    [-2^-25, 0), and fl(m * fl(1/m)) < 1 for this m).  The reference then leaves the CDF un-wrapped and un-rolled; a call
    of one slice reproduces that, value, rotation and gradients."""
    import numpy as np
    ms = [m for m in range(40, 400) if np.float32(m) * np.float32(1.0 / m) < np.float32(1.0)]
    assert ms, "no cloud size with a last CDF entry below 1 in the range"
    tm, tp = -2.0 ** -24, 0.0
    hit = 0
    for m in ms[:4]:
        n = m + 7
        u, v = _tie_free(1, n, 300 + m), _tie_free(1, m, 400 + m, lo=0.1, width=0.8)
        # the oracle must really be in the all-negative state at the first midpoint, otherwise the test tests nothing
        vc = torch.cumsum(torch.full((m,), 1 / m), -1)
        th0 = torch.tensor((tm + tp) / 2, dtype=torch.float32)
        hit += int(bool(((vc - (th0 - torch.floor(th0))) < 0).all()))
        ur, vr = u.clone().requires_grad_(True), v.clone().requires_grad_(True)
        wr, thr = oracle.sliced.binary_search_circle(ur, vr, p=2, tm=tm, tp=tp, return_theta=True)
        wr.sum().backward()
        ug, vg = u.clone().to(dev()).requires_grad_(True), v.clone().to(dev()).requires_grad_(True)
        us, _ = shwd.ops.SegmentedSortFn.apply(ug)
        vs, _ = shwd.ops.SegmentedSortFn.apply(vg)
        w, th = shwd.ops.CircularWpFn.apply(us, vs, 2.0, tm, tp, 1e-7)
        w.sum().backward()
        print("single-slice all-negative CDF m=%d: w %.3e theta %.3e gu %.3e gv %.3e" % (
            m, rel(w, wr), (th.cpu() - thr).abs().max().item(), rel(ug.grad, ur.grad), rel(vg.grad, vr.grad)))
        assert rel(w, wr) < TOL and (th.cpu() - thr).abs().max().item() < 2e-6
        assert rel(ug.grad, ur.grad) < 1e-4 and rel(vg.grad, vr.grad) < 1e-4
    assert hit > 0


def test_circular_wp_memoised_rounds_do_not_depend_on_the_company(shwd):
    """The bisection reuses a thread's partial sums of dCost from a bracket end when its searches return the end's
    results (csrc/circular_wp.cu, DcMemo): a slice must get the same bits whether it runs alone, first or last in a call,
    and whatever its neighbours are (the reuse decision is per CTA)."""
    for n, m in ((4096, 4096), (1500, 1333), (257, 4000)):
        u, v = _tie_free(6, n, 31 + n).to(dev()), _tie_free(6, m, 32 + m, lo=0.15, width=0.6).to(dev())
        us, vs = torch.sort(u, -1)[0].contiguous(), torch.sort(v, -1)[0].contiguous()
        w, th = shwd.ops.CircularWpFn.apply(us, vs, 2.0, -1.0, 1.0, 1e-7)
        for r in (0, 3, 5):
            w1, th1 = shwd.ops.CircularWpFn.apply(us[r:r + 1].contiguous(), vs[r:r + 1].contiguous(), 2.0, -1.0, 1.0, 1e-7)
            # a lone slice differs from a slice in company only through the all-negative-CDF rule, unreachable from [-1, 1]
            assert torch.equal(w1, w[r:r + 1]) and torch.equal(th1, th[r:r + 1])
        w2, th2 = shwd.ops.CircularWpFn.apply(us.flip(0).contiguous(), vs.flip(0).contiguous(), 2.0, -1.0, 1.0, 1e-7)
        assert torch.equal(w2.flip(0), w) and torch.equal(th2.flip(0), th)


# ------------------------------------------------------------------------------------------------- CUDA graphs ----
def _graph_cases(shwd):
    g = torch.Generator().manual_seed(5)
    U, _ = torch.linalg.qr(torch.randn(16, 3, 2, generator=g))
    U = U.to(dev())
    th = F.normalize(torch.randn(16, 3, generator=g), dim=-1).to(dev())
    geo = shwd.losses.Geodesic_distance_W(device=dev(), p=2, eps=0.02, max_iter=20)
    return {
        "chamfer": (lambda a, b: shwd.losses.chamfer_distance(a, b)[0], (4, 300, 3), (4, 260, 3)),
        "ssw1": (lambda a, b: shwd.losses.sliced_cost(a, b, U, p=1), (500, 3), (450, 3)),
        "ssw2": (lambda a, b: shwd.losses.sliced_cost(a, b, U, p=2), (500, 3), (450, 3)),
        "euclid_sw": (lambda a, b: shwd.losses.sliced_wasserstein_distance(a, b, p=2, device=dev(), projections=th), (400, 3), (400, 3)),
        "geodesic_w": (lambda a, b: geo(a, b), (3, 200, 3), (3, 160, 3)),  # the persistent cooperative sweeps
    }


@pytest.mark.parametrize("which", ["chamfer", "ssw1", "ssw2", "euclid_sw", "geodesic_w"])
def test_graphed_loss_replays_forward_and_backward_bit_exactly(shwd, which):
    """graphs.graphed_loss: forward + backward captured once, replayed on NEW inputs -> the same bits as the eager call
    (the same kernels run on the same data; only the launch mechanism differs)."""
    fn, sx, sy = _graph_cases(shwd)[which]
    gfn = shwd.graphed_loss(fn)
    g = torch.Generator().manual_seed(11)
    for trial in range(3):  # trial 0 captures, 1 and 2 replay with different data
        x = F.normalize(torch.randn(*sx, generator=g), dim=-1).to(dev())
        y = F.normalize(torch.randn(*sy, generator=g) + 0.3, dim=-1).to(dev())
        xe, ye = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
        le = fn(xe, ye)
        (2.0 * le).backward()
        xg, yg = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
        lg = gfn(xg, yg)
        (2.0 * lg).backward()
        assert torch.equal(lg.detach(), le.detach()), (which, trial)
        assert torch.equal(xg.grad, xe.grad) and torch.equal(yg.grad, ye.grad), (which, trial)
    assert len(gfn._captures) == 1
    # zero-copy form used by gradient-flow loops
    loss, (gx, gy) = gfn.value_and_grad(xg, yg)
    assert torch.equal(loss, le.detach()) and torch.equal(2.0 * gx, xe.grad) and torch.equal(2.0 * gy, ye.grad)


def test_graphed_loss_signatures_and_one_sided_gradients(shwd):
    fn = lambda a, b: shwd.losses.chamfer_distance(a, b, batch_reduction="sum")[0]  # noqa: E731
    gfn = shwd.graphed_loss(fn)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 128, 3, generator=g).to(dev())
    y = torch.randn(2, 96, 3, generator=g).to(dev())
    xg = x.clone().requires_grad_(True)
    out = gfn(xg, y)  # only x requires grad (the registration setting: y is the fixed template)
    out.backward()
    xe = x.clone().requires_grad_(True)
    fn(xe, y).backward()
    assert torch.equal(xg.grad, xe.grad)
    with torch.no_grad():
        assert torch.equal(gfn(x, y), fn(x, y))
        assert torch.equal(gfn(xg, y), fn(x, y))  # requires_grad input under no_grad: the gradient-free capture is used
    gfn(torch.randn(1, 50, 3, device=dev()), torch.randn(1, 70, 3, device=dev()))  # a new shape -> a new capture
    assert len(gfn._captures) == 3
    with pytest.raises(RuntimeError):
        gfn(x.cpu(), y.cpu())
    with pytest.raises(ValueError):
        shwd.graphed_loss(lambda a, b: shwd.chamfer_nn(a, b)[0])(x, y)  # not a scalar


# ------------------------------------------------------------------------------------- ot.emd2 on a cost matrix ----
@pytest.mark.parametrize("N,kind", [(1, "rand"), (7, "rand"), (64, "rand"), (300, "euclid"), (1024, "euclid"), (256, "ties")])
def test_dense_emd2_dropin_is_the_lp_optimum(shwd, N, kind):
    """dropin_ot/ot.emd2(a, b, M) (auction kernel on an explicit cost matrix) against scipy's exact assignment on the same
    float32 matrix: optimal value, the permutation (where it is unique) and d emd2 / dM = plan -- POT's semantics."""
    import sys
    from scipy.optimize import linear_sum_assignment
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin_ot"))
    import ot
    g = torch.Generator().manual_seed(N)
    if kind == "euclid":  # main_rotation.py:82-92 cost_matrix, p = 2
        x = F.normalize(torch.randn(N, 3, generator=g), dim=-1)
        y = F.normalize(torch.randn(N, 3, generator=g) + 0.2, dim=-1)
        M = torch.pow(torch.sum(torch.abs(x.unsqueeze(-2) - y.unsqueeze(-3)) ** 2, -1), 0.5)
    elif kind == "ties":
        M = torch.randint(0, 4, (N, N), generator=g).float()
    else:
        M = torch.rand(N, N, generator=g)
    r, c = linear_sum_assignment(M.double().numpy())
    ref = M.double()[r, c].sum().item() / N
    Mg = M.to(dev()).requires_grad_(True)
    w = torch.full((N,), 1.0 / N, device=dev())
    val = ot.emd2(w, w, Mg)
    assert val.shape == () and val.dtype == torch.float32
    assert abs(val.item() - ref) <= 2e-6 * max(abs(ref), 1e-30) + 1e-12
    val.backward()
    plan = Mg.grad.cpu()
    assert torch.equal(plan.sum(0), torch.full((N,), 1.0 / N)) and torch.equal(plan.sum(1), torch.full((N,), 1.0 / N))
    assert (plan > 0).sum().item() == N
    if kind != "ties":
        assert torch.equal(plan.argmax(1), torch.from_numpy(c).long())
    # batched form + the reference's POT_loss pattern (main_rotation.py:63-79): sum_b emd2_b^(1/p)
    Mb = torch.stack([M, M.t().contiguous()]).to(dev())
    vb = shwd.exact_emd2_dense(Mb)
    assert vb.shape == (2,) and abs(vb[0].item() - val.item()) < 1e-7 and abs(vb[1].item() - ref) <= 2e-6 * max(abs(ref), 1e-30) + 1e-12
    with pytest.raises(NotImplementedError):  # rectangular beyond the kernel: lcm(n, m) copies would not fit
        ot.emd2(torch.full((3000,), 1 / 3000, device=dev()), torch.full((7,), 1 / 7, device=dev()), torch.rand(3000, 7, device=dev()))
    if N > 1:
        w2 = w.clone()
        w2[0] *= 1.5
        with pytest.raises(NotImplementedError):
            ot.emd2(w2, w, Mg)
