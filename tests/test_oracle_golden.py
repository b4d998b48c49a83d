"""CPU: the oracle restatements reproduce the frozen outputs of the imported reference (tests/golden/*.npz,
made by tests/golden/make_golden.py) and the reference's one deterministic known answer."""
import os

import numpy as np
import pytest
import torch

import oracle

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return dict(np.load(os.path.join(G, name + ".npz"), allow_pickle=False))


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_known_answer_sinkhorn_fixed_smoke():
    # Point_Cloud_Resistration/losses/Sinkhorn_fixed.py:97-110
    d = load("sinkhorn_fixed_smoke")
    a, b = torch.from_numpy(d["a"]), torch.from_numpy(d["b"])
    l1 = oracle.log_sinkhorn(a, b, "euclid", 1, 0.1, 10, 1e-9, "mean").item()
    l2 = oracle.log_sinkhorn(a, b, "euclid", 2, 0.1, 10, 1e-9, "mean").item()
    assert l1 == pytest.approx(8.621342658996582, rel=1e-7) and l1 == pytest.approx(float(d["smoke_fixed_L1"]), rel=0, abs=0)
    assert l2 == pytest.approx(8.161666870117188, rel=1e-7) and l2 == pytest.approx(float(d["smoke_fixed_L2"]), rel=0, abs=0)


CASES = [
    # fixture, kind, p, thresh, n_power
    ("sinkhorn_cmp_L2", "sqeuclid", 2, 1e-9, 1),
    ("sinkhorn_cmp_L1_sum", "sqeuclid", 1, 1e-9, 1),
    ("sinkhorn_plain_L2_none", "sqeuclid", 2, None, 1),
    ("sinkhorn_fixed_L2", "euclid", 2, 1e-9, 1),
    ("sinkhorn_logN_2", "sqeuclid", 2, 1e-9, 2),
    ("sinkhorn_cmp_unbatched", "sqeuclid", 2, 1e-9, 1),
    ("geodesic_sinkhorn_p2", "geodesic", 2, None, 1),
    ("geodesic_sinkhorn_p1", "geodesic", 1, None, 1),
    ("geodesic_sinkhorn_p2_ragged", "geodesic", 2, None, 1),
]


@pytest.mark.parametrize("name,kind,p,thresh,n_power", CASES)
def test_oracle_sinkhorn_matches_reference(name, kind, p, thresh, n_power):
    d = load(name)
    x = torch.from_numpy(d["x"]).requires_grad_(True)
    y = torch.from_numpy(d["y"]).requires_grad_(True)
    loss = oracle.log_sinkhorn(x, y, kind, p, float(d["eps"]), int(d["max_iter"]), thresh, str(d["batch_reduction"]), n_power)
    total = loss if loss.dim() == 0 else loss.sum()
    gx, gy = torch.autograd.grad(total, (x, y))
    # same torch ops in the same order -> bit-equal on the machine that made the fixture; allow ulp-level drift
    # across CPU models / thread counts.
    assert rel(loss.detach().numpy(), d["loss"]) < 2e-6
    assert rel(gx.numpy(), d["gx"]) < 2e-5
    assert rel(gy.numpy(), d["gy"]) < 2e-5


def test_oracle_cost_matrices():
    d = load("cost_matrices")
    x, y = torch.from_numpy(d["x"]), torch.from_numpy(d["y"])
    for key, kind, p in (("geodesic_p2", "geodesic", 2), ("geodesic_p1", "geodesic", 1), ("sqeuclid_p2", "sqeuclid", 2),
                         ("sqeuclid_p1", "sqeuclid", 1)):
        C = oracle.cost_matrix(x, y, kind, p).numpy()
        assert np.array_equal(C, d[key]), key


def test_oracle_exact_emd_standin_matches_reference_wrapper():
    # informational row: Geodesic_distance_W / Cos_disimilarity_W through the scipy-backed ot shim
    d = load("cost_matrices")
    x, y = torch.from_numpy(d["x"]), torch.from_numpy(d["y"])
    for key, kind in (("exact_emd_geodesic_p2", "geodesic"), ("exact_emd_sqeuclid_p2", "sqeuclid")):
        C = oracle.cost_matrix(x, y, kind, 2)
        vals = [oracle.exact_emd2(C[b])[0] ** 0.5 for b in range(C.shape[0])]
        assert np.mean(vals) == pytest.approx(float(d[key]), rel=1e-5)


def test_oracle_regularizer():
    d = load("regularizer")
    x = torch.from_numpy(d["x"]).requires_grad_(True)
    r = oracle.flow_regularization(x)
    (g,) = torch.autograd.grad(r, x)
    assert r.item() == pytest.approx(float(d["reg"]), rel=1e-6)
    assert rel(g.numpy(), d["gx"]) < 1e-6


@pytest.mark.parametrize("p", [1, 2])
def test_oracle_spherical_sliced(p):
    d = load(f"ssw_p{p}")
    xs = torch.from_numpy(d["Xs"]).requires_grad_(True)
    xt = torch.from_numpy(d["Xt"]).requires_grad_(True)
    U = torch.from_numpy(d["U"])
    loss = oracle.sliced_wasserstein_sphere(xs, xt, U, p=p)
    gx, gy = torch.autograd.grad(loss, (xs, xt))
    assert loss.item() == pytest.approx(float(d["loss"]), rel=2e-6)
    assert rel(gx.numpy(), d["gx"]) < 1e-5
    assert rel(gy.numpy(), d["gy"]) < 1e-5


def test_oracle_emd1d_circle_and_binary_search():
    d = load("emd1d_circle")
    u = torch.from_numpy(d["u"]).requires_grad_(True)
    v = torch.from_numpy(d["v"]).requires_grad_(True)
    w = oracle.emd1d_circle(u, v)
    gu, gv = torch.autograd.grad(w.sum(), (u, v))
    assert np.array_equal(w.detach().numpy(), d["w"])
    assert np.array_equal(gu.numpy(), d["gu"]) and np.array_equal(gv.numpy(), d["gv"])
    d2 = load("binary_search_circle_p2")
    w2 = oracle.binary_search_circle(torch.from_numpy(d2["u"]), torch.from_numpy(d2["v"]), p=2)
    assert rel(w2.numpy(), d2["w"]) < 1e-6


def test_sphere_map_matches_cosine_similarity_normalisation():
    # normalise-then-dot is what F.cosine_similarity computes (s2_wasserstein.py:122)
    torch.manual_seed(0)
    x = torch.randn(4, 64, 3) * 3
    y = torch.randn(4, 48, 3)
    xh = oracle.sphere_map(x, center=False)
    yh = oracle.sphere_map(y, center=False)
    dot = (xh.unsqueeze(-2) * yh.unsqueeze(-3)).sum(-1)
    cs = torch.nn.functional.cosine_similarity(x.unsqueeze(-2), y.unsqueeze(-3), dim=-1)
    assert torch.equal(dot, cs)


def test_chamfer_oracle_manual_gradient():
    # manual scatter backward equals autograd of the dense restatement
    torch.manual_seed(1)
    x = torch.randn(2, 40, 3, requires_grad=True)
    y = torch.randn(2, 33, 3, requires_grad=True)
    loss, none = oracle.chamfer_distance(x, y)
    assert none is None
    gx, gy = torch.autograd.grad(loss, (x, y))
    with torch.no_grad():
        d = ((x.unsqueeze(2) - y.unsqueeze(1)) ** 2).sum(-1)
        jx = d.argmin(2)
        iy = d.argmin(1)
        B, N, M = 2, 40, 33
        mgx = torch.zeros_like(x)
        mgy = torch.zeros_like(y)
        for b in range(B):
            dx = 2 * (x[b] - y[b][jx[b]]) / N / B
            mgx[b] += dx
            mgy[b].index_add_(0, jx[b], -dx)
            dy = 2 * (y[b] - x[b][iy[b]]) / M / B
            mgy[b] += dy
            mgx[b].index_add_(0, iy[b], -dy)
    assert torch.allclose(gx, mgx, atol=1e-7) and torch.allclose(gy, mgy, atol=1e-7)


def test_oracle_rigid_transform_matches_reference():
    """Data side (8f #3): Dataset_Transformation / qrot / euler_to_quaternion (data_utils/Data_set_maker.py:40-230)."""
    d = load("rigid_transform")
    out, rot, trans = oracle.data.rigid_transform(torch.from_numpy(d["src"]), torch.from_numpy(d["poses"]))
    assert np.allclose(out.numpy(), d["out"], rtol=0, atol=2e-7)
    assert np.allclose(rot.numpy(), d["igt_rotation"], rtol=0, atol=2e-7)
    assert np.array_equal(trans.numpy(), d["igt_translation"])
    assert np.allclose(oracle.data.euler_to_quaternion(d["euler_in"], "xyz"), d["euler_quat"], rtol=0, atol=1e-15)
    # the pose generator consumes the numpy stream exactly like the reference (seed 1234 produced the fixture's poses)
    rng = np.random.RandomState(1234)
    poses = torch.cat([oracle.data.create_random_transform(rng) for _ in range(d["poses"].shape[0])], 0)
    assert np.array_equal(poses.numpy(), d["poses"])


# ---- rows a11 / a14 / f1: notebook sliced W, the learned sphere map, the max-over-phi wrapper ---------------------------
def _state(d, prefix):
    return {k[len(prefix):].replace("__", "."): torch.from_numpy(v) for k, v in d.items() if k.startswith(prefix)}


def test_oracle_euclid_sliced_w_matches_notebook_cell():
    """Flow_ellipsoid.ipynb cell 5 `sliced_wasserstein_distance`, executed from the notebook's own source by make_golden.py."""
    d = load("notebook_sliced_wasserstein")
    for p in (1, 2, 3):
        x = torch.from_numpy(d["x_p%d" % p]).requires_grad_(True)
        y = torch.from_numpy(d["y_p%d" % p]).requires_grad_(True)
        loss = oracle.euclid_sliced_wasserstein(x, y, torch.from_numpy(d["theta_p%d" % p]), p)
        gx, gy = torch.autograd.grad(loss, (x, y))
        assert loss.item() == pytest.approx(float(d["loss_p%d" % p]), rel=2e-6)
        assert rel(gx.numpy(), d["gx_p%d" % p]) < 2e-5 and rel(gy.numpy(), d["gy_p%d" % p]) < 2e-5


@pytest.mark.parametrize("name", ["Residual", "Planar"])
def test_flow_modules_replay_the_vendored_normflows(name):
    """losses/flows.py (the eager definition of phi the fused kernel is tested against) loaded with the reference's own
    state_dict -- same keys -- reproduces the vendored normflows' outputs and gradients (batched and un-batched inputs)."""
    import shwd
    d = load("flow_" + name.lower())
    phi = shwd.losses.Norm_Flow_structure(flow_name=name, n_flow_layer=3)
    missing = phi.load_state_dict(_state(d, "sd__"), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    phi.train()
    for tag in ("b", "u"):
        x = torch.from_numpy(d["x_" + tag]).requires_grad_(True)
        y = phi(x)
        named = [(n, q) for n, q in phi.named_parameters() if q.dtype == torch.float32 and q.dim() > 0]
        gs = torch.autograd.grad((y * torch.from_numpy(d["w_" + tag])).sum(), [x] + [q for _, q in named], allow_unused=True)
        assert rel(y.detach().numpy(), d["y_" + tag]) < 1e-6
        assert rel(gs[0].numpy(), d["gx_" + tag]) < 1e-5
        for (n, q), g in zip(named, gs[1:]):
            want = d["gp_%s__%s" % (tag, n.replace(".", "__"))]
            got = np.zeros_like(want) if g is None else g.numpy()
            assert np.linalg.norm(got - want) <= 2e-5 * max(np.linalg.norm(want), 1e-3), n


class _CpuExactCSW(torch.nn.Module):
    """Cos_disimilarity_W (s2_wasserstein.py:25-50) on the CPU from oracle pieces: cost matrix + exact plan per pair."""

    def forward(self, x, y):
        C = oracle.cost_matrix(x, y, "sqeuclid", 2)
        tot = 0
        for b in range(C.shape[0]):
            _, plan = oracle.exact_emd2(C[b])
            tot = tot + torch.pow((plan.to(C.dtype) * C[b]).sum(), 1.0 / 2)
        return tot / C.shape[0]


@pytest.mark.parametrize("name", ["Residual", "Planar"])
def test_max_wrapper_step_matches_reference_on_cpu(name):
    """The host orchestration of max_cos_disimilarity_wassersten_distance (s2_wasserstein.py:234-262: two SGD ascent steps
    on phi, then the outer distance) against one frozen step of the unmodified reference wrapper."""
    import shwd
    d = load("max_wrapper_" + name.lower())
    L = shwd.losses
    phi = L.Norm_Flow_structure(flow_name=name, n_flow_layer=int(d["n_flow_layer"]))
    phi.load_state_dict(_state(d, "sd0__"))
    op = torch.optim.SGD(list(phi.parameters()), lr=float(d["lr"]))
    crit = L.max_cos_disimilarity_wassersten_distance(phi=phi, CSW=_CpuExactCSW(), device="cpu", phi_op=op,
                                                       max_iter=int(d["max_iter"]), lam=float(d["lam"]))
    second = torch.from_numpy(d["second"]).requires_grad_(True)
    cswd, ft, st = crit(torch.from_numpy(d["first"]), second, "train")
    (g2,) = torch.autograd.grad(cswd, second)
    assert cswd.item() == pytest.approx(float(d["cswd"]), rel=2e-6)
    assert rel(ft.detach().numpy(), d["first_t"]) < 1e-5 and rel(st.detach().numpy(), d["second_t"]) < 1e-5
    assert rel(g2.numpy(), d["g_second"]) < 5e-5
    for k, v in _state(d, "sd1__").items():  # phi after the two ascent steps
        got = phi.state_dict()[k]
        if v.dtype == torch.float32 and v.dim() > 0 and "last_" not in k and not k.endswith("scale"):
            assert torch.allclose(got, v, rtol=2e-4, atol=2e-6), (k, (got - v).abs().max().item())


def test_oracle_weighted_emd1d_circle_matches_reference():
    d = load("emd1d_circle_weighted")
    u = torch.from_numpy(d["u"]).requires_grad_(True)
    v = torch.from_numpy(d["v"]).requires_grad_(True)
    uw = torch.from_numpy(d["uw"]).requires_grad_(True)
    vw = torch.from_numpy(d["vw"]).requires_grad_(True)
    w = oracle.emd1d_circle(u, v, u_weights=uw, v_weights=vw)
    gu, gv, guw, gvw = torch.autograd.grad(w.sum(), (u, v, uw, vw))
    assert rel(w.detach().numpy(), d["w"]) < 2e-6
    for got, want in ((gu, "gu"), (gv, "gv"), (guw, "guw"), (gvw, "gvw")):
        assert rel(got.numpy(), d[want]) < 2e-5


def test_oracle_weighted_binary_search_circle_matches_reference():
    """binary_search_circle / sliced_cost with non-uniform weights (max_spherical_sliced_w.py:117-207, 251-286): the oracle
    against the unmodified reference's frozen value and gradients (coordinates and weights)."""
    d = load("binary_search_circle_weighted")
    for p in (2, 3):
        t = {k: torch.from_numpy(d[f"{k}_p{p}"]).requires_grad_(True) for k in ("u", "v", "uw", "vw")}
        w = oracle.sliced.binary_search_circle(t["u"], t["v"], p=p, u_weights=t["uw"], v_weights=t["vw"])
        grads = torch.autograd.grad(w.sum(), (t["u"], t["v"], t["uw"], t["vw"]))
        assert rel(w.detach().numpy(), d[f"w_p{p}"]) < 2e-6
        for got, name in zip(grads, ("gu", "gv", "guw", "gvw")):
            assert rel(got.numpy(), d[f"{name}_p{p}"]) < 2e-5, (p, name)
    Xs = torch.from_numpy(d["Xs"]).requires_grad_(True)
    Xt = torch.from_numpy(d["Xt"]).requires_grad_(True)
    loss = oracle.sliced.sliced_wasserstein_sphere(Xs, Xt, torch.from_numpy(d["U"]), p=2, u_weights=torch.from_numpy(d["sc_uw"]),
                                                   v_weights=torch.from_numpy(d["sc_vw"]))
    gx, gy = torch.autograd.grad(loss, (Xs, Xt))
    assert abs(loss.item() - float(d["sc_loss"])) / float(d["sc_loss"]) < 2e-6
    assert rel(gx.numpy(), d["sc_gx"]) < 2e-5 and rel(gy.numpy(), d["sc_gy"]) < 2e-5


@pytest.mark.parametrize("p", [2, 1])
def test_oracle_sliced_path_reproduces_the_max_ssw_wrapper_fixture(p):
    """The outer value of the frozen max_spherical_wassersten_distance call (max_spherical_sliced_w.py:528-533) is a sum of
    per-pair sliced costs of the transformed clouds (40 against 33 points) with the fixture's frame cycle: calls 0-5 are the two
    ascent steps over three pairs, calls 6-8 the outer evaluation.  The oracle's sliced path must reproduce it."""
    d = load("max_ssw_wrapper")
    ft, st, Us = (torch.from_numpy(d[k % p]) for k in ("first_t_p%d", "second_t_p%d", "Us_p%d"))
    total = sum(oracle.sliced_wasserstein_sphere(ft[i], st[i], Us[(6 + i) % len(Us)], p=p) for i in range(len(ft)))
    assert float(total.detach()) == pytest.approx(float(d["ssw_p%d" % p]), rel=2e-6)


def test_oracle_sliced_path_reproduces_the_batched_fast_fixture():
    """max_spherical_sliced_w_fast.py:258-295: per-pair frames, value = sum over the pairs, p = 2 and 3, with gradients."""
    d = load("ssw_fast")
    for p in (2, 3):
        x = torch.from_numpy(d["x_p%d" % p]).requires_grad_(True)
        y = torch.from_numpy(d["y_p%d" % p]).requires_grad_(True)
        Us = torch.from_numpy(d["Us_p%d" % p])
        total = sum(oracle.sliced_wasserstein_sphere(x[i], y[i], Us[i], p=p) for i in range(len(x)))
        gx, gy = torch.autograd.grad(total, (x, y))
        assert total.item() == pytest.approx(float(d["w_p%d" % p][0]), rel=2e-6)
        assert rel(gx.numpy(), d["gx_p%d" % p]) < 1e-5 and rel(gy.numpy(), d["gy_p%d" % p]) < 1e-5
