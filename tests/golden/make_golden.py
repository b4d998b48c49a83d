"""Generate the golden fixtures in this directory by running the UNMODIFIED reference modules from
``/root/reference`` on seeded inputs (build container only: the reference tree does not exist on the GPU box).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Import shims, created in a temp dir, never inside the reference tree:
  * ``ot``                 -> ``emd2(a, b, M)`` backed by scipy's linear_sum_assignment (POT is not installed);
                              only used for the *informational* exact-EMD fixture.
  * ``normflows``          -> alias of the vendored ``losses/normflows_ishikawa`` package.
  * ``matplotlib.pyplot``  -> empty stub (imported, never used, by max_spherical_sliced_w.py:4).
The Sinkhorn files are loaded stand-alone by path.  Outputs: ``*.npz`` with inputs, outputs and autograd gradients.

Subsets (each regenerates its files bit for bit; the full run writes everything):
    --wrappers-only      flow_{residual,planar}, max_wrapper_{residual,planar}, notebook_sliced_wasserstein, emd1d_circle_weighted
    --weighted-wp-only   binary_search_circle_weighted
    --max-ssw-only       max_ssw_wrapper, ssw_fast        (max_spherical_sliced_w.py:498-536, max_spherical_sliced_w_fast.py)
    --mini-batch-only    mini_batch_mssw                  (mini_batch_Residual_MSSW.py)
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _install_shims():
    shim = tempfile.mkdtemp(prefix="shwd_shim_")
    os.symlink(os.path.join(REF, "Point_Cloud_Resistration/losses/normflows_ishikawa"), os.path.join(shim, "normflows"))
    os.makedirs(os.path.join(shim, "matplotlib"))
    open(os.path.join(shim, "matplotlib/__init__.py"), "w").close()
    open(os.path.join(shim, "matplotlib/pyplot.py"), "w").close()
    sys.path.insert(0, shim)
    sys.path.insert(0, os.path.join(REF, "Point_Cloud_Resistration"))

    ot = types.ModuleType("ot")

    class _Emd2(torch.autograd.Function):
        @staticmethod
        def forward(ctx, a, b, M):
            from scipy.optimize import linear_sum_assignment
            Mn = M.detach().double().numpy()
            r, c = linear_sum_assignment(Mn)
            plan = np.zeros_like(Mn)
            plan[r, c] = 1.0 / Mn.shape[0]
            ctx.save_for_backward(torch.from_numpy(plan).to(M.dtype))
            return torch.tensor((plan * Mn).sum(), dtype=M.dtype)

        @staticmethod
        def backward(ctx, g):
            (plan,) = ctx.saved_tensors
            return None, None, g * plan

    ot.emd2 = lambda a, b, M: _Emd2.apply(a, b, M)
    sys.modules["ot"] = ot


def sphere_pair(B, N, M, seed):
    g = torch.Generator().manual_seed(seed)
    x = F.normalize(torch.randn(B, N, 3, generator=g), dim=-1)
    # a rotated + noisy copy when N == M, an independent cloud otherwise
    if N == M:
        ang = 0.4
        R = torch.tensor([[1, 0, 0], [0, np.cos(ang), -np.sin(ang)], [0, np.sin(ang), np.cos(ang)]], dtype=torch.float32)
        y = x @ R.T + 0.05 * torch.randn(B, M, 3, generator=g)
    else:
        y = torch.randn(B, M, 3, generator=g)
    return x.contiguous(), y.contiguous()


def grads(loss, *leaves):
    gs = torch.autograd.grad(loss, leaves)
    return [g.numpy() for g in gs]


def main():
    _install_shims()
    torch.set_num_threads(8)
    sk_cmp = _load(os.path.join(REF, "Comparison_Wasserstein_with_Chamfer_distance/losses/sinkhorn.py"), "ref_sinkhorn_cmp")
    sk_fix = _load(os.path.join(REF, "Point_Cloud_Resistration/losses/Sinkhorn_fixed.py"), "ref_sinkhorn_fixed")
    sk_plain = _load(os.path.join(REF, "Point_Cloud_Resistration/losses/Sinkhorn.py"), "ref_sinkhorn_plain")
    ssw = _load(os.path.join(REF, "Point_Cloud_Resistration/losses/max_spherical_sliced_w.py"), "ref_ssw")
    import losses as ref_losses  # Point_Cloud_Resistration/losses/__init__.py:27-32

    out = {}

    # ---- 1. the one deterministic smoke of the reference: Sinkhorn_fixed.py:97-110 ------------------------------
    a = torch.tensor([[[i, 0, 0] for i in range(9)]] * 2)
    b = torch.tensor([[[i, 8, 0] for i in range(12)]] * 2)
    for norm in ("L1", "L2"):
        crit = sk_fix.log_Sinkhorn_Distance_Loss(eps=0.1, max_iter=10, batch_reduction="mean", type_of_cost_norm=norm)
        loss, P, C = crit(a, b, device="cpu")
        out[f"smoke_fixed_{norm}"] = np.float64(loss.item())
    np.savez(os.path.join(HERE, "sinkhorn_fixed_smoke.npz"), a=a.numpy(), b=b.numpy(), **out)
    print("smoke:", out)

    # ---- 2. log-Sinkhorn classes on random clouds, loss + autograd gradients ------------------------------------
    def run_sink(tag, cls, kwargs, x, y, cost_override=None, extra=None):
        x = x.clone().requires_grad_(True)
        y = y.clone().requires_grad_(True)
        crit = cls(**kwargs)
        if cost_override is not None:
            crit._cost_matrix = cost_override
        loss, P, C = crit(x, y, "cpu")
        total = loss if loss.dim() == 0 else loss.sum()
        gx, gy = grads(total, x, y)
        d = dict(x=x.detach().numpy(), y=y.detach().numpy(), loss=loss.detach().numpy(), gx=gx, gy=gy,
                 P=P.detach().numpy() if P.numel() <= 70000 else np.zeros(0, np.float32),
                 C=C.detach().numpy() if C.numel() <= 70000 else np.zeros(0, np.float32))
        d.update({k: np.asarray(v) for k, v in kwargs.items() if not isinstance(v, str)})
        d.update({k: np.asarray(v) for k, v in kwargs.items() if isinstance(v, str)})
        if extra:
            d.update(extra)
        np.savez(os.path.join(HERE, tag + ".npz"), **d)
        print(tag, "loss", loss.detach().numpy().ravel()[:4], "|gx|", np.linalg.norm(gx), "|gy|", np.linalg.norm(gy))

    x, y = sphere_pair(2, 96, 80, 11)
    run_sink("sinkhorn_cmp_L2", sk_cmp.log_Sinkhorn_Distance_Loss,
             dict(eps=0.01, max_iter=100, batch_reduction="mean", type_of_cost_norm="L2"), x, y)
    run_sink("sinkhorn_cmp_L1_sum", sk_cmp.log_Sinkhorn_Distance_Loss,
             dict(eps=0.05, max_iter=40, batch_reduction="sum", type_of_cost_norm="L1"), x, y)
    run_sink("sinkhorn_plain_L2_none", sk_plain.Sinkhorn_Distance_Loss,
             dict(eps=0.02, max_iter=60, batch_reduction="none", type_of_cost_norm="L2"), x, y)
    run_sink("sinkhorn_fixed_L2", sk_fix.log_Sinkhorn_Distance_Loss,
             dict(eps=0.05, max_iter=50, batch_reduction="mean", type_of_cost_norm="L2"), x, y)
    run_sink("sinkhorn_logN_2", sk_cmp.log_N_Sinkhorn_Distance_Loss,
             dict(eps=0.05, max_iter=50, batch_reduction="mean", type_of_cost_norm="L2", type_of_Wasserstein_N="2"), x, y)
    xs, ys = sphere_pair(1, 64, 64, 5)
    run_sink("sinkhorn_cmp_unbatched", sk_cmp.log_Sinkhorn_Distance_Loss,
             dict(eps=0.01, max_iter=30, batch_reduction="mean", type_of_cost_norm="L2"), xs[0], ys[0])

    # ---- 3. the north-star composition: reference Sinkhorn recurrence on the reference geodesic cost ------------
    geo = ref_losses.Geodesic_distance_W(device="cpu", p=2)
    geo1 = ref_losses.Geodesic_distance_W(device="cpu", p=1)
    x, y = sphere_pair(2, 128, 128, 1234)
    x = x * (1.0 + 0.3 * torch.rand(2, 128, 1, generator=torch.Generator().manual_seed(3)))  # not unit-norm on entry
    run_sink("geodesic_sinkhorn_p2", sk_plain.Sinkhorn_Distance_Loss,
             dict(eps=0.01, max_iter=100, batch_reduction="mean", type_of_cost_norm="L2"), x, y,
             cost_override=lambda x, y, p: geo.geodesic_cost_matrix(x, y, 2))
    run_sink("geodesic_sinkhorn_p1", sk_plain.Sinkhorn_Distance_Loss,
             dict(eps=0.05, max_iter=50, batch_reduction="none", type_of_cost_norm="L1"), x, y,
             cost_override=lambda x, y, p: geo1.geodesic_cost_matrix(x, y, 1))
    x3, y3 = sphere_pair(3, 200, 136, 77)
    run_sink("geodesic_sinkhorn_p2_ragged", sk_plain.Sinkhorn_Distance_Loss,
             dict(eps=0.02, max_iter=40, batch_reduction="sum", type_of_cost_norm="L2"), x3, y3,
             cost_override=lambda x, y, p: geo.geodesic_cost_matrix(x, y, 2))

    # ---- 4. cost matrices + exact-EMD wrappers (informational: needs the scipy-backed ``ot`` shim) ---------------
    cosw = ref_losses.Cos_disimilarity_W(device="cpu", p=2)
    x, y = sphere_pair(2, 64, 64, 21)
    Cg = geo.geodesic_cost_matrix(x, y, 2)
    Cg1 = geo1.geodesic_cost_matrix(x, y, 1)
    Cs = cosw.cos_cost_matrix(x, y, 2)
    np.savez(os.path.join(HERE, "cost_matrices.npz"), x=x.numpy(), y=y.numpy(), geodesic_p2=Cg.numpy(),
             geodesic_p1=Cg1.numpy(), sqeuclid_p2=Cs.numpy(), sqeuclid_p1=cosw.cos_cost_matrix(x, y, 1).numpy(),
             exact_emd_geodesic_p2=np.float64(geo(x, y).item()), exact_emd_sqeuclid_p2=np.float64(cosw(x, y).item()))

    # ---- 5. regulariser (s2_wasserstein.py:224-232) --------------------------------------------------------------
    crit = ref_losses.max_cos_disimilarity_wassersten_distance(phi=None, CSW=None, device="cpu", phi_op=None)
    xr = torch.randn(3, 50, 3, generator=torch.Generator().manual_seed(9)).requires_grad_(True)
    reg = crit.regularization_of_normalizing_flow(xr)
    (gr,) = grads(reg, xr)
    np.savez(os.path.join(HERE, "regularizer.npz"), x=xr.detach().numpy(), reg=reg.detach().numpy(), gx=gr)

    # ---- 6. spherical sliced W (explicit frames U) --------------------------------------------------------------
    g = torch.Generator().manual_seed(42)
    Xs = F.normalize(torch.randn(150, 3, generator=g), dim=-1)
    Xt = F.normalize(torch.randn(131, 3, generator=g) + torch.tensor([0.5, 0.0, 0.0]), dim=-1)
    U, _ = torch.linalg.qr(torch.randn(24, 3, 2, generator=g))
    for p in (1, 2):
        xs = Xs.clone().requires_grad_(True)
        xt = Xt.clone().requires_grad_(True)
        loss = ssw.sliced_cost(xs, xt, U, p=p)
        gx, gy = grads(loss, xs, xt)
        np.savez(os.path.join(HERE, f"ssw_p{p}.npz"), Xs=Xs.numpy(), Xt=Xt.numpy(), U=U.numpy(), loss=loss.detach().numpy(),
                 gx=gx, gy=gy)
        print(f"ssw p={p}", loss.item())
    # raw circular W1 on given circle coordinates
    u = torch.rand(9, 70, generator=g).requires_grad_(True)
    v = torch.rand(9, 55, generator=g).requires_grad_(True)
    w = ssw.emd1D_circle(u, v, p=1)
    gu, gv = grads(w.sum(), u, v)
    np.savez(os.path.join(HERE, "emd1d_circle.npz"), u=u.detach().numpy(), v=v.detach().numpy(), w=w.detach().numpy(), gu=gu, gv=gv)
    w2 = ssw.binary_search_circle(u.detach(), v.detach(), p=2)
    np.savez(os.path.join(HERE, "binary_search_circle_p2.npz"), u=u.detach().numpy(), v=v.detach().numpy(), w=w2.numpy())

    make_wrapper_fixtures(ref_losses)
    make_notebook_fixture()
    make_weighted_circle_fixture()
    make_weighted_wp_fixture()
    make_max_ssw_fixture()
    make_ssw_fast_fixture()
    make_mini_batch_mssw_fixture()


def _flat_state(prefix, sd):
    return {prefix + k.replace(".", "__"): v.detach().cpu().numpy() for k, v in sd.items()}


def make_wrapper_fixtures(ref_losses):
    """Rows a14 / f1: the learned sphere map and the max-over-phi wrapper, from the UNMODIFIED
    reference modules (vendored normflows 1.7.2 + s2_wasserstein.py:134-163, 211-262), with every weight saved so the
    fixtures replay through ``load_state_dict`` -- the state_dict keys are the reference's own."""
    # ---- 7. Norm_Flow_structure forward + gradients (Residual x3 as train_W_COS.py:390 builds it; Planar x3) -----
    for name in ("Residual", "Planar"):
        torch.manual_seed(101 if name == "Residual" else 202)
        np.random.seed(5)
        phi = ref_losses.Norm_Flow_structure(flow_name=name, n_flow_layer=3)
        with torch.no_grad():  # move the weights off their init so every term of the chain rule is exercised
            for prm in phi.parameters():
                if prm.dim() > 0 and prm.dtype == torch.float32:
                    prm.add_(0.3 * torch.randn_like(prm))
        phi.train()
        sd0 = {k: v.clone() for k, v in phi.state_dict().items()}
        g = torch.Generator().manual_seed(7)
        out = {}
        for tag, shape in (("b", (3, 64, 3)), ("u", (50, 3))):  # batched clouds and the un-batched (N,3) branch
            x = (torch.randn(*shape, generator=g) * 0.8).requires_grad_(True)
            w = torch.randn(*shape, generator=g)
            y = phi(x)
            params = [q for q in phi.parameters() if q.dtype == torch.float32 and q.dim() > 0]
            names = [n for n, q in phi.named_parameters() if q.dtype == torch.float32 and q.dim() > 0]
            gs = torch.autograd.grad((y * w).sum(), [x] + params, allow_unused=True)
            out.update({f"x_{tag}": x.detach().numpy(), f"w_{tag}": w.numpy(), f"y_{tag}": y.detach().numpy(),
                        f"gx_{tag}": gs[0].numpy()})
            for n, gq, q in zip(names, gs[1:], params):
                out[f"gp_{tag}__" + n.replace(".", "__")] = (torch.zeros_like(q) if gq is None else gq).numpy()
        out.update(_flat_state("sd__", sd0))
        np.savez(os.path.join(HERE, f"flow_{name.lower()}.npz"), **out)
        print("flow", name, "|y|", float(np.linalg.norm(out["y_b"])), "|gx|", float(np.linalg.norm(out["gx_b"])))

    # ---- 8. one training step of max_cos_disimilarity_wassersten_distance (s2_wasserstein.py:234-262) -------------
    # phi ascent (max_iter=2, SGD so the update is a plain function of the gradient) on detached inputs, then the outer
    # distance; CSW = Cos_disimilarity_W(p=2) as train_W_COS.py:393 (exact EMD through the scipy-backed ``ot`` shim)
    for name, layers in (("Residual", 2), ("Planar", 3)):
        torch.manual_seed(303 if name == "Residual" else 404)
        np.random.seed(6)
        phi = ref_losses.Norm_Flow_structure(flow_name=name, n_flow_layer=layers)
        phi_op = torch.optim.SGD([q for q in phi.parameters()], lr=0.05)
        csw = ref_losses.Cos_disimilarity_W(device="cpu", p=2)
        crit = ref_losses.max_cos_disimilarity_wassersten_distance(phi=phi, CSW=csw, device="cpu", phi_op=phi_op, max_iter=2, lam=0.1)
        sd0 = {k: v.clone() for k, v in phi.state_dict().items()}
        g = torch.Generator().manual_seed(8)
        first = torch.randn(2, 48, 3, generator=g)
        first = first - first.mean(1, keepdim=True)
        second = (first[:, torch.randperm(48, generator=g)] * 0.9 + 0.05 * torch.randn(2, 48, 3, generator=g)).requires_grad_(True)
        cswd, first_t, second_t = crit(first, second, "train")
        (g_second,) = torch.autograd.grad(cswd, second)
        sd1 = phi.state_dict()
        cswd_test, ft_test, st_test = crit(first, second.detach(), "test")
        out = dict(first=first.detach().numpy(), second=second.detach().numpy(), cswd=np.float64(cswd.item()), first_t=first_t.detach().numpy(),
                   second_t=second_t.detach().numpy(), g_second=g_second.numpy(), cswd_test=np.float64(cswd_test.item()),
                   lr=np.float64(0.05), lam=np.float64(0.1), max_iter=np.int64(2), n_flow_layer=np.int64(layers))
        out.update(_flat_state("sd0__", sd0))
        out.update(_flat_state("sd1__", sd1))
        np.savez(os.path.join(HERE, f"max_wrapper_{name.lower()}.npz"), **out)
        print("max wrapper", name, "cswd", cswd.item(), "test", cswd_test.item())


def make_weighted_circle_fixture():
    """emd1D_circle with non-uniform weights (max_spherical_sliced_w.py:210-247, weights gathered through the sorts :224-228)."""
    ssw = _load(os.path.join(REF, "Point_Cloud_Resistration/losses/max_spherical_sliced_w.py"), "ref_ssw_w")
    g = torch.Generator().manual_seed(77)
    u = torch.rand(6, 90, generator=g).requires_grad_(True)
    v = torch.rand(6, 64, generator=g).requires_grad_(True)
    uw = torch.rand(90, generator=g) + 0.1
    uw = (uw / uw.sum()).requires_grad_(True)
    vw = torch.rand(64, generator=g) + 0.1
    vw = (vw / vw.sum()).requires_grad_(True)
    w = ssw.emd1D_circle(u, v, u_weights=uw, v_weights=vw, p=1)
    gu, gv, guw, gvw = torch.autograd.grad(w.sum(), (u, v, uw, vw))
    np.savez(os.path.join(HERE, "emd1d_circle_weighted.npz"), u=u.detach().numpy(), v=v.detach().numpy(), uw=uw.detach().numpy(),
             vw=vw.detach().numpy(), w=w.detach().numpy(), gu=gu.numpy(), gv=gv.numpy(), guw=guw.numpy(), gvw=gvw.numpy())
    print("weighted emd1D_circle", w.detach().numpy()[:3])


def make_weighted_wp_fixture():
    """binary_search_circle with non-uniform weights (max_spherical_sliced_w.py:117-207; weights gathered through the sorts and
    accumulated, :156-170), p = 2 and p = 3: value, gradients w.r.t. the coordinates and w.r.t. the weights; and one
    sliced_cost(p=2, u_weights, v_weights) call (:251-286)."""
    ssw = _load(os.path.join(REF, "Point_Cloud_Resistration/losses/max_spherical_sliced_w.py"), "ref_ssw_wp")
    g = torch.Generator().manual_seed(78)
    out = {}
    for p in (2, 3):
        u = torch.rand(5, 83, generator=g).requires_grad_(True)
        v = torch.rand(5, 61, generator=g).requires_grad_(True)
        uw = torch.rand(83, generator=g) + 0.1
        uw = (uw / uw.sum()).requires_grad_(True)
        vw = torch.rand(61, generator=g) + 0.1
        vw = (vw / vw.sum()).requires_grad_(True)
        w = ssw.binary_search_circle(u, v, u_weights=uw, v_weights=vw, p=p)
        gu, gv, guw, gvw = torch.autograd.grad(w.sum(), (u, v, uw, vw))
        out.update({f"u_p{p}": u.detach().numpy(), f"v_p{p}": v.detach().numpy(), f"uw_p{p}": uw.detach().numpy(),
                    f"vw_p{p}": vw.detach().numpy(), f"w_p{p}": w.detach().numpy(), f"gu_p{p}": gu.numpy(), f"gv_p{p}": gv.numpy(),
                    f"guw_p{p}": guw.numpy(), f"gvw_p{p}": gvw.numpy()})
        print("weighted binary_search_circle p=%d" % p, w.detach().numpy()[:3])
    Xs = F.normalize(torch.randn(140, 3, generator=g), dim=-1).requires_grad_(True)
    Xt = F.normalize(torch.randn(110, 3, generator=g) + 0.4, dim=-1).requires_grad_(True)
    U, _ = torch.linalg.qr(torch.randn(11, 3, 2, generator=g))
    uw = torch.rand(140, generator=g) + 0.2
    uw = uw / uw.sum()
    vw = torch.rand(110, generator=g) + 0.2
    vw = vw / vw.sum()
    loss = ssw.sliced_cost(Xs, Xt, U, p=2, u_weights=uw, v_weights=vw)
    gx, gy = torch.autograd.grad(loss, (Xs, Xt))
    out.update(dict(Xs=Xs.detach().numpy(), Xt=Xt.detach().numpy(), U=U.numpy(), sc_uw=uw.numpy(), sc_vw=vw.numpy(),
                    sc_loss=np.float64(loss.item()), sc_gx=gx.numpy(), sc_gy=gy.numpy()))
    print("weighted sliced_cost p=2", loss.item())
    np.savez(os.path.join(HERE, "binary_search_circle_weighted.npz"), **out)


def make_notebook_fixture():
    """Row a11: the Euclidean sliced Wasserstein distance of the flow notebooks, by executing the SOURCE of cell 5 of
    Wasserstein_flow_problem/Flow_ellipsoid.ipynb (``rand_projections`` / ``sliced_wasserstein_distance``, raw JSON lines
    203-220) unmodified; only the random directions are pinned by seeding torch right before the call."""
    import json
    with open(os.path.join(REF, "Wasserstein_flow_problem/Flow_ellipsoid.ipynb")) as fh:
        nb = json.load(fh)
    src = "".join(nb["cells"][5]["source"])
    assert "def sliced_wasserstein_distance" in src
    ns = {"torch": torch, "np": np, "optim": torch.optim, "nn": torch.nn}
    exec(compile(src, "Flow_ellipsoid.ipynb#cell5", "exec"), ns)
    g = torch.Generator().manual_seed(55)
    out = {}
    for p in (1, 2, 3):
        P = 40
        ns["num_projections"] = P  # the cell reads this global inside rand_projections(dim, num_projections) (:214)
        x = (torch.randn(300, 3, generator=g) * torch.tensor([2.0, 1.0, 1.0])).requires_grad_(True)
        y = (torch.randn(300, 3, generator=g) + 0.3).requires_grad_(True)
        torch.manual_seed(1000 + p)
        theta = ns["rand_projections"](3, P)        # the directions the call below will draw (same seed)
        torch.manual_seed(1000 + p)
        loss = ns["sliced_wasserstein_distance"](x, y, num_projection=P, p=p, device="cpu")
        gx, gy = torch.autograd.grad(loss, (x, y))
        out.update({f"x_p{p}": x.detach().numpy(), f"y_p{p}": y.detach().numpy(), f"theta_p{p}": theta.numpy(),
                    f"loss_p{p}": np.float64(loss.item()), f"gx_p{p}": gx.numpy(), f"gy_p{p}": gy.numpy()})
        print("notebook SWD p=%d" % p, loss.item())
    np.savez(os.path.join(HERE, "notebook_sliced_wasserstein.npz"), **out)


class FixedFramesSSW:
    """An ``SSW`` callable with the signature the max-SSW wrapper calls (max_spherical_sliced_w.py:518,532) that replaces the
    random frames of sliced_wasserstein_sphere (:307-308) by a fixed cycle of frames, so a wrapper run is a function of its
    inputs.  ``sliced_cost`` is the reference's (fixture generation) or the drop-in's (the GPU test)."""

    def __init__(self, sliced_cost, Us):
        self.sliced_cost, self.Us, self.k = sliced_cost, Us, 0

    def __call__(self, Xs, Xt, num_projections, device, p=2):
        U = self.Us[self.k % len(self.Us)]
        self.k += 1
        return self.sliced_cost(Xs, Xt, U.to(Xs.device), p=p)


def make_max_ssw_fixture():
    """SURVEY.md 8f #4 tail: one training call of max_spherical_wassersten_distance (max_spherical_sliced_w.py:498-536) with
    the reference's own sphere map transform_to_sphere (:334-350), from the UNMODIFIED reference: two SGD ascent steps on
    phi over detached inputs, then the outer sum over the batch of per-pair sliced costs; p = 2 and p = 1."""
    ssw = _load(os.path.join(REF, "Point_Cloud_Resistration/losses/max_spherical_sliced_w.py"), "ref_ssw_max")
    out = {}
    for p in (2, 1):
        torch.manual_seed(600 + p)
        phi = ssw.transform_to_sphere()
        phi_op = torch.optim.SGD(phi.parameters(), lr=0.05)
        g = torch.Generator().manual_seed(60 + p)
        Us = torch.linalg.qr(torch.randn(5, 24, 3, 2, generator=g)).Q
        first = torch.randn(3, 40, 3, generator=g)
        second = (first[:, torch.randperm(40, generator=g)][:, :33] * 0.8 + 0.1 * torch.randn(3, 33, 3, generator=g)).requires_grad_(True)
        crit = ssw.max_spherical_wassersten_distance(24, phi, FixedFramesSSW(ssw.sliced_cost, Us), phi_op, p=p, max_iter=2, device="cpu")
        sd0 = {k: v.clone() for k, v in phi.state_dict().items()}
        val, ft, st = crit(first, second, "train")
        (g_second,) = torch.autograd.grad(val, second)
        sd1 = {k: v.clone() for k, v in phi.state_dict().items()}
        val_test, _, _ = crit(first, second.detach(), "test")
        out.update({f"first_p{p}": first.numpy(), f"second_p{p}": second.detach().numpy(), f"Us_p{p}": Us.numpy(),
                    f"ssw_p{p}": np.float64(val.item()), f"first_t_p{p}": ft.detach().numpy(), f"second_t_p{p}": st.detach().numpy(),
                    f"g_second_p{p}": g_second.numpy(), f"ssw_test_p{p}": np.float64(val_test.item())})
        out.update(_flat_state(f"sd0_p{p}__", sd0))
        out.update(_flat_state(f"sd1_p{p}__", sd1))
        print("max SSW wrapper p=%d" % p, val.item(), "test", val_test.item())
    out.update(lr=np.float64(0.05), max_iter=np.int64(2), num_projections=np.int64(24))
    np.savez(os.path.join(HERE, "max_ssw_wrapper.npz"), **out)


def make_ssw_fast_fixture():
    """The batched variant max_spherical_sliced_w_fast.py:258-295 (`sliced_cost` with per-pair frames (B,P,3,2): projection by
    one broadcast matmul, then a loop over the batch of mean_P binary_search_circle; returns the SUM over the pairs, shape (1,))
    and one training call of max_spherical_wassersten_distance_fast (:346-382) on it, from the unmodified reference."""
    fast = _load(os.path.join(REF, "Point_Cloud_Resistration/losses/max_spherical_sliced_w_fast.py"), "ref_ssw_fast")
    g = torch.Generator().manual_seed(91)
    out = {}
    for p in (2, 3):
        Us = torch.linalg.qr(torch.randn(3, 16, 3, 2, generator=g)).Q
        x = F.normalize(torch.randn(3, 40, 3, generator=g), dim=-1).requires_grad_(True)
        y = F.normalize(torch.randn(3, 33, 3, generator=g), dim=-1).requires_grad_(True)
        w = fast.sliced_cost(x, y, Us, p=p)
        gx, gy = torch.autograd.grad(w.sum(), (x, y))
        out.update({f"Us_p{p}": Us.numpy(), f"x_p{p}": x.detach().numpy(), f"y_p{p}": y.detach().numpy(), f"w_p{p}": w.detach().numpy(),
                    f"gx_p{p}": gx.numpy(), f"gy_p{p}": gy.numpy()})
        print("ssw fast p=%d" % p, w.detach().numpy())
    # the wrapper, with the frames of every SSW call fixed (one (B,P,3,2) set per call, cycled)
    torch.manual_seed(92)
    phi = fast.transform_to_sphere_fast()
    phi_op = torch.optim.SGD(phi.parameters(), lr=0.05)
    Uc = torch.linalg.qr(torch.randn(4, 3, 16, 3, 2, generator=g)).Q
    first = torch.randn(3, 40, 3, generator=g)
    second = (first[:, :33] * 0.8 + 0.1 * torch.randn(3, 33, 3, generator=g)).requires_grad_(True)
    crit = fast.max_spherical_wassersten_distance_fast(16, phi, FixedFramesSSW(fast.sliced_cost, Uc), phi_op, p=2, max_iter=2, device="cpu")
    sd0 = {k: v.clone() for k, v in phi.state_dict().items()}
    val, ft, st = crit(first, second, "train")
    (g_second,) = torch.autograd.grad(val.sum(), second)
    out.update(Uc=Uc.numpy(), first=first.numpy(), second=second.detach().numpy(), ssw=val.detach().numpy(), first_t=ft.detach().numpy(),
               second_t=st.detach().numpy(), g_second=g_second.numpy())
    out.update(_flat_state("sd0__", sd0))
    out.update(_flat_state("sd1__", phi.state_dict()))
    # the same call in float64 (same code, same weights): how far the reference's float32 gradient is from its own exact value
    phi64 = fast.transform_to_sphere_fast().double()
    phi64.load_state_dict({k: v.double() for k, v in sd0.items()})
    op64 = torch.optim.SGD(phi64.parameters(), lr=0.05)
    crit64 = fast.max_spherical_wassersten_distance_fast(16, phi64, FixedFramesSSW(fast.sliced_cost, Uc.double()), op64, p=2, max_iter=2,
                                                         device="cpu")
    second64 = second.detach().double().requires_grad_(True)
    val64, _, _ = crit64(first.double(), second64, "train")
    (g64,) = torch.autograd.grad(val64.sum(), second64)
    out.update(g_second_f64=g64.numpy(), ssw_f64=val64.detach().double().numpy())
    print("max SSW fast wrapper", val.detach().numpy(), "float32 vs float64 reference: d/dsecond",
          float((g_second.double() - g64).norm() / g64.norm()))
    np.savez(os.path.join(HERE, "ssw_fast.npz"), **out)


def make_mini_batch_mssw_fixture():
    """mini_batch_Residual_MSSW.py from the UNMODIFIED reference: (A) its sphere map transform_to_sphere (:327-408: MLP -> flows on
    R^2 -> angles -> S^2) for both flow kinds -- the first forward runs ActNorm's data-dependent initialisation, the second one is
    differentiated; (B) one training call of max_spherical_wassersten_distance_Residual (:413-452) with a Planar map, the
    mini-batches drawn by np.random.choice after np.random.seed(12), the frames of every SSW call fixed."""
    mb = _load(os.path.join(REF, "Point_Cloud_Resistration/losses/mini_batch_Residual_MSSW.py"), "ref_mini_batch_mssw")
    out = {}
    for fl in ("Planar", "Residual"):
        torch.manual_seed(700 if fl == "Planar" else 701)
        np.random.seed(7)
        phi = mb.transform_to_sphere(fl, n_flow_layer=2)
        with torch.no_grad():
            for prm in phi.parameters():
                if prm.dim() > 0 and prm.dtype == torch.float32:
                    prm.add_(0.2 * torch.randn_like(prm))
        phi.train()
        sd0 = {k: v.clone() for k, v in phi.state_dict().items()}
        g = torch.Generator().manual_seed(70)
        x1 = torch.randn(4, 30, 3, generator=g)
        y1 = phi(x1)
        x2 = torch.randn(4, 30, 3, generator=g).requires_grad_(True)
        w = torch.randn(4, 30, 3, generator=g)
        y2 = phi(x2)
        named = [(n, q) for n, q in phi.named_parameters() if q.dtype == torch.float32 and q.dim() > 0]
        gs = torch.autograd.grad((y2 * w).sum(), [x2] + [q for _, q in named], allow_unused=True)
        out.update({f"{fl}_x1": x1.numpy(), f"{fl}_y1": y1.detach().numpy(), f"{fl}_x2": x2.detach().numpy(), f"{fl}_w": w.numpy(),
                    f"{fl}_y2": y2.detach().numpy(), f"{fl}_gx": gs[0].numpy()})
        for (n, q), gq in zip(named, gs[1:]):
            out[f"{fl}_gp__" + n.replace(".", "__")] = (torch.zeros_like(q) if gq is None else gq).numpy()
        out.update(_flat_state(f"{fl}_sd0__", sd0))
        print("mini-batch MSSW sphere map", fl, "|y2|", float(y2.norm()), "|gx|", float(gs[0].norm()))
    torch.manual_seed(702)
    phi = mb.transform_to_sphere("Planar", n_flow_layer=2)
    phi_op = torch.optim.SGD(phi.parameters(), lr=0.05)
    g = torch.Generator().manual_seed(71)
    Us = torch.linalg.qr(torch.randn(5, 24, 3, 2, generator=g)).Q
    first = torch.randn(4, 40, 3, generator=g)
    second = (first[:, :33] * 0.8 + 0.1 * torch.randn(4, 33, 3, generator=g)).requires_grad_(True)
    crit = mb.max_spherical_wassersten_distance_Residual(24, phi, phi_op, SSW=FixedFramesSSW(mb.sliced_cost, Us), p=2, max_iter=2,
                                                         psi_minibatch_size=2, device="cpu")
    sd0 = {k: v.clone() for k, v in phi.state_dict().items()}
    np.random.seed(12)
    val, ft, st = crit(first, second, "train")
    (g_second,) = torch.autograd.grad(val, second)
    out.update(Us=Us.numpy(), first=first.numpy(), second=second.detach().numpy(), ssw=np.float64(val.item()), first_t=ft.detach().numpy(),
               second_t=st.detach().numpy(), g_second=g_second.numpy())
    out.update(_flat_state("w_sd0__", sd0))
    out.update(_flat_state("w_sd1__", phi.state_dict()))
    print("mini-batch MSSW wrapper", val.item())
    np.savez(os.path.join(HERE, "mini_batch_mssw.npz"), **out)


if __name__ == "__main__":
    if "--mini-batch-only" in sys.argv:
        _install_shims()
        torch.set_num_threads(8)
        make_mini_batch_mssw_fixture()
    elif "--max-ssw-only" in sys.argv:  # the fixtures added in the last session of round 2
        _install_shims()
        torch.set_num_threads(8)
        make_max_ssw_fixture()
        make_ssw_fast_fixture()
    elif "--wrappers-only" in sys.argv:  # regenerate only the fixtures added in round 2 (the older ones are unchanged)
        _install_shims()
        torch.set_num_threads(8)
        import losses as _ref_losses
        make_wrapper_fixtures(_ref_losses)
        make_notebook_fixture()
        make_weighted_circle_fixture()
    elif "--weighted-wp-only" in sys.argv:  # the fixture added last in round 2
        _install_shims()
        torch.set_num_threads(8)
        make_weighted_wp_fixture()
    else:
        main()
