"""CPU: host-side mirror of the reference interface -- names, signatures, error behaviour, the phi flows, and the
batch-sharding helper under a world_size-2 gloo group.  No CUDA compute happens here; the product path must refuse
CPU tensors instead of falling back."""
import inspect
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import shwd

L = shwd.losses


def test_exported_names_match_reference_packages():
    # Point_Cloud_Resistration/losses/__init__.py:27-32 and Comparison_.../losses/__init__.py:1-2
    for name in ("Cos_disimilarity_W", "Geodesic_distance_W", "Norm_Flow_structure", "Norm_Flow_structure_optuna",
                 "max_cos_disimilarity_wassersten_distance", "pseudo_max_cos_disimilarity_wassersten_distance",
                 "log_Sinkhorn_Distance_Loss", "log_N_Sinkhorn_Distance_Loss"):
        assert hasattr(L, name), name


def test_signatures_keep_the_reference_positional_arguments():
    def params(fn):
        return list(inspect.signature(fn).parameters)
    assert params(L.Geodesic_distance_W.__init__)[:3] == ["self", "device", "p"]
    assert params(L.Cos_disimilarity_W.__init__)[:3] == ["self", "device", "p"]
    assert params(L.max_cos_disimilarity_wassersten_distance.__init__)[:8] == [
        "self", "phi", "CSW", "device", "phi_op", "max_iter", "lam", "psi_minibatch_size"]
    assert params(L.max_cos_disimilarity_wassersten_distance.forward) == ["self", "first_samples", "second_samples", "train_or_test"]
    assert params(L.pseudo_max_cos_disimilarity_wassersten_distance.__init__)[:8] == [
        "self", "CSW", "device", "phi_num", "lam", "n_flow_layer", "flow_name", "mean_or_max_or_softmax"]
    assert params(L.log_Sinkhorn_Distance_Loss.__init__)[:5] == ["self", "eps", "max_iter", "batch_reduction", "type_of_cost_norm"]
    assert params(L.log_Sinkhorn_Distance_Loss.forward) == ["self", "x", "y", "device"]
    assert params(L.log_N_Sinkhorn_Distance_Loss.__init__)[:6] == ["self", "eps", "max_iter", "batch_reduction", "type_of_cost_norm",
                                                                    "type_of_Wasserstein_N"]
    assert params(L.sliced_wasserstein_sphere)[:7] == ["Xs", "Xt", "num_projections", "device", "u_weights", "v_weights", "p"]
    assert params(L.Norm_Flow_structure.__init__) == ["self", "input_dim", "flow_name", "n_flow_layer"]
    sig = inspect.signature(L.chamfer_distance)
    assert sig.parameters["batch_reduction"].default == "mean" and sig.parameters["point_reduction"].default == "mean"


def test_error_behaviour_matches_reference():
    with pytest.raises(ValueError, match="Flow name is not valid"):  # s2_wasserstein.py:158
        L.Norm_Flow_structure(flow_name="Nope")
    with pytest.raises(ValueError, match="type_of_cost_norm must be 'L1' or 'L2'"):  # Sinkhorn.py:91
        L.Sinkhorn_Distance_Loss(eps=0.1, max_iter=3, type_of_cost_norm="L7")
    crit = L.pseudo_max_cos_disimilarity_wassersten_distance(None, "cpu", phi_num=1, n_flow_layer=1, flow_name="Planar",
                                                             mean_or_max_or_softmax="bogus")
    with pytest.raises(ValueError, match="mean_or_max_or_softmax is not valid"):  # s2_wasserstein.py:344
        crit(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))


def test_no_cpu_fallback():
    x = torch.randn(2, 8, 3)
    for call in (lambda: shwd.entropic_ot(x, x), lambda: shwd.sphere_map(x), lambda: shwd.chamfer_nn(x, x),
                 lambda: L.Geodesic_distance_W("cpu", 2)(x, x), lambda: L.chamfer_distance(x, x)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
    with pytest.raises(ValueError):
        shwd.ops._as_cloud(torch.zeros(2, 3, 4).cuda() if torch.cuda.is_available() else _FakeCuda((2, 3, 4)), "x")


class _FakeCuda(torch.Tensor):
    """A CPU tensor that claims to be CUDA, to reach the shape validation without a GPU."""
    @staticmethod
    def __new__(cls, shape):
        return torch.Tensor._make_subclass(cls, torch.zeros(shape))

    @property
    def is_cuda(self):
        return True


def test_flows_forward_and_parameter_counts():
    torch.manual_seed(0)
    x = torch.randn(2, 16, 3)
    res = L.Norm_Flow_structure(flow_name="Residual", n_flow_layer=3)
    # 1284 parameters for Residual x3 in the reference = 1278 network parameters + 2 (geom_p, lamb) per block;
    # the module tree mirrors the reference's, so parameters() / state_dict() have its names, shapes and order
    assert sum(p.numel() for p in res.parameters()) == 1284
    assert [n for n, _ in res.named_parameters()][:4] == ["net.0.iresblock.geom_p", "net.0.iresblock.lamb",
                                                          "net.0.iresblock.nnet.net.0.beta", "net.0.iresblock.nnet.net.1.weight"]
    y = res(x)
    assert y.shape == x.shape and torch.isfinite(y).all()
    # Lipschitz < 1 residual branch: the map is a contraction-perturbed identity
    a, b = torch.randn(64, 3), torch.randn(64, 3)
    for flow in res.net:
        ga, gb = flow.net(a), flow.net(b)
        assert ((ga - gb).norm(dim=-1) <= 1.0 * (a - b).norm(dim=-1) + 1e-6).all()
    pl = L.Norm_Flow_structure(flow_name="Planar", n_flow_layer=3)
    assert sum(p.numel() for p in pl.parameters()) == 3 * 7
    assert pl(x).shape == x.shape
    opt = L.Norm_Flow_structure_optuna(flow_name="Residual", n_flow_layer=2, Residual_hidden_units=4, Residual_hidden_layers=3)
    assert opt(x).shape == x.shape


def test_regularizer_cpu_path_of_wrapper_matches_formula():
    crit = L.max_cos_disimilarity_wassersten_distance(phi=None, CSW=None, device="cpu", phi_op=None)
    x = torch.randn(3, 10, 3)
    assert torch.allclose(crit.regularization_of_normalizing_flow(x), (x.norm(dim=-1) - 1).abs().sum())


def test_shard_range_partitions_exactly():
    from shwd_b200.dist import shard_range
    for n in (0, 1, 7, 32, 33):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from shwd_b200.dist import sharded_pair_loss
    torch.manual_seed(0)
    x = torch.randn(5, 6, 3, requires_grad=True)  # identical on both ranks; 5 pairs -> shards of 3 and 2
    y = torch.randn(5, 6, 3)

    def per_pair(a, b):  # stand-in for the CUDA loss: any per-pair function
        return ((a - b) ** 2).sum(dim=(1, 2))

    loss = sharded_pair_loss(per_pair, x, y)
    loss.backward()
    g = x.grad.clone()
    dist.all_reduce(g)  # DDP would sum the shards' gradients
    ref = per_pair(x.detach(), y).mean()
    ok = torch.allclose(loss.detach(), ref, rtol=1e-6) and torch.allclose(g, 2 * (x.detach() - y) / 5, rtol=1e-5, atol=1e-7)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_mean_equals_single_process_mean_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def _ddp_worker(rank, world, port, out):
    """sharded_pair_loss inside a real DistributedDataParallel step: DDP AVERAGES parameter gradients over the ranks, so the
    loss is built with grad_reduction="mean"; the averaged gradient must be the single-process gradient of the batch mean."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from shwd_b200.dist import sharded_pair_loss
    torch.manual_seed(0)
    lin = torch.nn.Linear(3, 3)
    ref = torch.nn.Linear(3, 3)
    ref.load_state_dict(lin.state_dict())
    model = torch.nn.parallel.DistributedDataParallel(lin)
    x = torch.randn(5, 6, 3)
    y = torch.randn(5, 6, 3)

    def per_pair(a, b):
        return ((a - b) ** 2).sum(dim=(1, 2))

    loss = sharded_pair_loss(per_pair, model(x), y, grad_reduction="mean")
    loss.backward()
    want = per_pair(ref(x), y).mean()
    want.backward()
    ok = torch.allclose(loss.detach(), want.detach(), rtol=1e-6)
    ok = ok and all(torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-7) for a, b in zip(lin.parameters(), ref.parameters()))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_loss_inside_ddp_averaging_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ddp_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def _slice_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from shwd_b200.dist import sharded_slice_loss
    torch.manual_seed(0)
    x = torch.randn(40, 3, requires_grad=True)  # the same clouds and frames on both ranks; 7 slices -> shards of 4 and 3
    y = torch.randn(40, 3)
    th = torch.nn.functional.normalize(torch.randn(7, 3), dim=-1)

    def slice_mean(a, b, frames):  # stand-in for the CUDA sliced loss: mean over the given slices of a per-slice value
        pa, pb = torch.sort(a @ frames.T, dim=0)[0], torch.sort(b @ frames.T, dim=0)[0]
        return ((pa - pb) ** 2).sum(0).mean()

    loss = sharded_slice_loss(slice_mean, x, y, th)
    loss.backward()
    g = x.grad.clone()
    dist.all_reduce(g)
    xr = x.detach().clone().requires_grad_(True)
    ref = slice_mean(xr, y, th)
    ref.backward()
    out[rank] = bool(torch.allclose(loss.detach(), ref.detach(), rtol=1e-6) and torch.allclose(g, xr.grad, rtol=1e-5, atol=1e-7))
    dist.destroy_process_group()


def test_slice_sharded_mean_equals_single_process_mean_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_slice_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_fused_phi_parameter_layout_matches_the_c_abi():
    """The parameter tensors Norm_Flow_structure hands to the fused kernel, concatenated in order, must fill exactly the
    layout include/shwd.h documents (426 raw parameters and 102 power-iteration entries per Residual flow)."""
    from shwd_b200 import _lib
    from shwd_b200.losses import flows
    lib = _lib.lib()
    per, uvper = lib.shwd_resflow_params_per_layer(), lib.shwd_resflow_uv_per_layer()
    assert (per, uvper) == (8 * 3 + 8 + 5 * (64 + 8) + 3 * 8 + 3 + 7, (8 + 3) + 5 * 16 + (3 + 8))
    phi = L.Norm_Flow_structure(flow_name="Residual", n_flow_layer=3)
    assert flows.is_standard_residual_stack(phi.net)
    raw = []
    for f in phi.net:
        raw += flows._raw_params(f)
    assert sum(p.numel() for p in raw) == 3 * per
    # every network parameter, once (geom_p / lamb of the discarded log-determinant estimator never reach the kernel)
    assert {id(p) for p in raw} == {id(p) for n, p in phi.named_parameters() if not n.endswith(("geom_p", "lamb"))}
    assert flows._uv_buffer(phi.net).numel() == 3 * uvper
    assert lib.shwd_resflow_workspace_bytes(1000, 3) == (8 + 1) * 3 * per * 4
    # non-standard shapes and Planar stacks stay on the eager modules
    assert not flows.is_standard_residual_stack(L.Norm_Flow_structure(flow_name="Planar").net)
    assert not flows.is_standard_residual_stack(L.Norm_Flow_structure_optuna(flow_name="Residual", Residual_hidden_units=4).net)
    x = torch.randn(5, 3)
    assert torch.equal(phi(x), phi.forward_eager(x))  # CPU tensors: the eager path


def test_exact_solver_option_and_limits():
    with pytest.raises(ValueError, match="solver must be"):
        L.Cos_disimilarity_W("cpu", p=2, solver="simplex")
    assert L.Geodesic_distance_W("cpu", p=2, solver="exact").solver == "exact"
    from shwd_b200 import _lib
    assert _lib.lib().shwd_exact_assignment_max_points() >= 2048
    # clouds of different sizes: the assignment of lcm(n, m) copies, when that fits the kernel
    from shwd_b200 import ops
    assert ops.exact_copies(512, 1024) == (2, 1) and ops.exact_copies(6, 4) == (2, 3) and ops.exact_copies(1000, 1024) is None
    assert list(inspect.signature(L.binary_search_circle).parameters) == [
        "u_values", "v_values", "u_weights", "v_weights", "p", "Lm", "Lp", "tm", "tp", "eps", "require_sort"]


def test_pose_generator_consumes_numpy_like_the_reference():
    """shwd.data.random_poses against the frozen reference poses (numpy seed 1234, tests/golden/make_golden_data.py)."""
    import numpy as np
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "rigid_transform.npz"))
    poses = shwd.data.random_poses(d["poses"].shape[0], 45, 1, np.random.RandomState(1234))
    assert np.array_equal(poses.numpy(), d["poses"])
    assert np.allclose(shwd.data.euler_to_quaternion_xyz(d["euler_in"]), d["euler_quat"], rtol=0, atol=1e-15)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        shwd.data.rigid_transform(torch.zeros(1, 4, 3), poses[:1])


def test_graphed_loss_refuses_cpu_tensors():
    gfn = shwd.graphed_loss(lambda a, b: (a - b).pow(2).sum())
    with pytest.raises(RuntimeError, match="CUDA"):
        gfn(torch.randn(4, 3), torch.randn(4, 3))
    assert len(gfn._captures) == 0


def test_ot_dropin_exports_emd2_and_refuses_cpu_matrices():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin_ot"))
    import ot
    assert ot.__all__ == ["emd2"]
    w = torch.full((4,), 0.25)
    with pytest.raises(RuntimeError, match="CUDA"):
        ot.emd2(w, w, torch.rand(4, 4))
    with pytest.raises(NotImplementedError):
        ot.emd2(w, w, torch.rand(4, 5))


def test_stiefel_frames_equal_torch_qr():
    """sliced_wasserstein_sphere draws qr(randn(P,d,2)).Q (max_spherical_sliced_w.py:307-308); the elementwise restatement
    reproduces LAPACK's Householder signs, so the frames equal torch.linalg.qr's to rounding."""
    from shwd_b200.losses.sliced import stiefel_frames
    torch.manual_seed(0)
    for d in (3, 4):
        Z = torch.randn(2000, d, 2)
        Q, _ = torch.linalg.qr(Z)
        U = stiefel_frames(Z)
        assert (U - Q).abs().max().item() < 2e-5
        assert (U.transpose(1, 2) @ U - torch.eye(2)).abs().max().item() < 2e-5


def test_bench_reference_arm_and_meta_describe_the_same_workload():
    """The driver compares the two arms' `config`: both must come from workload_meta, for every --config."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for cfg in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"):
        m = bench.workload_meta(cfg, "weak", 1)
        assert set(m) == {"metric", "unit", "config"} and "workload" in m["config"]
        assert cfg in bench.CPU_SAMPLE
    assert bench.workload_meta("cfg2", "strong", 8)["config"]["global_batch"] == 32
    assert bench.workload_meta("cfg2", "weak", 8)["config"]["global_batch"] == 256


def test_weight_cdf_equals_the_reference_cumsum_bit_for_bit():
    """ops.weight_cdf (the per-slice CDF tables of the weighted circular W_p) against the reference's own expression
    ``torch.cumsum(weights[..., sorter], -1)`` on the CPU (max_spherical_sliced_w.py:166-170): the same float32 bits, for (n,)
    weights gathered through a (S,n) sorter and for the uniform default."""
    g = torch.Generator().manual_seed(4)
    S, n = 7, 1531
    w = torch.rand(n, generator=g) + 0.01
    w = w / w.sum()
    perm = torch.argsort(torch.rand(S, n, generator=g), dim=-1)
    got = shwd.ops.weight_cdf(w, perm.int(), S, n, torch.device("cpu"))
    want = torch.cumsum(w[..., perm], -1)
    assert got.dtype == torch.float32 and torch.equal(got, want)
    uni = shwd.ops.weight_cdf(None, None, S, n, torch.device("cpu"))
    assert torch.equal(uni, torch.cumsum(torch.full((n,), 1 / n, dtype=torch.float32), -1).expand(S, n))
    # and the closed form the uniform kernel evaluates in registers: fl((i + 1) * fl(1 / n))
    closed = (torch.arange(1, n + 1, dtype=torch.float64) * float(torch.tensor(1 / n, dtype=torch.float32))).float()
    assert torch.equal(uni[0], closed)


def test_transform_to_sphere_is_the_reference_module():
    """max_spherical_sliced_w.py:334-350: same state_dict keys / shapes, and -- loaded with the weights the unmodified reference
    ended its ascent with -- the same points on the sphere (fixture of tests/golden/make_golden.py::make_max_ssw_fixture)."""
    import numpy as np
    d = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "max_ssw_wrapper.npz")))
    phi = L.transform_to_sphere()
    sd = {k[len("sd1_p2__"):].replace("__", "."): torch.from_numpy(v) for k, v in d.items() if k.startswith("sd1_p2__")}
    assert list(phi.state_dict().keys()) == list(sd.keys()) == ["net.0.weight", "net.0.bias", "net.2.weight", "net.2.bias",
                                                               "net.4.weight", "net.4.bias"]
    phi.load_state_dict(sd)
    y = phi(torch.from_numpy(d["first_p2"]))
    assert torch.allclose(y, torch.from_numpy(d["first_t_p2"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(y.norm(dim=-1), torch.ones(3, 40), atol=1e-6)


@pytest.mark.parametrize("fl", ["Planar", "Residual"])
def test_mini_batch_mssw_sphere_map_matches_the_reference_fixture(fl):
    """mini_batch_Residual_MSSW.py:327-408 (MLP -> flows on R^2, incl. ActNorm's data-dependent first-batch initialisation ->
    angles -> S^2): loaded with the reference's own state_dict, the first forward (initialising) and a second, differentiated
    one against outputs and autograd gradients frozen from the unmodified reference (make_golden.py::make_mini_batch_mssw_fixture)."""
    import numpy as np
    from shwd_b200.losses import mini_batch_mssw as M
    d = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mini_batch_mssw.npz")))
    pre = fl + "_sd0__"
    sd = {k[len(pre):].replace("__", "."): torch.from_numpy(v) for k, v in d.items() if k.startswith(pre)}
    phi = M.transform_to_sphere(fl, n_flow_layer=2)
    res = phi.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    phi.train()
    y1 = phi(torch.from_numpy(d[fl + "_x1"]))
    x2 = torch.from_numpy(d[fl + "_x2"]).requires_grad_(True)
    y2 = phi(x2)
    named = [(n, q) for n, q in phi.named_parameters() if q.dtype == torch.float32 and q.dim() > 0]
    gs = torch.autograd.grad((y2 * torch.from_numpy(d[fl + "_w"])).sum(), [x2] + [q for _, q in named], allow_unused=True)

    def rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-3)).item()

    assert rel(y1.detach(), torch.from_numpy(d[fl + "_y1"])) < 1e-6 and rel(y2.detach(), torch.from_numpy(d[fl + "_y2"])) < 1e-6
    assert rel(gs[0], torch.from_numpy(d[fl + "_gx"])) < 1e-5
    for (n, q), g in zip(named, gs[1:]):
        want = torch.from_numpy(d[fl + "_gp__" + n.replace(".", "__")])
        got = torch.zeros_like(want) if g is None else g
        assert got.shape == want.shape and rel(got, want) < 2e-5, n
