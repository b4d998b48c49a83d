"""Run the reference's UNMODIFIED training entry point (Point_Cloud_Resistration/train_W_COS.py: train -> train_one_epoch
/ test_one_epoch / the snapshot code, load_checkpoint) for two epochs on a tiny synthetic dataset, with either

    mode = reference   the reference's own `losses` package (exact EMD through a scipy-backed `ot` shim), or
    mode = dropin      this repository's drop-in `losses` package (dropin/ first on sys.path),

and print one JSON line: per-epoch losses as the reference's own run.log reports them, a digest of the trained PCRNet /
phi weights, and what a cross-loaded checkpoint of the OTHER mode does.  Executed in a subprocess by
tests/test_reference_entry.py -- only in the build container (needs /root/reference; pure CPU).

The drop-in's CUDA kernels cannot run here, so in `dropin` mode the innermost distance CSW is the oracle's exact solve on
the CPU; everything around it -- the import of `losses`, Norm_Flow_structure, max_cos_disimilarity_wassersten_distance,
.phi / .phi_op, state_dict round trips through the reference's torch.save / load_checkpoint -- is the shipped host code.
"""
import json
import os
import sys
import tempfile
import types

REF = "/root/reference/Point_Cloud_Resistration"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mode, workdir, other_ckpt = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
sys.dont_write_bytecode = True

# ---- stubs for what the entry script imports but this path never executes ------------------------------------------
for name in ("torch_geometric", "torch_geometric.transforms", "torch_geometric.datasets", "torch_geometric.data",
             "torch_geometric.utils", "transforms3d", "transforms3d.euler", "tensorboardX"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["torch_geometric"].__path__ = []  # a package: `from torch_geometric.data import Data` must resolve
sys.modules["torch_geometric.datasets"].ModelNet = object
sys.modules["torch_geometric.data"].Data = sys.modules["torch_geometric.data"].Batch = object
sys.modules["torch_geometric.utils"].to_networkx = None
sys.modules["transforms3d"].euler = sys.modules["transforms3d.euler"]


class _Writer:
    def __init__(self, *a, **k):
        self.rows = []

    def add_scalar(self, tag, value, step):
        self.rows.append((tag, float(value), int(step)))


sys.modules["tensorboardX"].SummaryWriter = _Writer

import numpy as np  # noqa: E402
import torch  # noqa: E402


def _euler2mat_stub(*a, **k):
    return np.eye(3)


sys.modules["transforms3d.euler"].mat2euler = lambda m, axes="sxyz": (0.0, 0.0, 0.0)
sys.modules["transforms3d.euler"].euler2mat = _euler2mat_stub


def _mat2axangle(m):  # what test_one_epoch's error metric needs from transforms3d (train_W_COS.py:83): the rotation angle
    c = np.clip((np.trace(np.asarray(m, dtype=np.float64)) - 1.0) / 2.0, -1.0, 1.0)
    return np.array([0.0, 0.0, 1.0]), float(np.arccos(c))


sys.modules["transforms3d.axangles"] = types.ModuleType("transforms3d.axangles")
sys.modules["transforms3d.axangles"].mat2axangle = _mat2axangle
sys.modules["transforms3d"].axangles = sys.modules["transforms3d.axangles"]

shim = tempfile.mkdtemp(prefix="shwd_entry_")
if mode == "reference":
    os.symlink(os.path.join(REF, "losses/normflows_ishikawa"), os.path.join(shim, "normflows"))
    os.makedirs(os.path.join(shim, "matplotlib"))
    open(os.path.join(shim, "matplotlib/__init__.py"), "w").close()
    open(os.path.join(shim, "matplotlib/pyplot.py"), "w").close()
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
    sys.path[:0] = [shim, REF]
else:
    sys.path[:0] = [os.path.join(ROOT, "dropin"), REF, ROOT]  # `import losses` -> dropin/losses

os.chdir(workdir)
import train_W_COS as entry  # noqa: E402  (the unmodified reference script)
import losses  # noqa: E402

if mode == "dropin":
    assert os.path.realpath(losses.__file__).startswith(os.path.realpath(ROOT)), losses.__file__
    import oracle

    class CSW(torch.nn.Module):  # CPU stand-in for the CUDA solve: the oracle's exact EMD on the reference's cost matrix
        def forward(self, x, y):
            C = oracle.cost_matrix(x, y, "sqeuclid", 2)
            tot = 0
            for b in range(C.shape[0]):
                _, plan = oracle.exact_emd2(C[b])
                tot = tot + torch.pow((plan.to(C.dtype) * C[b]).sum(), 1.0 / 2)
            return tot / C.shape[0]
    csw = CSW()
else:
    assert losses.__file__.startswith(REF), losses.__file__
    csw = losses.Cos_disimilarity_W("cpu", p=2)


class Pairs(torch.utils.data.Dataset):
    """What Dataset_pytorch.__getitem__ returns (data_utils/Data_set_maker.py:238-259): template, source, rotation, translation."""

    def __init__(self, n, pts, seed):
        g = torch.Generator().manual_seed(seed)
        self.t = torch.nn.functional.normalize(torch.randn(n, pts, 3, generator=g), dim=-1) * torch.tensor([1.0, 0.7, 0.5])
        ang = 0.3 * torch.randn(n, generator=g)
        R = torch.zeros(n, 3, 3)
        R[:, 0, 0], R[:, 0, 1], R[:, 1, 0], R[:, 1, 1], R[:, 2, 2] = ang.cos(), -ang.sin(), ang.sin(), ang.cos(), 1.0
        self.R, self.tr = R, 0.1 * torch.randn(n, 1, 3, generator=g)
        self.s = self.t @ R.transpose(1, 2) + self.tr + 0.01 * torch.randn(n, pts, 3, generator=g)

    def __len__(self):
        return self.t.shape[0]

    def __getitem__(self, i):
        return self.t[i], self.s[i], self.R[i], self.tr[i]


entry.fix_seed(1234)
np.random.seed(7)
torch.manual_seed(7)
dev = torch.device("cpu")
model = entry.PCRNet(feature_model=entry.MLP_Architecture()).to(dev)
optimizer = torch.optim.Adam(model.parameters(), lr=1e-3)
phi = losses.Norm_Flow_structure(flow_name="Residual", n_flow_layer=3).to(dev)   # train_W_COS.py:390
phi_op = torch.optim.Adam(phi.parameters(), lr=1e-3)                              # :391-392
crit = losses.max_cos_disimilarity_wassersten_distance(phi, csw, dev, phi_op, max_iter=1, lam=0.1)  # :404
train_loader = torch.utils.data.DataLoader(Pairs(4, 48, 1), batch_size=2, shuffle=False)
test_loader = torch.utils.data.DataLoader(Pairs(2, 48, 2), batch_size=2, shuffle=False)
os.makedirs(os.path.join(workdir, "log/models"), exist_ok=True)
writer = _Writer()


class Log:
    lines = []

    def cprint(self, text):
        self.lines.append(text)


init = {k: v.clone() for k, v in phi.state_dict().items()}
entry.train(crit, crit, 2, dev, optimizer, 2, model, train_loader, test_loader, writer, os.path.join(workdir, "log"), Log(), None)


def digest(sd):
    # (last_n_samples / last_firmom / last_secmom are the random bookkeeping of the log-determinant estimator whose result
    #  the loss path discards -- x, _ = flow(x), s2_wasserstein.py:160-163 -- and the drop-in never runs)
    return float(sum(v.double().abs().sum().item() for k, v in sd.items() if v.is_floating_point() and ".last_" not in k))


out = {"mode": mode, "scalars": [r for r in writer.rows if r[0] in ("Train Loss", "Test Loss")],
       "model_digest": digest(model.state_dict()), "phi_digest": digest(phi.state_dict()),
       "phi_keys": list(phi.state_dict().keys())[:6], "n_phi_params": sum(p.numel() for p in phi.parameters())}
snap = os.path.join(workdir, "log/models/best_model_snap.t7")
out["snapshot"] = snap
# ---- resume through the reference's own load_checkpoint (train_W_COS.py:252-276), from this run's snapshot and from the
# snapshot the OTHER mode wrote (a reference .t7 must load into the drop-in's phi / phi_op, and vice versa)
for tag, path in (("own", snap), ("other", other_ckpt)):
    if not path:
        continue
    m2 = entry.PCRNet(feature_model=entry.MLP_Architecture())
    o2 = torch.optim.Adam(m2.parameters(), lr=1e-3)
    p2 = losses.Norm_Flow_structure(flow_name="Residual", n_flow_layer=3)
    po2 = torch.optim.Adam(p2.parameters(), lr=1e-3)
    m2, o2, p2, po2, ep = entry.load_checkpoint(m2, o2, p2, po2, dev, path)
    x = Pairs(1, 48, 3).t
    p2.eval()
    out["resume_" + tag] = {"epoch": int(ep), "phi_digest": digest(p2.state_dict()), "phi_out": float(p2(x).double().abs().sum().item()),
                            "phi_op_states": len(po2.state_dict()["state"])}
print("RESULT " + json.dumps(out))
