"""The learned sphere map phi (caller of the hot path): compact eager-torch restatement of the two
flows ``Norm_Flow_structure`` builds from the vendored normflows 1.7.2 copy
(``Point_Cloud_Resistration/losses/normflows_ishikawa``): ``flows.Planar`` (flows/planar.py:8-64) and
``flows.Residual`` over ``nets.LipschitzMLP`` (flows/residual.py:12-124, nets/lipschitz.py:14-68,132-293,642-648).
Only the forward map is kept -- the reference discards the log-determinant (``x, _ = flow(x)``,
s2_wasserstein.py:160-163).  The modules below own the parameters and define the map; on CUDA inputs a stack of
standard Residual flows (hidden 8, 7 layers -- what ``Norm_Flow_structure`` builds) runs as ONE fused kernel per direction
(``csrc/resflow.cu``, ``fused_residual_stack``) instead of ~45 eager kernels per flow layer, and a stack of Planar flows
on (N,3) or (B,N,3) inputs likewise (``csrc/planar.cu``, ``fused_planar_stack``); anything else (other widths, CPU tensors)
runs the eager modules.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn


class PlanarFlow(nn.Module):
    """f(z) = z + u^ tanh(lin + b), u^ = u + (softplus(<w,u>) - 1 - <w,u>) w / |w|^2  (flows/planar.py:49-60).

    ``lin`` is the reference's expression verbatim in behaviour: ``sum(w * z, dims 1 .. w.dim()-1, keepdim=True)`` with
    ``w`` of shape (1, dim) -- i.e. a sum over dim 1 of the input.  For (N, dim) inputs that is <w, z_n>; for the batched
    (B, N, dim) clouds the loss wrappers feed it (s2_wasserstein.py:243-246) it sums over the N points per coordinate.
    That is what the reference computes, so it is what is computed here."""

    def __init__(self, dim=3):
        super().__init__()
        self.u = nn.Parameter(torch.empty(1, dim).uniform_(-math.sqrt(2), math.sqrt(2)))
        lim_w = math.sqrt(2.0 / dim)
        self.w = nn.Parameter(torch.empty(1, dim).uniform_(-lim_w, lim_w))
        self.b = nn.Parameter(torch.zeros(1))

    def forward(self, z):
        lin = torch.sum(self.w * z, list(range(1, self.w.dim())), keepdim=True) + self.b
        inner = torch.sum(self.w * self.u)
        u = self.u + (torch.log(1 + torch.exp(inner)) - 1 - inner) * self.w / torch.sum(self.w ** 2)
        return z + u * torch.tanh(lin)


class Swish(nn.Module):
    """x * sigmoid(x * softplus(beta)) / 1.1  (nets/lipschitz.py:642-648)."""

    def __init__(self):
        super().__init__()
        self.beta = nn.Parameter(torch.tensor([0.5]))

    def forward(self, x):
        return (x * torch.sigmoid(x * F.softplus(self.beta))).div(1.1)


class SpectralLinear(nn.Module):
    """Linear layer softly normalised to spectral norm <= coeff: W / max(1, (u^T W v)/coeff) with u, v frozen after 200
    power iterations at construction (InducedNormLinear with domain = codomain = 2, nets/lipschitz.py:132-274;
    the reference never calls update_lipschitz, so u and v stay fixed).  Parameters and buffers carry the reference's
    names (weight, bias; scale, u, v), so a reference checkpoint's entries load unchanged."""

    def __init__(self, in_features, out_features, coeff=0.95, zero_init=False):
        super().__init__()
        self.coeff = coeff
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if zero_init:
            self.weight.data.div_(1000)
        bound = 1 / math.sqrt(in_features)
        nn.init.uniform_(self.bias, -bound, bound)
        u = F.normalize(torch.randn(out_features), dim=0)
        v = F.normalize(torch.randn(in_features), dim=0)
        with torch.no_grad():
            for _ in range(200):
                u = F.normalize(torch.mv(self.weight, v), dim=0)
                v = F.normalize(torch.mv(self.weight.t(), u), dim=0)
            sigma = torch.dot(u, torch.mv(self.weight, v))
        self.register_buffer("scale", sigma.detach().clone())  # the last u^T W v seen (nets/lipschitz.py:266-268)
        self.register_buffer("u", u)
        self.register_buffer("v", v)

    def forward(self, x):
        sigma = torch.dot(self.u, torch.mv(self.weight, self.v))
        with torch.no_grad():
            self.scale.copy_(sigma)
        factor = torch.clamp(sigma / self.coeff, min=1.0)
        return F.linear(x, self.weight / factor, self.bias)


class LipschitzMLP(nn.Module):
    """[Swish, SpectralLinear] per layer, last layer initialised near zero (nets/lipschitz.py:14-68)."""

    def __init__(self, channels, lipschitz_const=0.95):
        super().__init__()
        layers = []
        for i in range(len(channels) - 1):
            layers += [Swish(), SpectralLinear(channels[i], channels[i + 1], lipschitz_const, zero_init=(i == len(channels) - 2))]
        self.net = nn.Sequential(*layers)

    def forward(self, x):
        return self.net(x)


class _IResBlock(nn.Module):
    """Container with the reference's iResBlock state (flows/residual.py:82-124): the Lipschitz net under ``nnet`` plus
    the parameters / buffers of the log-determinant estimator.  The loss path discards the log-determinant
    (``x, _ = flow(x)``, s2_wasserstein.py:160-163), so ``geom_p``, ``lamb`` and the moment buffers are carried only to
    keep ``state_dict()`` / ``parameters()`` identical in names, shapes and order to the reference's -- what
    train_W_COS.py:204-205,260-263 saves and restores for phi and its optimiser."""

    def __init__(self, nnet, geom_p=0.5, lamb=2.0, n_samples=1):
        super().__init__()
        self.nnet = nnet
        self.geom_p = nn.Parameter(torch.tensor(math.log(geom_p) - math.log(1.0 - geom_p), dtype=torch.float64))
        self.lamb = nn.Parameter(torch.tensor(lamb))
        self.register_buffer("last_n_samples", torch.zeros(n_samples))
        self.register_buffer("last_firmom", torch.zeros(1))
        self.register_buffer("last_secmom", torch.zeros(1))

    def forward(self, x):
        return x + self.nnet(x)


class ResidualFlow(nn.Module):
    """x + LipschitzMLP(x) (flows/residual.py:12-68 with reverse=False); module tree as in the reference:
    ``iresblock.nnet.net.{2i}`` = Swish, ``.{2i+1}`` = the spectrally normalised linear layer."""

    def __init__(self, dim=3, hidden_units=8, hidden_layers=7, lipschitz_const=0.95, reverse=False):
        super().__init__()
        channels = [dim] + [hidden_units] * (hidden_layers - 1) + [dim]
        self.iresblock = _IResBlock(LipschitzMLP(channels, lipschitz_const))
        # reverse=True is normflows' constructor default (flows/residual.py:21,63-65): ``forward`` then applies the INVERSE of
        # x + g(x); s2_wasserstein.py:152 passes reverse=False, mini_batch_Residual_MSSW.py:381 keeps the default
        self.reverse = reverse

    @property
    def net(self):
        return self.iresblock.nnet.net

    def _inverse_fixed_point(self, y, atol=1e-5, rtol=1e-5):
        """x with x + g(x) = y by the iteration x <- y - g(x) until every entry moved by less than the tolerance (at most 1000
        rounds), differentiated through its iterations like the reference's (flows/residual.py:133-142)."""
        g = self.iresblock.nnet
        x, x_prev = y - g(y), y
        i = 0
        tol = atol + y.abs() * rtol
        while not torch.all((x - x_prev) ** 2 / tol < 1):
            x, x_prev = y - g(x), x
            i += 1
            if i > 1000:
                break
        return x

    def forward(self, x):
        if self.reverse:
            return self._inverse_fixed_point(x)
        return self.iresblock(x)


def _raw_params(flow):
    """Raw parameter tensors of one standard ResidualFlow in the layout order of include/shwd.h:
    W0, b0, ..., W6, b6, beta0..beta6."""
    mods = list(flow.net)
    swish, lins = mods[0::2], mods[1::2]
    out = []
    for lin in lins:
        out += [lin.weight, lin.bias]
    return out + [sw.beta for sw in swish]


def _uv_buffer(flows):
    """The frozen power-iteration vectors of every linear layer, flat: per flow [u0 | v0 | ... | u6 | v6]."""
    parts = []
    for f in flows:
        for lin in list(f.net)[1::2]:
            parts += [lin.u, lin.v]
    return torch.cat(parts).float().contiguous()


def is_standard_residual_stack(flows):
    """True when every flow is a ResidualFlow with channels [3, 8 x6, 3] -- the shape the fused kernel implements."""
    for f in flows:
        if not isinstance(f, ResidualFlow) or f.reverse:
            return False
        lins = list(f.net)[1::2]
        shapes = [tuple(l.weight.shape) for l in lins]
        if shapes != [(8, 3)] + [(8, 8)] * 5 + [(3, 8)]:
            return False
    return 0 < len(flows) <= 8


def fused_residual_stack(flows, x, uv=None):
    """phi(x) for a standard Residual stack through the fused CUDA kernel (x: (...,3) CUDA tensor).  ``uv``: the cached
    result of ``_uv_buffer`` (the vectors are buffers that never change after construction)."""
    from .. import ops
    if uv is None:
        uv = _uv_buffer(flows)
    params = []
    for f in flows:
        params += _raw_params(f)
    coeff = list(flows[0].net)[1].coeff
    return ops.ResidualFlowStackFn.apply(x, uv.to(x.device), len(flows), coeff, *params).reshape(x.shape)


def is_planar_stack(flows):
    """True when every flow is a 3-D PlanarFlow and the stack is no deeper than the fused kernel's limit (8)."""
    return 0 < len(flows) <= 8 and all(isinstance(f, PlanarFlow) and tuple(f.w.shape) == (1, 3) for f in flows)


def fused_planar_stack(flows, x):
    """phi(x) for a Planar stack through the fused CUDA kernels (x: (N,3) or (B,N,3) CUDA tensor)."""
    from .. import ops
    params = []
    for f in flows:
        params += [f.u, f.w, f.b]
    return ops.PlanarFlowStackFn.apply(x, len(flows), *params)

