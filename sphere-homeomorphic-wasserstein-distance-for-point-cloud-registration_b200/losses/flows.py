"""The learned sphere map phi (caller of the hot path): compact eager-torch restatement of the two
flows ``Norm_Flow_structure`` builds from the vendored normflows 1.7.2 copy
(``Point_Cloud_Resistration/losses/normflows_ishikawa``): ``flows.Planar`` (flows/planar.py:8-64) and
``flows.Residual`` over ``nets.LipschitzMLP`` (flows/residual.py:12-124, nets/lipschitz.py:14-68,132-293,642-648).
Only the forward map is kept -- the reference discards the log-determinant (``x, _ = flow(x)``,
s2_wasserstein.py:160-163).  This stays PyTorch: it is a tiny per-point MLP upstream of the CUDA kernels.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn


class PlanarFlow(nn.Module):
    """f(z) = z + u^ tanh(<w, z> + b), u^ = u + (softplus(<w,u>) - 1 - <w,u>) w / |w|^2  (flows/planar.py:49-60)."""

    def __init__(self, dim=3):
        super().__init__()
        self.u = nn.Parameter(torch.empty(1, dim).uniform_(-math.sqrt(2), math.sqrt(2)))
        lim_w = math.sqrt(2.0 / dim)
        self.w = nn.Parameter(torch.empty(1, dim).uniform_(-lim_w, lim_w))
        self.b = nn.Parameter(torch.zeros(1))

    def forward(self, z):
        lin = torch.sum(self.w * z, dim=-1, keepdim=True) + self.b
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
    the reference never calls update_lipschitz, so u and v stay fixed)."""

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
        self.register_buffer("u", u)
        self.register_buffer("v", v)

    def forward(self, x):
        sigma = torch.dot(self.u, torch.mv(self.weight, self.v))
        factor = torch.clamp(sigma / self.coeff, min=1.0)
        return F.linear(x, self.weight / factor, self.bias)


class ResidualFlow(nn.Module):
    """x + LipschitzMLP(x): [Swish, SpectralLinear] per layer, last layer initialised near zero (nets/lipschitz.py:47-63)."""

    def __init__(self, dim=3, hidden_units=8, hidden_layers=7, lipschitz_const=0.95):
        super().__init__()
        channels = [dim] + [hidden_units] * (hidden_layers - 1) + [dim]
        layers = []
        for i in range(len(channels) - 1):
            layers += [Swish(), SpectralLinear(channels[i], channels[i + 1], lipschitz_const,
                                               zero_init=(i == len(channels) - 2))]
        self.net = nn.Sequential(*layers)

    def forward(self, x):
        return x + self.net(x)
