"""The mini-batch max-SSW variant, ``Point_Cloud_Resistration/losses/mini_batch_Residual_MSSW.py`` (a caller of the sliced path,
SURVEY.md 8f #4 tail): the sphere map  R^3 --MLP--> R^2 --flows--> R^2 --angles--> S^2  (:327-408) and the wrapper whose ascent
steps use a random mini-batch of the pairs (:413-452).  Host orchestration in the reference's order on top of the fused sliced
kernels; the modules keep the reference's module tree, so its ``state_dict`` loads unchanged.
"""
import math

import numpy as np
import torch
from torch import nn

from .flows import PlanarFlow, ResidualFlow
from .sliced import sliced_cost_fast, sliced_wasserstein_sphere, stiefel_frames
from .. import ops


class MLP_Architecture(nn.Module):
    """mini_batch_Residual_MSSW.py:327-355 -- point-wise MLP 3 -> 8 -> 8 -> emb_dims as 1x1 convolutions with ReLU.  (The
    convolutions are registered both as attributes and inside ``net``, as in the reference: both key sets exist.)"""

    def __init__(self, emb_dims=2):
        super().__init__()
        self.emb_dims = emb_dims
        self.conv1 = nn.Conv1d(3, 8, 1)
        self.conv2 = nn.Conv1d(8, 8, 1)
        self.conv3 = nn.Conv1d(8, self.emb_dims, 1)
        self.relu = nn.ReLU()
        self.layers = [self.conv1, self.relu, self.conv2, self.relu, self.conv3]
        self.net = nn.Sequential(*self.layers)

    def forward(self, input_data):
        return self.net(input_data.permute(0, 2, 1)).permute(0, 2, 1)


class ActNorm(nn.Module):
    """z * exp(s) + t with the data-dependent initialisation of the first batch (vendored normflows, flows/normalization.py:7-29
    over flows/affine/coupling.py:9-45): ``s = -log(std(z, dim 0) + 1e-6)``, ``t = -mean(z, dim 0) exp(s)`` -- statistics over
    dim 0 only (the dims where the (1, d) parameter has extent 1), so on (B, N, d) input the parameters take shape (1, N, d), as
    in the reference."""

    def __init__(self, dim):
        super().__init__()
        self.s = nn.Parameter(torch.zeros(dim)[None])
        self.t = nn.Parameter(torch.zeros(dim)[None])
        self.batch_dims = [0]
        self.register_buffer("data_dep_init_done", torch.tensor(0.0))

    def forward(self, z):
        if not self.data_dep_init_done > 0.0:
            s_init = -torch.log(z.std(dim=self.batch_dims, keepdim=True) + 1e-6)
            self.s.data = s_init.data
            self.t.data = (-z.mean(dim=self.batch_dims, keepdim=True) * torch.exp(self.s)).data
            self.data_dep_init_done = torch.tensor(1.0, device=z.device)
        return z * torch.exp(self.s) + self.t


class Flow_structure(nn.Module):
    """mini_batch_Residual_MSSW.py:359-390 -- flows on R^2: Planar x n, or [Residual (2 -> 4 -> 2, Lipschitz 0.9), ActNorm] x n.
    ``nf.flows.Residual(net, reduce_memory=True)`` keeps normflows' default ``reverse=True``: its forward is the fixed-point
    INVERSE of x + g(x) (flows/residual.py:63-65,133-142), reproduced by ``ResidualFlow(reverse=True)``."""

    def __init__(self, input_dim=2, flow_name="Planar", n_flow_layer=3):
        super().__init__()
        if flow_name == "Planar":
            flows = [PlanarFlow(input_dim) for _ in range(n_flow_layer)]
        elif flow_name == "Residual":
            flows = []
            for _ in range(n_flow_layer):
                flows += [ResidualFlow(input_dim, hidden_units=4, hidden_layers=2, lipschitz_const=0.9, reverse=True), ActNorm(input_dim)]
        else:
            raise ValueError("Flow name is not valid")
        self.net = nn.ModuleList(flows)

    def forward(self, x):
        for flow in self.net:
            x = flow(x)
        return x


class transform_to_sphere(nn.Module):
    """mini_batch_Residual_MSSW.py:392-408."""

    def __init__(self, flow_name, n_flow_layer=3, two_d_encoder=MLP_Architecture, flow=Flow_structure):
        super().__init__()
        self.net = two_d_encoder()
        self.flows = flow(flow_name=flow_name, n_flow_layer=n_flow_layer)

    def forward(self, x):
        x = self.flows(self.net(x).contiguous())
        t1 = math.pi * (torch.tanh(x[:, :, 0]) / 2 + 0.5)
        t2 = math.pi * torch.tanh(x[:, :, 1])
        return torch.stack([torch.sin(t1) * torch.cos(t2), torch.sin(t1) * torch.sin(t2), torch.cos(t1)], dim=2)


class max_spherical_wassersten_distance_Residual(nn.Module):
    """mini_batch_Residual_MSSW.py:413-452 -- every ascent step sums the sliced cost over a random mini-batch of
    ``psi_minibatch_size`` pairs (``np.random.choice(B, size, replace=False)``, the reference's draw), the outer value over all
    pairs.  With this package's ``sliced_wasserstein_sphere`` as ``SSW`` the selected pairs go through ONE fused call with
    per-pair frames (same distribution as the reference's per-pair draws); any other callable is called per pair."""

    def __init__(self, num_projections, phi, phi_op, SSW=sliced_wasserstein_sphere, p=2, max_iter=10, psi_minibatch_size=5,
                 device="cuda", verbose=True):
        super().__init__()
        self.num_projections = num_projections
        self.phi = phi
        self.SSW = SSW
        self.phi_op = phi_op
        self.p = p
        self.max_iter = max_iter
        self.psi_minibatch_size = psi_minibatch_size
        self.device = device
        self.verbose = verbose

    def _sum(self, a, b, idx):
        if self.SSW is sliced_wasserstein_sphere and a.shape[1] + b.shape[1] <= ops.CIRCULAR_W1_MAX:
            sel = torch.as_tensor(np.asarray(idx), device=a.device, dtype=torch.long)
            Z = torch.randn((len(sel), self.num_projections, a.shape[-1], 2), device=a.device)
            return sliced_cost_fast(a[sel], b[sel], stiefel_frames(Z), p=self.p).reshape(())
        ssw = 0
        for i in idx:
            ssw = ssw + self.SSW(a[i], b[i], self.num_projections, self.device, p=self.p)
        return ssw

    def forward(self, first_samples, second_samples, train_or_test="train"):
        if train_or_test == "train":
            f0, s0 = first_samples.detach(), second_samples.detach()
            for _ in range(self.max_iter):
                ft, st = self.phi(f0), self.phi(s0)
                mini_batch = np.random.choice(len(ft), size=self.psi_minibatch_size, replace=False)
                ssw = self._sum(ft, st, mini_batch)
                loss = -ssw  # gradient ascent
                self.phi_op.zero_grad()
                loss.backward(retain_graph=True)
                self.phi_op.step()
                if self.verbose:
                    print(ssw.item())
        first_t = self.phi(first_samples)
        second_t = self.phi(second_samples)
        return self._sum(first_t, second_t, range(len(first_t))), first_t, second_t
