"""Sliced losses with the reference's function names and signatures.

  sliced_wasserstein_sphere / sliced_cost / emd1D_circle   Point_Cloud_Resistration/losses/max_spherical_sliced_w.py:210-310
  sliced_wasserstein_distance                               Wasserstein_flow_problem/Flow_ellipsoid.ipynb:203-220 (cell 5)
  transform_to_sphere / max_spherical_wassersten_distance   max_spherical_sliced_w.py:334-350, 498-536 (SURVEY.md 8f #4 tail)

The frames / directions are drawn with torch exactly as the reference draws them (``qr(randn(P,d,2))``,
row-normalised ``randn(P,d)``); projection, sort and the 1-D reductions run in CUDA.
"""
import math

import torch
from torch import nn

from .. import ops


def emd1D_circle(u_values, v_values, u_weights=None, v_weights=None, p=1, require_sort=True):
    """Circular W1 (level median) per row: (S,n),(S,m) -> (S,), p == 1 (max_spherical_sliced_w.py:210-247).  With
    ``u_weights`` / ``v_weights`` ((n,) or (S,n); the reference gathers them through the sort permutations, :224-228) the
    level median runs as the reference's four-sort composition on the device sort instead of the fused uniform kernel."""
    if p != 1:
        raise NotImplementedError("emd1D_circle implements p == 1 only, like the reference (it returns None otherwise)")
    weighted = u_weights is not None or v_weights is not None
    pu = pv = None
    if require_sort:
        u_values, pu = ops.SegmentedSortFn.apply(u_values.contiguous().float())
        v_values, pv = ops.SegmentedSortFn.apply(v_values.contiguous().float())
    if weighted:
        def sorted_w(w, perm):
            if w is None:
                return None
            w = w.to(u_values.device)
            return w[..., perm] if perm is not None else w  # (:227-228)
        return ops.circular_w1_large(u_values.contiguous(), v_values.contiguous(), sorted_w(u_weights, pu), sorted_w(v_weights, pv))
    if u_values.shape[-1] + v_values.shape[-1] > ops.CIRCULAR_W1_MAX:
        return ops.circular_w1_large(u_values.contiguous(), v_values.contiguous())
    return ops.CircularW1Fn.apply(u_values.contiguous(), v_values.contiguous())


def binary_search_circle(u_values, v_values, u_weights=None, v_weights=None, p=1, Lm=10, Lp=10, tm=-1, tp=1, eps=1e-6,
                         require_sort=True):
    """Circular W_p^p per row by bisection on the rotation: (S,n),(S,m) -> (S,)  (max_spherical_sliced_w.py:117-207).
    Every round of the reference's host-synchronised loop runs inside one kernel launch.  With ``u_weights`` / ``v_weights``
    ((n,) or (S,n); gathered through the sort permutations, :166-170) the kernel searches per-slice CDF tables instead of the
    closed-form uniform CDFs; the weights receive gradients through the final cost, like the reference's autograd."""
    pu = pv = None
    if require_sort:
        u_values, pu = ops.SegmentedSortFn.apply(u_values.contiguous().float())
        v_values, pv = ops.SegmentedSortFn.apply(v_values.contiguous().float())
    if u_weights is not None or v_weights is not None:
        S, n = u_values.shape
        m = v_values.shape[-1]
        ucdf = ops.weight_cdf(u_weights, pu, S, n, u_values.device)
        vcdf = ops.weight_cdf(v_weights, pv, S, m, u_values.device)
        w, _ = ops.CircularWpWeightedFn.apply(u_values.contiguous(), v_values.contiguous(), ucdf, vcdf, float(p), float(tm),
                                              float(tp), eps / max(Lm, Lp))
        return w
    w, _ = ops.CircularWpFn.apply(u_values.contiguous(), v_values.contiguous(), float(p), float(tm), float(tp),
                                  eps / max(Lm, Lp))
    return w


def sliced_cost(Xs, Xt, Us, p=2, u_weights=None, v_weights=None):
    """(n,3),(m,3) clouds and (P,3,2) frames -> mean over slices of the circular W (max_spherical_sliced_w.py:251-286)."""
    if u_weights is not None or v_weights is not None:
        xs, unb = ops._as_cloud(Xs, "Xs")
        xt, _ = ops._as_cloud(Xt, "Xt")
        ks = ops.ProjectCircleFn.apply(xs, Us.to(device=xs.device, dtype=torch.float32))  # (B,P,n)
        kt = ops.ProjectCircleFn.apply(xt, Us.to(device=xs.device, dtype=torch.float32))
        B, P, n = ks.shape
        if p == 1:
            w = emd1D_circle(ks.reshape(B * P, n), kt.reshape(B * P, -1), u_weights, v_weights, p=1)
        else:
            w = binary_search_circle(ks.reshape(B * P, n), kt.reshape(B * P, -1), u_weights, v_weights, p=p)
        w = w.reshape(B, P).mean(1)
        return w.reshape(()) if unb else w
    w = ops.spherical_sliced_w1(Xs, Xt, Us) if p == 1 else ops.spherical_sliced_wp(Xs, Xt, Us, float(p))
    return w.reshape(()) if Xs.dim() == 2 else w


def stiefel_frames(Z):
    """The Q factor of ``torch.linalg.qr(Z)`` for a batch of (d,2) matrices, as a dozen elementwise device ops.

    ``sliced_wasserstein_sphere`` (max_spherical_sliced_w.py:307-308) draws ``randn(P,d,2)`` and runs a batched QR inside
    every loss call; on the GPU that is a cuSOLVER batched factorisation of P tiny matrices -- 15 ms at P = 512, eight
    times the whole cfg3 loss.  For two columns Q is Gram-Schmidt up to LAPACK's Householder signs
    (R11 = -sign(a11)|a1|, R22 = -sign((H1 a2)_2)|w|), reproduced here, so the frames equal torch's to rounding."""
    a1, a2 = Z[..., 0], Z[..., 1]
    n1 = torch.linalg.vector_norm(a1, dim=-1, keepdim=True)
    s1 = torch.where(a1[..., :1] >= 0, 1.0, -1.0)
    q1 = -s1 * a1 / n1
    w = a2 - (q1 * a2).sum(-1, keepdim=True) * q1
    beta = -s1 * n1
    h2 = a2[..., 1:2] - a1[..., 1:2] * ((a1 * a2).sum(-1, keepdim=True) - beta * a2[..., :1]) / (n1 * (n1 + a1[..., :1].abs()))
    s2 = torch.where(h2 >= 0, 1.0, -1.0)
    q2 = -s2 * w / torch.linalg.vector_norm(w, dim=-1, keepdim=True)
    return torch.stack((q1, q2), dim=-1)


def sliced_wasserstein_sphere(Xs, Xt, num_projections, device, u_weights=None, v_weights=None, p=2):
    """max_spherical_sliced_w.py:289-310 -- draws ``U = qr(randn(P,d,2)).Q`` on ``device`` and calls sliced_cost."""
    d = Xs.shape[-1]
    Z = torch.randn((num_projections, d, 2), device=device)
    U = stiefel_frames(Z)
    return sliced_cost(Xs, Xt, U, p=p, u_weights=u_weights, v_weights=v_weights)


def rand_projections(dim, num_projections=100, device=None):
    projections = torch.randn((num_projections, dim), device=device)
    return projections / torch.sqrt(torch.sum(projections ** 2, dim=1, keepdim=True))


def sliced_wasserstein_distance(first_samples, second_samples, num_projection=100, p=2, device="cuda", projections=None):
    """Flow_ellipsoid.ipynb:208-220.  ``projections`` (P,3) may be supplied for reproducible tests."""
    if projections is None:
        projections = rand_projections(second_samples.size(-1), num_projection, device=device)
    w = ops.euclid_sliced_w(first_samples.to(device), second_samples.to(device), projections, float(p))
    return w.reshape(()) if first_samples.dim() == 2 else w


class transform_to_sphere(nn.Module):
    """max_spherical_sliced_w.py:334-350 -- the small sphere map of the max-SSW wrapper: an MLP 3 -> 16 -> 4 -> 2 (tanh) read as
    two angles, ``theta_1 = pi (tanh(.)/2 + 1/2)``, ``theta_2 = pi tanh(.)``, mapped to a point of S^2.  Same module tree
    (``net.0``, ``net.2``, ``net.4``) as the reference, so its ``state_dict`` loads unchanged."""

    def __init__(self):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(3, 16), nn.Tanh(), nn.Linear(16, 4), nn.Tanh(), nn.Linear(4, 2))

    def forward(self, x):
        a = self.net(x)
        t1 = math.pi * (torch.tanh(a[:, :, 0]) / 2 + 0.5)
        t2 = math.pi * torch.tanh(a[:, :, 1])
        return torch.stack([torch.sin(t1) * torch.cos(t2), torch.sin(t1) * torch.sin(t2), torch.cos(t1)], dim=2)


class max_spherical_wassersten_distance(nn.Module):
    """max_spherical_sliced_w.py:498-536 -- gradient ascent of ``phi`` on the summed sliced cost of the detached clouds
    (``max_iter`` optimiser steps), then the outer sum over the batch; returns ``(ssw, phi(first), phi(second))``.

    ``SSW`` is called the reference's way, once per pair: ``SSW(first_t[i], second_t[i], num_projections, device, p=p)``
    (any callable); every such call is one fused project + sort + reduce launch sequence (``ops.SlicedLossFn``).  When
    ``SSW`` is this package's ``sliced_wasserstein_sphere`` the B independent frame draws are made as one (B,P,d,2) tensor
    and the batch goes through ONE fused call with per-pair frames (same distribution).  ``shared_frames=True`` (an extension, off by default) draws ONE set of frames per evaluation and
    sends the whole batch through a single fused call instead of B calls -- same expectation per pair, but the pairs then
    share their slices, which the reference's per-pair draws (:307-308) do not."""

    def __init__(self, num_projections, phi, SSW, phi_op, p=2, max_iter=10, device="cuda", shared_frames=False, verbose=True):
        super().__init__()
        self.num_projections = num_projections
        self.phi = phi
        self.SSW = SSW
        self.phi_op = phi_op
        self.p = p
        self.max_iter = max_iter
        self.device = device
        self.shared_frames = shared_frames
        self.verbose = verbose

    def _sum_over_pairs(self, a, b):
        if self.shared_frames:
            Z = torch.randn((self.num_projections, a.shape[-1], 2), device=a.device)
            return sliced_cost(a, b, stiefel_frames(Z), p=self.p).sum()
        if self.SSW is sliced_wasserstein_sphere and a.dim() == 3 and a.shape[1] + b.shape[1] <= ops.CIRCULAR_W1_MAX:
            # the reference's B calls draw B independent frame sets; drawing them as one (B,P,d,2) tensor is the same
            # distribution, and the kernels take one frame set per pair -- one fused call instead of B
            Z = torch.randn((a.shape[0], self.num_projections, a.shape[-1], 2), device=a.device)
            return sliced_cost_fast(a, b, stiefel_frames(Z), p=self.p).reshape(())
        ssw = 0
        for i in range(len(a)):
            ssw = ssw + self.SSW(a[i], b[i], self.num_projections, self.device, p=self.p)
        return ssw

    def forward(self, first_samples, second_samples, train_or_test="train"):
        if train_or_test == "train":
            f0, s0 = first_samples.detach(), second_samples.detach()
            for _ in range(self.max_iter):
                ssw = self._sum_over_pairs(self.phi(f0), self.phi(s0))
                loss = -ssw  # gradient ascent
                self.phi_op.zero_grad()
                loss.backward(retain_graph=True)
                self.phi_op.step()
                if self.verbose:
                    print(ssw.item())  # (:524)
        first_t = self.phi(first_samples)
        second_t = self.phi(second_samples)
        return self._sum_over_pairs(first_t, second_t), first_t, second_t


def sliced_cost_fast(Xs, Xt, Us, p=2, u_weights=None, v_weights=None):
    """max_spherical_sliced_w_fast.py:258-295 (there named ``sliced_cost``) -- batched clouds (B,n,3), (B,m,3) with PER-PAIR
    frames ``Us`` (B,P,3,2); returns the SUM over the pairs of the per-pair mean over slices, shape (1,), as the reference's
    loop ``w1 += mean(binary_search_circle(Xps[i], Xpt[i]))`` does.  One fused sliced call for the whole batch (per pair when weights are given).  (The reference's
    p == 1 branch fails on batched keys -- its ``gather`` index is flattened, :250 -- the same sum is returned here.)"""
    if Us.dim() != 4 or Us.shape[0] != Xs.shape[0]:
        raise ValueError("Us must be (batch, num_projections, d, 2) with one frame set per pair")
    if u_weights is None and v_weights is None and Xs.shape[1] + Xt.shape[1] <= ops.CIRCULAR_W1_MAX:
        # the whole batch in ONE fused call: the kernels index the frames by pair (shwd_*_pp, include/shwd.h)
        w = ops.spherical_sliced_w1(Xs, Xt, Us) if p == 1 else ops.spherical_sliced_wp(Xs, Xt, Us, float(p))
        return w.sum().reshape(1)
    w1 = torch.zeros((1,), device=Xs.device)
    for i in range(Xs.shape[0]):
        w1 = w1 + sliced_cost(Xs[i], Xt[i], Us[i], p=p, u_weights=u_weights, v_weights=v_weights)
    return w1


def sliced_wasserstein_sphere_fast(Xs, Xt, num_projections, device, u_weights=None, v_weights=None, p=2):
    """max_spherical_sliced_w_fast.py:298-319 -- one frame draw ``qr(randn(B,P,d,2))`` for the whole batch, then sliced_cost_fast."""
    Z = torch.randn((Xs.shape[0], num_projections, Xs.shape[2], 2), device=device)
    return sliced_cost_fast(Xs, Xt, stiefel_frames(Z), p=p, u_weights=u_weights, v_weights=v_weights)


class transform_to_sphere_fast(transform_to_sphere):
    """max_spherical_sliced_w_fast.py:323-340 -- the same module as transform_to_sphere."""


class max_spherical_wassersten_distance_fast(max_spherical_wassersten_distance):
    """max_spherical_sliced_w_fast.py:346-382 -- as max_spherical_wassersten_distance, with ``SSW`` called ONCE per evaluation on
    the whole batch: ``SSW(phi(first), phi(second), num_projections, device, p=p)`` (pass sliced_wasserstein_sphere_fast)."""

    def _sum_over_pairs(self, a, b):
        return self.SSW(a, b, self.num_projections, self.device, p=self.p)
