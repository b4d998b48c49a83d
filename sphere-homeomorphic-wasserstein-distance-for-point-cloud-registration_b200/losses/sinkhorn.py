"""Entropic OT losses with the reference's class names and signatures; the recurrence runs in the CUDA sweeps.

  Sinkhorn_Distance_Loss            Point_Cloud_Resistration/losses/Sinkhorn.py:3-92       (no early stop, sum|d|^p)
  log_Sinkhorn_Distance_Loss        Comparison_Wasserstein_with_Chamfer_distance/losses/sinkhorn.py:3-86
                                                                                            (early stop 1e-9, sum|d|^p)
  log_Sinkhorn_Distance_Loss_fixed  Point_Cloud_Resistration/losses/Sinkhorn_fixed.py:10-93 (early stop, (sum|d|^p)^(1/p))
  log_N_Sinkhorn_Distance_Loss      Comparison_.../losses/sinkhorn.py:92-184                (cost C^N, result cost^(1/N))

``forward(x, y, device) -> (cost, P, C)`` as in the reference.  P and C are (B,N,M) tensors, which this path never
needs: ``dense_outputs`` chooses whether they are materialised -- "auto" (default) only while B*N*M <= 2**22,
True always, False never (then None is returned in their place; the reference's callers ignore both,
main_rotation.py:210).
"""
import torch

from .. import ops

_AUTO_DENSE_LIMIT = 1 << 22


class _SinkhornBase(torch.nn.Module):
    _kind = "sqeuclid"
    _thresh = 0.0

    def __init__(self, eps, max_iter, batch_reduction="none", type_of_cost_norm="L2", dense_outputs="auto"):
        super().__init__()
        self.eps = eps
        self.max_iter = max_iter
        self.batch_reduction = batch_reduction
        self.p = self._type_of_cost_norm(type_of_cost_norm)
        self.dense_outputs = dense_outputs
        self.n_power = 1

    @staticmethod
    def _type_of_cost_norm(type_of_cost_norm="L2"):
        return int(type_of_cost_norm[-1])

    def forward(self, x, y, device=None):
        if device is not None:
            x = x.to(device)
            y = y.to(device)
        unbatched = x.dim() == 2
        res = ops.entropic_ot(x, y, self._kind, float(self.p), float(self.eps), int(self.max_iter), float(self.n_power),
                              float(self._thresh))
        cost = res.cost
        if self.n_power != 1:
            cost = torch.pow(cost, 1 / self.n_power)
        if unbatched:
            cost = cost.reshape(())
        if self.batch_reduction == "mean":
            cost = cost.mean()
        elif self.batch_reduction == "sum":
            cost = cost.sum()
        P = C = None
        B = 1 if unbatched else x.shape[0]
        dense = self.dense_outputs
        if dense == "auto":
            dense = B * x.shape[-2] * y.shape[-2] <= _AUTO_DENSE_LIMIT
        if dense:
            P, C = res.dense()
            if unbatched:
                P, C = P[0], C[0]
        return cost, P, C


class Sinkhorn_Distance_Loss(_SinkhornBase):
    """losses/Sinkhorn.py:3-92 -- fixed ``max_iter`` iterations, cost ``sum_k |x_k - y_k|^p``."""

    @staticmethod
    def _type_of_cost_norm(type_of_cost_norm="L2"):
        if type_of_cost_norm == "L2":
            return 2
        if type_of_cost_norm == "L1":
            return 1
        raise ValueError("type_of_cost_norm must be 'L1' or 'L2'")


class log_Sinkhorn_Distance_Loss(_SinkhornBase):
    """Comparison_.../losses/sinkhorn.py:3-86 -- early stop on mean_b sum_i |du| < 1e-9, cost ``sum_k |x_k - y_k|^p``."""
    _thresh = 1e-9


class log_Sinkhorn_Distance_Loss_fixed(_SinkhornBase):
    """losses/Sinkhorn_fixed.py:10-93 -- early stop, cost ``(sum_k |x_k - y_k|^p)^(1/p)``."""
    _kind = "euclid"
    _thresh = 1e-9


class log_N_Sinkhorn_Distance_Loss(_SinkhornBase):
    """Comparison_.../losses/sinkhorn.py:92-184 -- cost matrix ``C^N``, returns ``cost^(1/N)``."""
    _thresh = 1e-9

    def __init__(self, eps, max_iter, batch_reduction="none", type_of_cost_norm="L2", type_of_Wasserstein_N="2",
                 dense_outputs="auto"):
        super().__init__(eps, max_iter, batch_reduction, type_of_cost_norm, dense_outputs)
        self.N = int(type_of_Wasserstein_N)
        self.n_power = self.N
