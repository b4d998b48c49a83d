"""Oracle: the reference's log-domain Sinkhorn recurrence.  Test infrastructure only."""
import torch
from .costs import cost_matrix


def log_sinkhorn(x, y, kind="sqeuclid", p=2, eps=0.01, max_iter=100, thresh=None, reduction="none", n_power=1,
                 return_plan=False):
    """Restates ``log_Sinkhorn_Distance_Loss.forward``
    (``Comparison_Wasserstein_with_Chamfer_distance/losses/sinkhorn.py:14-60``; identical recurrences at
    ``Point_Cloud_Resistration/losses/Sinkhorn.py:14-60`` -- no early stop -- and ``losses/Sinkhorn_fixed.py:21-67``)
    with the cost matrix pluggable (``kind``), so the same recurrence can run on the geodesic cost of
    ``s2_wasserstein.py:119-122``.

    ``thresh``: None -> fixed ``max_iter`` iterations (Sinkhorn.py, lines 42-44 commented out there);
    a float (1e-9 in sinkhorn.py:32 / Sinkhorn_fixed.py:39) -> break when the batch-mean L1 change of u drops below
    it, tested after the v-update of that iteration (sinkhorn.py:42-44).
    ``n_power``: ``log_N_Sinkhorn_Distance_Loss`` (sinkhorn.py:104-152): cost matrix ``C^N``, result ``cost^(1/N)``.
    Autograd runs through the whole unrolled loop, exactly like the reference.
    """
    C = cost_matrix(x, y, kind, p, n_power)
    n, m = x.shape[-2], y.shape[-2]
    batch = 1 if x.dim() == 2 else x.shape[0]
    dt = C.dtype
    # sinkhorn.py:25-26 -- float32 fill of 1.0/n, .squeeze()
    a = torch.empty(batch, n, dtype=torch.float).fill_(1.0 / n).squeeze().to(dtype=dt, device=C.device)
    b = torch.empty(batch, m, dtype=torch.float).fill_(1.0 / m).squeeze().to(dtype=dt, device=C.device)
    u = torch.zeros_like(a)
    v = torch.zeros_like(b)

    def M(u, v):  # sinkhorn.py:62-69
        return (-C + u.unsqueeze(-1) + v.unsqueeze(-2)) / eps

    iters_run = 0
    for _ in range(max_iter):
        u_init = u
        u = eps * (torch.log(a + 1e-8) - torch.logsumexp(M(u, v), dim=-1)) + u
        v = eps * (torch.log(b + 1e-8) - torch.logsumexp(M(u, v).transpose(-2, -1), dim=-1)) + v
        iters_run += 1
        if thresh is not None:
            err = (u - u_init).abs().sum(-1).mean()
            if err.item() < thresh:
                break
    P = torch.exp(M(u, v))
    cost = torch.sum(P * C, dim=(-2, -1))
    if n_power != 1:
        cost = torch.pow(cost, 1 / n_power)
    if reduction == "mean":
        cost = cost.mean()
    elif reduction == "sum":
        cost = cost.sum()
    if return_plan:
        return cost, P, C, u, v, iters_run
    return cost


def entropic_w(x, y, kind="geodesic", p=2, eps=0.01, max_iter=100):
    """The drop-in semantics of ``Geodesic_distance_W`` / ``Cos_disimilarity_W`` when the per-pair exact solve
    ``ot.emd2(a_i, b_i, C_i) ** (1/p)`` (``s2_wasserstein.py:39-50`` / ``:99-110``) is replaced by the entropic
    solve above: ``mean_b cost_b ** (1/p)`` for batched input, ``cost ** (1/p)`` for un-batched input.
    """
    cost = log_sinkhorn(x, y, kind, p, eps, max_iter, None, "none")
    w = torch.pow(cost, 1.0 / p)
    if x.dim() == 3 and x.shape[0] >= 2:
        return w.sum() / int(x.shape[0])
    return w.reshape(())
