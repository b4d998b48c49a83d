"""Oracle: exact EMD stand-in for POT's ``ot.emd2``.  Test infrastructure only.

PARITY UNPINNED: POT is a third-party dependency of the reference (un-vendored, version un-pinned, not installed).
For uniform weights with n == m the LP optimum is attained at a permutation, so ``emd2 = (1/n) * min-assignment``;
``scipy.optimize.linear_sum_assignment`` on the float64 cost gives that optimum.  Used only to *report* the entropic
gap of the drop-in against the exact solve the reference runs at ``s2_wasserstein.py:41-43`` -- never as a parity
target of the CUDA path.
"""
import numpy as np
import torch


def exact_emd2(C):
    """C: (n,n) tensor -> (value, plan) with uniform marginals."""
    from scipy.optimize import linear_sum_assignment
    Cn = C.detach().double().cpu().numpy()
    n, m = Cn.shape
    if n != m:
        raise ValueError("exact_emd2 stand-in needs n == m (uniform weights)")
    r, c = linear_sum_assignment(Cn)
    plan = np.zeros_like(Cn)
    plan[r, c] = 1.0 / n
    return float((plan * Cn).sum()), torch.from_numpy(plan)


def exact_emd2_lp(C):
    """C: (n,m) tensor, n != m allowed -> (value, plan): the transport LP with uniform marginals 1/n, 1/m solved by
    scipy.optimize.linprog (HiGHS) on the float64 cost -- the general problem ``ot.emd2`` solves when the two clouds of
    s2_wasserstein.py:39-50 have different sizes.  Small problems only (n * m variables)."""
    from scipy.optimize import linprog
    Cn = C.detach().double().cpu().numpy()
    n, m = Cn.shape
    A = np.zeros((n + m, n * m))
    for i in range(n):
        A[i, i * m:(i + 1) * m] = 1.0
    for j in range(m):
        A[n + j, j::m] = 1.0
    rhs = np.concatenate([np.full(n, 1.0 / n), np.full(m, 1.0 / m)])
    res = linprog(Cn.reshape(-1), A_eq=A[:-1], b_eq=rhs[:-1], bounds=(0, None), method="highs")  # (one constraint is redundant)
    if res.status != 0:
        raise RuntimeError("linprog failed: %s" % res.message)
    return float(res.fun), torch.from_numpy(res.x.reshape(n, m))
