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
