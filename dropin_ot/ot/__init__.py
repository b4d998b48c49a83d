"""``import ot`` drop-in for the ONE function of POT the reference's loss path calls: ``ot.emd2(a, b, M)``
(Point_Cloud_Resistration/losses/s2_wasserstein.py:42,47,102,107; Comparison_Wasserstein_with_Chamfer_distance/
main_rotation.py:213 via POT_loss :63-79; the W2 metric of the notebooks, Flow_ellipsoid.ipynb cell 8).

Put this file's parent directory on ``sys.path`` (``PYTHONPATH=<repo>/dropin_ot``) on a machine without POT, or to keep
the exact solve on the GPU: uniform weights and a square cost matrix are solved by the float64 auction kernel of
libshwd_b200.so (``shwd_exact_assignment_dense``); the value is POT's ((1/n) * optimal assignment cost, float64
accumulation, returned in M's dtype) and the gradient w.r.t. M is the optimal plan, as POT's torch backend attaches it.
Anything else POT offers is not provided: non-uniform weights or a rectangular M raise ``NotImplementedError``.
POT is un-vendored and version-unpinned in the reference: parity is against the LP optimum
(scipy.optimize.linear_sum_assignment in the tests), not against a pinned POT build.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import torch  # noqa: E402

import shwd  # noqa: E402,F401

__all__ = ["emd2"]


def _uniform(w, n):
    if w is None:
        return True
    if not isinstance(w, torch.Tensor):
        w = torch.as_tensor(w)
    if w.numel() == 0:  # POT: an empty weight vector means uniform
        return True
    return w.numel() == n and bool(torch.allclose(w.detach().float().cpu(), torch.full((n,), 1.0 / n), rtol=1e-4, atol=1e-7))


def emd2(a, b, M, processes=1, numItermax=100000, log=False, return_matrix=False, center_dual=True, numThreads=1):
    """Exact OT cost <G*, M> for uniform marginals ``a`` (n,), ``b`` (m,) and the cost matrix ``M`` (n,m) on a CUDA device."""
    if log or return_matrix:
        raise NotImplementedError("ot.emd2(log=True / return_matrix=True) is not used on the reference's path")
    if not isinstance(M, torch.Tensor):
        raise TypeError("this ot.emd2 takes torch tensors (the reference passes torch tensors)")
    if M.dim() != 2:
        raise NotImplementedError("ot.emd2 takes one (n, m) cost matrix, got %s" % (tuple(M.shape),))
    n, m = M.shape
    if not (_uniform(a, n) and _uniform(b, m)):
        raise NotImplementedError("only uniform weights are solved on the GPU (what the reference passes)")
    # n != m: the assignment of lcm(n, m) copies (raises NotImplementedError beyond the kernel's size)
    return shwd.exact_emd2_dense(M)
