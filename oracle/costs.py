"""Oracle: the cost matrices of the loss path.  Test infrastructure only."""
import torch
import torch.nn.functional as F

COST_KINDS = ("geodesic", "sqeuclid", "euclid", "one_minus_cos")


def cost_matrix(x, y, kind="geodesic", p=2, n_power=1):
    """(B,N,3),(B,M,3) -> (B,N,M) (or un-batched).

    geodesic       acos(cos_sim)^p          Point_Cloud_Resistration/losses/s2_wasserstein.py:112-123
    sqeuclid       sum_k |x_k - y_k|^p      s2_wasserstein.py:52-63 (active line 62); losses/Sinkhorn.py:72-82;
                                            Comparison_Wasserstein_with_Chamfer_distance/losses/sinkhorn.py:71-82
    euclid         (sum_k |x_k-y_k|^p)^(1/p)  losses/Sinkhorn_fixed.py:79-89
    one_minus_cos  (1 - cos_sim)^p          losses/max_spherical_w_cos_with_regulation.py:745
    n_power        C^N of log_N_Sinkhorn    Comparison_.../losses/sinkhorn.py:165-176
    """
    x_col = x.unsqueeze(-2)
    y_lin = y.unsqueeze(-3)
    if kind == "geodesic":
        C = torch.acos(F.cosine_similarity(x_col, y_lin, dim=-1)) ** p
    elif kind == "sqeuclid":
        C = torch.sum(torch.abs(x_col - y_lin) ** p, -1)
    elif kind == "euclid":
        C = torch.pow(torch.sum(torch.abs(x_col - y_lin) ** p, -1), 1 / p)
    elif kind == "one_minus_cos":
        C = (1 - F.cosine_similarity(x_col, y_lin, dim=-1)) ** p
    else:
        raise ValueError(kind)
    if n_power != 1:
        C = torch.pow(C, n_power)
    return C
