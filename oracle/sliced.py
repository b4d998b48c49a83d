"""Oracle: sliced paths (spherical sliced W on great circles, Euclidean sliced W).  Test infrastructure only.

All citations are to ``Point_Cloud_Resistration/losses/max_spherical_sliced_w.py`` unless noted.
"""
import math
import torch
import torch.nn.functional as F


def project_circle(X, U):
    """(n,3) points, (P,3,2) Stiefel frames -> (P,n) circle coordinates in [0,1].

    ``sliced_cost`` :270-279: ``Xp[p,n,k] = sum_d U[p,d,k] X[n,d]``, L2-normalise over k,
    ``t = (atan2(-Xp1, -Xp0) + pi) / (2 pi)``.
    """
    Xp = torch.einsum("pdk,nd->pnk", U, X)
    Xp = F.normalize(Xp, p=2, dim=-1)
    return (torch.atan2(-Xp[:, :, 1], -Xp[:, :, 0]) + math.pi) / (2 * math.pi)


def emd1d_circle(u_values, v_values, stable=False, u_weights=None, v_weights=None):
    """Circular W1 by the level-median formula -- ``emd1D_circle`` :210-247 (p == 1 branch); uniform weights unless
    ``u_weights`` (n,) / (P,n) and ``v_weights`` are given (:217-228: gathered through the sort permutations).

    (P,n),(P,m) -> (P,).  Quirk kept: ``delta`` pads the sorted merged values with 1 at the END only, so the arc
    from 0 to the first point is omitted (:238-239).
    ``stable`` selects ``torch.sort(stable=True)``; the reference calls plain ``torch.sort`` (equal for tie-free keys).
    """
    n, m = u_values.shape[-1], v_values.shape[-1]
    dt = u_values.dtype
    uw = torch.full((n,), 1 / n, dtype=dt) if u_weights is None else u_weights
    vw = torch.full((m,), 1 / m, dtype=dt) if v_weights is None else v_weights
    u_sorted, u_perm = torch.sort(u_values, dim=-1, stable=stable)
    v_sorted, v_perm = torch.sort(v_values, dim=-1, stable=stable)
    uw = uw[..., u_perm]
    vw = vw[..., v_perm]
    merged, merged_perm = torch.sort(torch.cat((u_sorted, v_sorted), -1), dim=-1, stable=stable)
    cdf_diff = torch.cumsum(torch.gather(torch.cat((uw, -vw), -1), -1, merged_perm), -1)
    cdf_sorted, cdf_perm = torch.sort(cdf_diff, dim=-1, stable=stable)
    merged_pad = F.pad(merged, (0, 1), value=1)
    delta = merged_pad[..., 1:] - merged_pad[..., :-1]
    w_sorted = torch.gather(delta, -1, cdf_perm)
    cw = torch.cumsum(w_sorted, dim=-1) - 0.5
    cw[cw < 0] = float("inf")
    k = torch.argmin(cw, dim=-1)
    lev_med = torch.gather(cdf_sorted, -1, k.view(-1, 1))
    return torch.sum(delta * torch.abs(cdf_diff - lev_med), dim=-1)


def sliced_wasserstein_sphere_p1(Xs, Xt, U, stable=False):
    """``sliced_cost`` :251-286 with p == 1 and an explicit frame tensor ``U`` (the reference draws
    ``U = qr(randn(P,3,2)).Q`` inside ``sliced_wasserstein_sphere`` :307-308).  Returns ``mean_P W1``."""
    return torch.mean(emd1d_circle(project_circle(Xs, U), project_circle(Xt, U), stable=stable))


def euclid_sliced_wasserstein(x, y, theta, p=2):
    """Euclidean sliced W -- ``Wasserstein_flow_problem/Flow_ellipsoid.ipynb:208-220`` (cell 5
    ``sliced_wasserstein_distance``) with explicit unit directions ``theta`` (P,3) (``rand_projections`` :203-206).
    ``(mean_P sum_n |sort(x theta)_n - sort(y theta)_n|^p)^(1/p)``; needs n == m."""
    xp = x.matmul(theta.transpose(0, 1)).transpose(0, 1)
    yp = y.matmul(theta.transpose(0, 1)).transpose(0, 1)
    d = torch.abs(torch.sort(xp, dim=1)[0] - torch.sort(yp, dim=1)[0])
    w = torch.pow(torch.sum(torch.pow(d, p), dim=1), 1.0 / p)
    return torch.pow(torch.pow(w, p).mean(), 1.0 / p)


# ----------------------------------------------------------------------------------------------------------------
# Circular W_p, p != 1 (Delon-Salomon-Sobolevski bisection) -- :9-207.  Behavioural restatement, quirks preserved.
# ----------------------------------------------------------------------------------------------------------------
def _roll_rows(mat, shifts):
    """Row-wise circular shift to the right by ``shifts`` (``roll_by_gather`` :9-22, dim == 1)."""
    cols = mat.shape[1]
    idx = (torch.arange(cols).view(1, cols) - shifts.view(-1, 1)) % cols
    return torch.gather(mat, 1, idx)


def _shifted(theta, v_values, v_cdf):
    """Common prelude of ``dCost`` :25-50 and ``Cost`` :68-92: shift v's CDF by frac(theta), lift the values by
    floor(theta) (+1 where the shifted CDF went negative), re-wrap negatives only when the WHOLE batch has both
    signs (:41-42 / :82-83, batch-global quirk), rotate each row so its smallest non-negative CDF entry leads,
    append ``v_0 + 1``."""
    v_values = v_values.clone()
    fl = torch.floor(theta)
    cdf = v_cdf - (theta - fl)
    neg = cdf < 0
    pos = ~neg
    v_values[neg] += fl[neg] + 1
    v_values[pos] += fl[pos]
    if bool(neg.any()) and bool(pos.any()):
        cdf[neg] += 1
    key = cdf.clone()
    key[neg] = float("inf")
    shift = -torch.argmin(key, dim=-1)
    cdf = _roll_rows(cdf, shift)
    v_values = _roll_rows(v_values, shift)
    v_values = torch.cat([v_values, v_values[:, :1] + 1], dim=1)
    return cdf, v_values


def _dcost(theta, u_values, v_values, u_cdf, v_cdf, p):
    n = u_values.shape[-1]
    cdf, vv = _shifted(theta, v_values, v_cdf)
    iu = torch.searchsorted(u_cdf, cdf)
    u_icdf = torch.gather(u_values, -1, iu.clip(0, n - 1))
    u_cdf1 = torch.cat([u_cdf, u_cdf[:, :1] + 1], dim=1)
    u_val1 = torch.cat([u_values, u_values[:, :1] + 1], dim=1)
    ium = torch.searchsorted(u_cdf1, cdf, right=True)
    u_icdfm = torch.gather(u_val1, -1, ium.clip(0, n))
    dcp = torch.sum(torch.abs(u_icdf - vv[:, 1:]) ** p - torch.abs(u_icdf - vv[:, :-1]) ** p, dim=-1)
    dcm = torch.sum(torch.abs(u_icdfm - vv[:, 1:]) ** p - torch.abs(u_icdfm - vv[:, :-1]) ** p, dim=-1)
    return dcp.reshape(-1, 1), dcm.reshape(-1, 1)


def _cost(theta, u_values, v_values, u_cdf, v_cdf, p):
    n = u_values.shape[-1]
    m = v_values.shape[-1]
    cdf, vv = _shifted(theta, v_values, v_cdf)
    axis, _ = torch.sort(torch.cat((u_cdf, cdf), -1), -1)
    axis_pad = F.pad(axis, (1, 0))
    delta = axis_pad[..., 1:] - axis_pad[..., :-1]
    iu = torch.searchsorted(u_cdf, axis)
    u_icdf = torch.gather(u_values, -1, iu.clip(0, n - 1))
    vv = torch.cat([vv, vv[:, :1] + 1], dim=1)  # second append, :103
    iv = torch.searchsorted(cdf, axis)
    v_icdf = torch.gather(vv, -1, iv.clip(0, m))
    return torch.sum(delta * torch.abs(u_icdf - v_icdf) ** p, dim=-1)


def binary_search_circle(u_values, v_values, p=2, Lm=10, Lp=10, tm=-1.0, tp=1.0, eps=1e-6, return_theta=False,
                         u_weights=None, v_weights=None):
    """``binary_search_circle`` :117-207 with ``require_sort=True``; uniform weights unless ``u_weights`` (n,) /
    ``v_weights`` (m,) are given (:156-170: gathered through the sort permutations, then cumsum).  (P,n),(P,m) -> (P,)."""
    n, m = u_values.shape[-1], v_values.shape[-1]
    dt = u_values.dtype
    rows = u_values.shape[0]
    uw = torch.full((n,), 1 / n, dtype=dt) if u_weights is None else u_weights
    vw = torch.full((m,), 1 / m, dtype=dt) if v_weights is None else v_weights
    u_values, u_perm = torch.sort(u_values, -1)
    v_values, v_perm = torch.sort(v_values, -1)
    u_cdf = torch.cumsum(uw[..., u_perm], -1)
    v_cdf = torch.cumsum(vw[..., v_perm], -1)
    L = max(Lm, Lp)
    tm = torch.full((rows, m), tm, dtype=dt)
    tp = torch.full((rows, m), tp, dtype=dt)
    tc = (tm + tp) / 2
    done = torch.zeros((rows, m))
    while bool(torch.any(1 - done)):
        dcp, dcm = _dcost(tc, u_values, v_values, u_cdf, v_cdf, p)
        done = ((dcp * dcm) <= 0) * 1
        mask = ((tp - tm) < eps / L) * (1 - done)
        if bool(torch.any(mask)):
            dcptp, dcmtp = _dcost(tp, u_values, v_values, u_cdf, v_cdf, p)
            dcptm, dcmtm = _dcost(tm, u_values, v_values, u_cdf, v_cdf, p)
            ctm = _cost(tm, u_values, v_values, u_cdf, v_cdf, p).reshape(-1, 1)
            ctp = _cost(tp, u_values, v_values, u_cdf, v_cdf, p).reshape(-1, 1)
            mask_end = mask * (torch.abs(dcptm - dcmtp) > 0.001)
            tc[mask_end > 0] = ((ctp - ctm + tm * dcptm - tp * dcmtp) / (dcptm - dcmtp))[mask_end > 0]
            done[torch.prod(mask, dim=-1) > 0] = 1
        elif bool(torch.any(1 - done)):
            lo = ((1 - mask) * (dcp < 0)) > 0
            hi = ((1 - mask) * (dcp >= 0)) > 0
            tm[lo] = tc[lo]
            tp[hi] = tc[hi]
            go = ((1 - mask) * (1 - done)) > 0
            tc[go] = (tm[go] + tp[go]) / 2
    w = _cost(tc.detach(), u_values, v_values, u_cdf, v_cdf, p)
    return (w, tc[:, 0].detach()) if return_theta else w


def sliced_wasserstein_sphere(Xs, Xt, U, p=2, u_weights=None, v_weights=None):
    """``sliced_cost`` :251-286 for any p with an explicit frame tensor ``U``: mean over slices of W1 (p == 1) or
    of W_p^p (p != 1; no root is taken, :284-286)."""
    a = project_circle(Xs, U)
    b = project_circle(Xt, U)
    if p == 1:
        w = emd1d_circle(a, b, u_weights=u_weights, v_weights=v_weights)
    else:
        w = binary_search_circle(a, b, p=p, u_weights=u_weights, v_weights=v_weights)
    return torch.mean(w)
