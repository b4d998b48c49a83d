"""Oracle: Chamfer distance with pytorch3d semantics.  Test infrastructure only.

PARITY UNPINNED: ``pytorch3d`` is a third-party dependency of the reference (un-vendored, version un-pinned, not
installed here).  This restates the published semantics of ``pytorch3d.loss.chamfer_distance`` (squared L2, K=1
nearest neighbour in each direction, ``point_reduction`` in {"mean","sum"}, ``batch_reduction`` in {"mean","sum",
None}, returns ``(loss, None)`` when no normals are given), anchored on the reference's call sites
``Point_Cloud_Resistration/train_CD.py:123,161`` and
``Comparison_Wasserstein_with_Chamfer_distance/main_rotation.py:203``.
"""
import torch


def chamfer_distance(x, y, batch_reduction="mean", point_reduction="mean"):
    if x.dim() == 2:
        x = x.unsqueeze(0)
        y = y.unsqueeze(0)
    B, N, _ = x.shape
    M = y.shape[1]
    d = ((x.unsqueeze(2) - y.unsqueeze(1)) ** 2).sum(-1)  # (B,N,M)
    cham_x = d.min(dim=2).values  # (B,N)
    cham_y = d.min(dim=1).values  # (B,M)
    cham_x = cham_x.sum(1)
    cham_y = cham_y.sum(1)
    if point_reduction == "mean":
        cham_x = cham_x / N
        cham_y = cham_y / M
    if batch_reduction is not None:
        cham_x = cham_x.sum()
        cham_y = cham_y.sum()
        if batch_reduction == "mean":
            cham_x = cham_x / max(B, 1)
            cham_y = cham_y / max(B, 1)
    return cham_x + cham_y, None
