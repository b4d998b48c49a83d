"""``chamfer_distance`` with pytorch3d's call signature as the reference uses it
(Point_Cloud_Resistration/train_CD.py:123,161; Comparison_.../main_rotation.py:203): squared L2, K = 1."""
from .. import ops


def chamfer_distance(x, y, x_lengths=None, y_lengths=None, x_normals=None, y_normals=None, weights=None,
                     batch_reduction="mean", point_reduction="mean", norm=2, single_directional=False,
                     abs_cosine=True):
    """Returns ``(loss, None)``.  Only what the reference exercises is supported: equal-length clouds, no normals,
    no weights, ``norm == 2``; anything else raises (as pytorch3d raises on invalid arguments)."""
    if x_lengths is not None or y_lengths is not None or x_normals is not None or y_normals is not None or weights is not None:
        raise NotImplementedError("lengths / normals / weights are not used on the reference's path")
    if norm != 2:
        raise ValueError("Support for 1 or 2 norm.") if norm not in (1, 2) else NotImplementedError("norm=1 is not used on this path")
    if batch_reduction is not None and batch_reduction not in ("mean", "sum"):
        raise ValueError('batch_reduction must be one of ["mean", "sum"] or None')
    if point_reduction not in ("mean", "sum"):
        raise ValueError('point_reduction must be one of ["mean", "sum"]')
    # one autograd node: nearest neighbours + every reduction below in two launches, backward in one (csrc/chamfer.cu)
    return ops.chamfer_loss(x, y, point_reduction == "mean", batch_reduction, single_directional), None
