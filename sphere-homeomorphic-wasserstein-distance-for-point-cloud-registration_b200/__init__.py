"""B200-native sphere-homeomorphic Wasserstein loss path (sm_100a CUDA behind a C ABI, ctypes-bound).

The directory name carries hyphens (it mirrors the upstream repository name), so import it through the ``shwd``
alias module at the repository root, or put this directory on ``sys.path`` and ``import losses`` for the drop-in
replacement of the reference's ``losses`` package.
"""
from . import _lib, ops, data
from .graphs import graphed_loss, GraphedLoss
from .ops import (sphere_map, flow_regularization, entropic_ot, chamfer_nn, segmented_sort_raw, spherical_sliced_w1,
                  spherical_sliced_wp, euclid_sliced_w, exact_assignment, exact_emd2, exact_assignment_dense, exact_emd2_dense)

__all__ = ["_lib", "ops", "data", "sphere_map", "flow_regularization", "entropic_ot", "chamfer_nn", "segmented_sort_raw",
           "spherical_sliced_w1", "spherical_sliced_wp", "euclid_sliced_w", "exact_assignment", "exact_emd2", "exact_assignment_dense", "exact_emd2_dense", "build_library", "graphed_loss", "GraphedLoss"]


def build_library(force=False):
    from . import build as _build
    return _build.build(force=force)
