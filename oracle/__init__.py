"""CPU oracle for the sphere-homeomorphic Wasserstein loss path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this directory; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may.

Every function is a plain-torch restatement of the reference's own algorithm and cites the reference
``file:line`` it follows (paths relative to the reference tree root).  The reference is pure Python/torch, so
the restatement is torch on CPU (float32 for parity, float64 for diagnosis).

Pinning: the reference ships no golden vectors and no tests for this path.  The oracle is
pinned against (i) the one deterministic smoke input of ``Point_Cloud_Resistration/losses/Sinkhorn_fixed.py:97-110``
(8.621342658996582 for L1 / 8.161666870117188 for L2) and (ii) outputs of the *imported, unmodified* reference
modules run in the build container on seeded inputs, frozen under ``tests/golden/*.npz`` by
``tests/golden/make_golden.py``.  Third-party pieces the reference calls but does not vendor (POT ``ot.emd2``,
``pytorch3d.loss.chamfer_distance``; both un-pinned, neither installed) are restated from their published semantics
-- those two rows are "parity unpinned" and say so where they are defined.
"""
from .sphere import sphere_map, flow_regularization
from .costs import cost_matrix, COST_KINDS
from .sinkhorn import log_sinkhorn, entropic_w
from .sliced import (project_circle, emd1d_circle, sliced_wasserstein_sphere_p1, euclid_sliced_wasserstein,
                     binary_search_circle, sliced_wasserstein_sphere)
from .chamfer import chamfer_distance
from .emd import exact_emd2, exact_emd2_lp
from . import data

__all__ = [
    "sphere_map", "flow_regularization", "cost_matrix", "COST_KINDS", "log_sinkhorn", "entropic_w",
    "project_circle", "emd1d_circle", "sliced_wasserstein_sphere_p1", "euclid_sliced_wasserstein",
    "binary_search_circle", "sliced_wasserstein_sphere", "chamfer_distance", "exact_emd2", "exact_emd2_lp", "data",
]
