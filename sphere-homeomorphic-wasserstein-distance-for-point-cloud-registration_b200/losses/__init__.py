"""Drop-in replacement for the reference's ``losses`` packages on the B200 path.

Exports the names of ``Point_Cloud_Resistration/losses/__init__.py:27-32`` and of
``Comparison_Wasserstein_with_Chamfer_distance/losses/__init__.py:1-2`` with the same call signatures, plus the
loss functions the reference imports from elsewhere on this path (``pytorch3d.loss.chamfer_distance``,
``sliced_wasserstein_sphere`` of ``max_spherical_sliced_w.py``, the notebooks' ``sliced_wasserstein_distance``).
Put the parent directory of this package first on ``sys.path`` and the reference's ``train_W_COS.py``,
``train_Pseudo_W_COS.py``, ``train_CD.py`` and ``Comparison_*/main_*.py`` import it unchanged.
"""
from .s2_wasserstein import (Cos_disimilarity_W, Geodesic_distance_W, Norm_Flow_structure, Norm_Flow_structure_optuna,
                             max_cos_disimilarity_wassersten_distance, pseudo_max_cos_disimilarity_wassersten_distance)
from .sinkhorn import (Sinkhorn_Distance_Loss, log_Sinkhorn_Distance_Loss, log_N_Sinkhorn_Distance_Loss,
                       log_Sinkhorn_Distance_Loss_fixed)
from .sliced import (sliced_wasserstein_sphere, sliced_cost, emd1D_circle, binary_search_circle,
                     sliced_wasserstein_distance, transform_to_sphere, max_spherical_wassersten_distance,
                     sliced_cost_fast, sliced_wasserstein_sphere_fast, transform_to_sphere_fast,
                     max_spherical_wassersten_distance_fast)
from .chamfer import chamfer_distance

__all__ = [
    "Cos_disimilarity_W", "Geodesic_distance_W", "Norm_Flow_structure", "Norm_Flow_structure_optuna",
    "max_cos_disimilarity_wassersten_distance", "pseudo_max_cos_disimilarity_wassersten_distance",
    "Sinkhorn_Distance_Loss", "log_Sinkhorn_Distance_Loss", "log_N_Sinkhorn_Distance_Loss",
    "log_Sinkhorn_Distance_Loss_fixed", "sliced_wasserstein_sphere", "sliced_cost", "emd1D_circle", "binary_search_circle",
    "sliced_wasserstein_distance", "chamfer_distance", "transform_to_sphere", "max_spherical_wassersten_distance",
    "sliced_cost_fast", "sliced_wasserstein_sphere_fast", "transform_to_sphere_fast", "max_spherical_wassersten_distance_fast",
]
