"""``from losses.mini_batch_Residual_MSSW import ...`` drop-in (Point_Cloud_Resistration/losses/__init__.py:14-15)."""
from shwd_b200.losses.mini_batch_mssw import (Flow_structure, MLP_Architecture, max_spherical_wassersten_distance_Residual,  # noqa: F401
                                              transform_to_sphere)
from shwd_b200.losses.sliced import binary_search_circle, emd1D_circle, sliced_cost, sliced_wasserstein_sphere  # noqa: F401
