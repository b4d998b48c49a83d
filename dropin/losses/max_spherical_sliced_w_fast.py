"""``from losses.max_spherical_sliced_w_fast import ...`` drop-in (Point_Cloud_Resistration/losses/__init__.py:9-10): the batched
variant with per-pair frames; ``sliced_cost`` here is that module's batched function (max_spherical_sliced_w_fast.py:258-295)."""
from shwd_b200.losses.sliced import (binary_search_circle, emd1D_circle, max_spherical_wassersten_distance_fast,  # noqa: F401
                                     sliced_wasserstein_sphere_fast, transform_to_sphere_fast)
from shwd_b200.losses.sliced import sliced_cost_fast as sliced_cost  # noqa: F401
