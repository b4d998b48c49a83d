"""``from losses.max_spherical_sliced_w import ...`` drop-in (the import lines the reference keeps, commented out, in
Point_Cloud_Resistration/losses/__init__.py:4-5): the sliced functions, the small sphere map and the max-SSW wrapper."""
from shwd_b200.losses.sliced import (binary_search_circle, emd1D_circle, max_spherical_wassersten_distance,  # noqa: F401
                                     sliced_cost, sliced_wasserstein_sphere, transform_to_sphere)
