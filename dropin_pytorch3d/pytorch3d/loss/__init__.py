import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import shwd  # noqa: E402,F401
from shwd_b200.losses.chamfer import chamfer_distance  # noqa: E402,F401
