"""``import losses`` drop-in: put this file's parent directory first on ``sys.path`` (``PYTHONPATH=<repo>/dropin``) and
the reference's train_W_COS.py / train_Pseudo_W_COS.py / Comparison_*/main_*.py pick up the B200 losses unchanged."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import shwd  # noqa: E402,F401  (registers the package as ``shwd_b200``)
from shwd_b200.losses import *  # noqa: E402,F401,F403
from shwd_b200.losses import __all__  # noqa: E402,F401
