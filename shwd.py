"""Import alias: ``import shwd`` loads the package directory
``sphere-homeomorphic-wasserstein-distance-for-point-cloud-registration_b200/`` (not a valid identifier) as the
module ``shwd_b200`` and re-exports it."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "sphere-homeomorphic-wasserstein-distance-for-point-cloud-registration_b200")
_NAME = "shwd_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_pkg = sys.modules[_NAME]

PACKAGE_DIR = _PKG_DIR
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})


def __getattr__(name):  # lazy sub-modules: shwd.losses, shwd.build, ...
    import importlib
    try:
        return importlib.import_module(_NAME + "." + name)
    except ModuleNotFoundError as e:
        raise AttributeError(name) from e
