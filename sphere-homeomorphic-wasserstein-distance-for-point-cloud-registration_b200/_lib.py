"""ctypes binding of libshwd_b200.so (the C ABI declared in include/shwd.h).

There is no CPU fallback: if the library is missing or a call fails, this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SHWD_B200_LIB: diagnostics only (A/B timing of differently built libraries); the product always loads the in-tree build
LIB_PATH = os.environ.get("SHWD_B200_LIB") or os.path.join(_HERE, "libshwd_b200.so")

COST_GEODESIC, COST_SQEUCLID, COST_EUCLID, COST_ONE_MINUS_COS = 0, 1, 2, 3
COST_KINDS = {"geodesic": COST_GEODESIC, "sqeuclid": COST_SQEUCLID, "euclid": COST_EUCLID,
              "one_minus_cos": COST_ONE_MINUS_COS}
MAP_CENTER, MAP_NORMALIZE = 1, 2

_vp, _i, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/shwd.h declares (tests/test_abi.py checks this)
SIGNATURES = {
    "shwd_version": (_i, []),
    "shwd_error_string": (ctypes.c_char_p, [_i]),
    "shwd_device_sm_count": (_i, []),
    "shwd_sphere_map_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "shwd_sphere_map_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "shwd_sinkhorn_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "shwd_sinkhorn_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _f, _f, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "shwd_sinkhorn_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _f, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "shwd_sinkhorn_status_offset": (_i, []),
    "shwd_sinkhorn_lean_regime": (_i, [_i, _i, _i]),
    "shwd_sinkhorn_set_path": (_i, [_i]),
    "shwd_sinkhorn_plan_dense": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _f, _f, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "shwd_chamfer_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "shwd_chamfer_bwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "shwd_chamfer_reduce_workspace_bytes": (_sz, [_i]),
    "shwd_chamfer_reduce": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _f, _i, _vp, _vp, _vp]),
    "shwd_chamfer_bwd_uniform": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _f, _f, _vp, _vp, _vp]),
    "shwd_project_circle": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "shwd_project_circle_pp": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "shwd_project_circle_bwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "shwd_project_line": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "shwd_project_line_bwd": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "shwd_project_circle_bwd_scaled": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "shwd_project_circle_bwd_scaled_pp": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "shwd_project_line_bwd_scaled": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "shwd_segmented_sort_workspace_bytes": (_sz, [_i, _i]),
    "shwd_segmented_sort": (_i, [_vp, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "shwd_segmented_sort_i32": (_i, [_vp, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "shwd_sort_projected_max_points": (_i, []),
    "shwd_sort_set_method": (_i, [_i]),
    "shwd_sort_projected": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "shwd_sort_projected_pp": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "shwd_circular_w1_scatter": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "shwd_circular_wp_scatter": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "shwd_euclid_sw_scatter": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "shwd_circular_w1_workspace_bytes": (_sz, [_i, _i, _i]),
    "shwd_circular_w1": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "shwd_circular_wp_workspace_bytes": (_sz, [_i, _i, _i]),
    "shwd_circular_wp_set_dyadic": (_i, [_i]),
    "shwd_project_bwd_set_wide": (_i, [_i]),
    "shwd_circular_wp_weighted": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "shwd_circular_wp": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "shwd_euclid_sw": (_i, [_vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "shwd_unsort": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "shwd_resflow_params_per_layer": (_i, []),
    "shwd_resflow_workspace_bytes": (_sz, [_i, _i]),
    "shwd_resflow_uv_per_layer": (_i, []),
    "shwd_resflow_fwd": (_i, [_vp, _i, _vp, _vp, _i, _f, _vp, _vp]),
    "shwd_resflow_bwd": (_i, [_vp, _vp, _i, _vp, _vp, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    "shwd_planar_max_layers": (_i, []),
    "shwd_planar_params_per_layer": (_i, []),
    "shwd_planar_workspace_bytes": (_sz, [_i, _i, _i]),
    "shwd_planar_fwd": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "shwd_planar_bwd": (_i, [_vp, _vp, _vp, _i, _i, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "shwd_exact_assignment_max_points": (_i, []),
    "shwd_exact_assignment": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "shwd_exact_assignment_dense": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "shwd_rigid_transform": (_i, [_vp, _vp, _i, _i, _f, ctypes.c_ulonglong, _vp, _vp, _vp]),
    "shwd_peak_fp32": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_double), _vp]),
    "shwd_peak_mufu": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_double), _vp]),
}

_lib = None


def lib():
    """Load (once) and return the ctypes handle.  Raises if the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libshwd_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
                "`python <package>/build.py`; there is no CPU fallback." % LIB_PATH)
        h = ctypes.CDLL(LIB_PATH)
        missing = [name for name in SIGNATURES if not hasattr(h, name)]
        if missing:
            raise RuntimeError("libshwd_b200.so is stale: missing symbols %s -- rebuild it" % missing)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().shwd_error_string(code)
        raise RuntimeError("%s failed: %s (code %d)" % (what, msg.decode() if msg else "?", code))
