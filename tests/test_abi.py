"""CPU: the C-ABI shared library loads and exports every symbol include/shwd.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "shwd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(shwd_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path_entry_points():
    fns = header_functions()
    for must in ("shwd_sphere_map_fwd", "shwd_sphere_map_bwd", "shwd_sinkhorn_fwd", "shwd_sinkhorn_bwd", "shwd_chamfer_fwd",
                 "shwd_chamfer_bwd", "shwd_project_circle", "shwd_segmented_sort", "shwd_circular_w1", "shwd_euclid_sw"):
        assert must in fns


def test_binding_table_matches_header():
    import shwd
    assert sorted(shwd._lib.SIGNATURES) == header_functions()


def test_library_builds_and_exports_every_symbol():
    import __graft_entry__ as g
    path = g.build()
    h = ctypes.CDLL(path)
    for name in header_functions():
        assert hasattr(h, name), name
    h.shwd_version.restype = ctypes.c_int
    assert h.shwd_version() >= 100
    h.shwd_error_string.restype = ctypes.c_char_p
    assert h.shwd_error_string(-1) == b"invalid argument"
    # argument validation happens before any CUDA call, so it is safe without a GPU
    h.shwd_sinkhorn_workspace_bytes.restype = ctypes.c_size_t
    assert h.shwd_sinkhorn_workspace_bytes(32, 1024, 1024, 100) > 0
    assert h.shwd_sinkhorn_workspace_bytes(0, 1024, 1024, 100) == 0
    assert h.shwd_segmented_sort_workspace_bytes(8, 4096) == 0
    h.shwd_segmented_sort_workspace_bytes.restype = ctypes.c_size_t
    assert h.shwd_segmented_sort_workspace_bytes(2, 10000) == 0  # rows up to 16384 keys are sorted in shared memory
    assert h.shwd_segmented_sort_workspace_bytes(2, 20000) == 2 * 2 * 20000 * 8  # global scratch beyond


def test_no_product_code_imports_the_oracle():
    pkg = os.path.join(ROOT, "sphere-homeomorphic-wasserstein-distance-for-point-cloud-registration_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M):
                    offenders.append(os.path.join(base, f))
    assert not offenders, offenders
