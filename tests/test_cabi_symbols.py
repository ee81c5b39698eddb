"""The C-ABI library loads and exports every symbol include/open3d_b200.h declares.
No compute calls (runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "open3d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(o3db_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from open3d_b200 import _lib
    names = _declared_functions()
    assert len(names) >= 40
    handle = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing
    # and the python binding declares a signature for each of them
    assert sorted(_lib.EXPORTED_SYMBOLS) == names


def test_version_and_error_plumbing():
    from open3d_b200 import _lib
    assert _lib.lib.o3db_version() == 1
    assert _lib.lib.o3db_kernel_launch_count() >= 0
    # argument validation happens before any CUDA call and reports through o3db_last_error
    rc = _lib.lib.o3db_nns_create(None, 0, 0.1, None, ctypes.byref(ctypes.c_void_p()))
    assert rc == _lib.ERR_INVALID and "empty" in _lib.last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc)


def test_pose_to_transformation_host_math_matches_oracle():
    import numpy as np
    import oracle
    from open3d_b200 import _lib
    pose = np.array([0.03, -0.02, 0.01, 0.1, -0.2, 0.3])
    T = np.zeros((4, 4))
    _lib.lib.o3db_pose_to_transformation(_lib.dptr(pose), _lib.dptr(T))
    np.testing.assert_array_equal(T, oracle.pose_to_transformation(pose))


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "open3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "orc_" not in src, f
