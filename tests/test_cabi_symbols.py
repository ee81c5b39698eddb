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


def test_shipped_library_carries_the_blackwell_instructions_the_design_claims():
    """SASS of the in-tree libo3db200.so (sm_100a cubins): the TMA bulk copy of the ICP chunk ring (UBLKCP), the tiled
    TMA load of the integrate kernel (UTMALDG.2D), mbarrier transactions (SYNCS), programmatic dependent launch
    (ACQBULK / PREEXIT) and the cluster barrier of the level-resident odometry kernel (UCGABAR) — DESIGN.md §4,
    profiles/r02_sass_excerpts.md.  A build that silently lost one of them (a knob left on, an #if) fails here."""
    import shutil
    import subprocess
    from open3d_b200 import _lib
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    elfs = subprocess.run([tool, "-lelf", _lib.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
    assert "sm_100a" in elfs
    sass = subprocess.run([tool, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    current, per_kernel = None, {}
    for line in sass.splitlines():
        if "Function :" in line:
            current = line.split("Function :")[1].strip()
            per_kernel[current] = set()
        elif current is not None:
            for m in ("UBLKCP", "UTMALDG.2D", "SYNCS.ARRIVE.TRANS64", "ACQBULK", "PREEXIT", "UCGABAR_ARV", "UCGABAR_WAIT",
                      "LDGSTS"):
                if m in line:
                    per_kernel[current].add(m)

    def has(kernel_substring, *mnemonics):
        hits = [k for k, v in per_kernel.items() if kernel_substring in k and all(m in v for m in mnemonics)]
        assert hits, (kernel_substring, mnemonics)

    has("icp_iteration_kernel", "UBLKCP", "SYNCS.ARRIVE.TRANS64", "LDGSTS", "ACQBULK", "PREEXIT")
    has("integrate16_kernel", "UTMALDG.2D", "ACQBULK")
    has("touch_kernel", "ACQBULK")
    has("odometry_level_kernel", "UCGABAR_ARV", "UCGABAR_WAIT", "ACQBULK", "PREEXIT")
    has("pyramid_level_kernel", "ACQBULK", "PREEXIT")
