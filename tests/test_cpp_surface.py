"""The C++17 mirror of the reference interface (include/open3d_b200.hpp): compiled against the
in-tree library, argument validation on the host, and — on a GPU — ICP + Model::Integrate
checked against the oracle from C++ (tests/cpp/test_cpp_api.cpp)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
BIN = os.path.join(CPP, "test_cpp_api")


def _build():
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        if os.path.exists(BIN):
            return
        pytest.skip("nvcc not available and no prebuilt test binary")
    import oracle
    oracle.build()
    subprocess.run(["make", "-C", CPP], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def test_cpp_surface_host_mode():
    _build()
    out = subprocess.run([BIN, "host"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "cpp host-mode ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_surface_gpu_mode():
    _build()
    out = subprocess.run([BIN, "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "cpp gpu-mode ok" in out.stdout, out.stdout + out.stderr
