"""Runs the C++ parity driver of the drop-in host classes (tests/cpp/test_drop_in.cpp) on the GPU box."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "cpp", "_build", "test_drop_in")
EXE_EIGEN = os.path.join(HERE, "cpp", "_build", "test_drop_in_eigen_flavour")


@pytest.mark.parametrize("exe", [EXE, EXE_EIGEN], ids=["shim", "eigen_flavour"])
def test_cpp_drop_in_classes(exe):
    """The drop-in headers over the C ABI, built twice: with the column-major shim, and with -DERL_GP_USE_EIGEN against a
    stand-in that has real Eigen's type structure (bool masks, alias templates, Ref<> class; tests/cpp/eigen_standin)."""
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(HERE, "cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(out.stdout)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-2000:]
    assert "ALL PASS" in out.stdout
    assert out.stdout.count("PASS ") >= 16
