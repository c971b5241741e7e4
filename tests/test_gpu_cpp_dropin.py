"""Runs the C++ parity driver of the drop-in host classes (tests/cpp/test_drop_in.cpp) on the GPU box."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "cpp", "_build", "test_drop_in")


def test_cpp_drop_in_classes():
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", os.path.join(HERE, "cpp")])
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    print(out.stdout)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-2000:]
    assert "ALL PASS" in out.stdout
    assert out.stdout.count("PASS ") >= 6
