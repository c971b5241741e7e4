"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/erl_gp_b200.h declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import shutil
import subprocess

import pytest

import erl_gaussian_process_b200 as gp
from erl_gaussian_process_b200 import _capi


def test_library_exports_every_declared_symbol():
    lib = gp.load()
    declared = _capi.declared_symbols()
    assert len(declared) >= 90
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in include/erl_gp_b200.h but not exported: {missing}"


def test_no_undeclared_exports():
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    exported = {s for s in exported if s.startswith("erl_gp_")}
    assert exported == set(_capi.declared_symbols())


def test_version_and_status_strings():
    lib = gp.load()
    assert lib.erl_gp_version() == 100
    for code in range(7):
        assert lib.erl_gp_status_string(code).decode() != "unknown status"
    assert "no CPU fallback" in lib.erl_gp_status_string(5).decode()


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert _capi.device_count() == 0
    with pytest.raises(gp.ErlGpError) as e:
        gp.Context(0)
    assert e.value.status == 5  # ERL_GP_STATUS_NO_DEVICE
    ctx = C.c_void_p()
    assert gp.load().erl_gp_context_create(0, C.byref(ctx)) == 5 and not ctx.value


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under erl_gaussian_process_b200/ may import, link or call it."""
    root = os.path.dirname(os.path.abspath(gp.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower(), f"{os.path.join(dirpath, f)} mentions the oracle"
    out = subprocess.run(["ldd", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_tc_sass_is_tcgen05():
    """The fused kernel's tensor work is tcgen05.mma with TMEM accumulators: its SASS carries UTCHMMA, LDTM and STTM."""
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    obj = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "csrc", "erl_gp_rowgp_tc_x3.o")
    if not os.path.exists(cuobjdump) or not os.path.exists(obj):
        pytest.skip("cuobjdump or the object file is not available on this box")
    sass = subprocess.run([cuobjdump, "-sass", obj], capture_output=True, text=True, timeout=300).stdout
    assert "RowGpTcKernel" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTCBAR"):
        assert mnemonic in sass, mnemonic


def test_tma_bulk_copies_in_sass():
    """The TMA engine moves L to HBM in both row-GP kernels (UBLKCP.G.S: shared -> global) and feeds the operands of the dense
    FP64 A B^T GEMM (UBLKCP.S.G: global -> shared, completing on an mbarrier: SYNCS)."""
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    build = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "csrc")
    objs = {"erl_gp_dense.o": ("GemmKernelDmmaTma", "UBLKCP.S.G", "UBLKPF", "SYNCS", "DMMA"), "erl_gp_rowgp64_x3.o": ("RowGp64Kernel", "UBLKCP.G.S", "DMMA")}
    if not os.path.exists(cuobjdump) or not all(os.path.exists(os.path.join(build, o)) for o in objs):
        pytest.skip("cuobjdump or the object files are not available on this box")
    for obj, needles in objs.items():
        sass = subprocess.run([cuobjdump, "-sass", os.path.join(build, obj)], capture_output=True, text=True, timeout=300).stdout
        for needle in needles:
            assert needle in sass, (obj, needle)
