"""The C-ABI shared library loads on a CPU-only box and exports exactly what include/tnpy_cuda.h
declares (no compute calls here)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from tnpy_b200._cuda import build

    build.build()
    from tnpy_b200 import _cuda

    return _cuda.load()


def declared_functions():
    text = (ROOT / "include" / "tnpy_cuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tnpy_[a-z0-9_]+)\s*\(", text)))


def test_header_and_library_agree(lib):
    from tnpy_b200 import _cuda

    names = declared_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in tnpy_cuda.h but not exported"
        assert name in _cuda.SIGNATURES, f"{name} has no ctypes prototype"
    assert sorted(_cuda.SIGNATURES) == names


def test_host_only_entry_points(lib):
    assert lib.tnpy_version() >= 100
    assert lib.tnpy_launch_count() == 0
    assert lib.tnpy_heff_workspace_bytes(2048, 2048, 5, 5, 2) >= 2 * 5 * 2 * 2048 * 2048 * 8
    assert lib.tnpy_set_gemm_algo(7) < 0
    assert b"unknown algo" in lib.tnpy_last_error()


def test_solver_switches_and_workspace_queries_without_gpu(lib):
    # process-wide switch of the fused small-site / mid-size steps: returns the previous setting
    assert lib.tnpy_set_fused_steps(0) == 1
    assert lib.tnpy_set_fused_steps(1) == 0
    assert lib.tnpy_steps_trace(None) == 0
    # workspace queries are pure host arithmetic (they must not need a device)
    small = lib.tnpy_eig_workspace_bytes(60, 60, 5, 5, 2, 0)
    assert small >= 33 * 7200 * 8
    assert lib.tnpy_geig_chol_workspace_bytes(8192) >= 4 * 8192 * 8192 * 8
    assert lib.tnpy_heff_dense_workspace_bytes(64, 64, 25, 25, 2) >= 4 * 64 * 64 * 25 * 8
    assert lib.tnpy_geig_chol_workspace_bytes(0) == 0


def test_argument_validation_without_gpu(lib):
    rc = lib.tnpy_gemm_tn(None, 1, None, 1, None, 1, 1, 1, 1, 0, 0, None)
    assert rc == -1 and b"invalid argument" in lib.tnpy_last_error()
    rc = lib.tnpy_heff_apply(None, None, None, None, None, 0, 1, 1, 1, 2, 0, None, 0, None)
    assert rc == -1


def test_python_wrappers_refuse_cpu_tensors():
    import torch

    from tnpy_b200 import _cuda

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _cuda.gemm_tn(torch.zeros(2, 2, dtype=torch.float64), torch.zeros(2, 2, dtype=torch.float64))
