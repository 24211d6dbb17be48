"""The C-ABI libraries load and export every symbol their headers declare (no compute, CPU only)."""
import ctypes as C
from pathlib import Path

import pytest

from conftest import has_gpu
from voxeltracing_b200 import abi

ROOT = Path(__file__).resolve().parent.parent


def test_cuda_library_exports_every_declared_symbol():
    lib = abi.load_cuda()
    names = abi.declared_symbols(ROOT / "include" / "vxrt_cuda.h")
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/vxrt_cuda.h but not exported: {missing}"


def test_host_library_exports_every_declared_symbol():
    lib = abi.load_host()
    names = abi.declared_symbols(ROOT / "voxeltracing_b200" / "host" / "vxrt_host.h")
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing


def test_oracle_library_exports_every_declared_symbol():
    from oracle import binding as ob

    lib = ob.lib()
    names = abi.declared_symbols(ROOT / "oracle" / "vxrt_oracle.h")
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing


def test_product_library_does_not_link_the_oracle():
    import subprocess

    out = subprocess.run(["ldd", str(abi.CUDA_LIB_PATH)], capture_output=True, text=True).stdout
    assert "oracle" not in out and "vxrt_ref" not in out
    syms = subprocess.run(["nm", "-D", "--undefined-only", str(abi.CUDA_LIB_PATH)], capture_output=True, text=True).stdout
    assert "vxo_" not in syms


def test_argument_errors_are_status_codes_not_crashes():
    lib = abi.load_cuda()
    assert lib.vxrt_cuda_create(None, 0, None) < 0
    assert b"NULL" in lib.vxrt_cuda_last_error()
    h = C.c_void_p()
    bad = (C.c_int32 * 3)(30, 16, 16)  # nx not a multiple of 16
    assert lib.vxrt_cuda_create(C.byref(h), 0, bad) < 0 and not h.value
    assert lib.vxrt_cuda_upload_world(None, None) < 0
    assert lib.vxrt_cuda_generate_distance_field(None) < 0
    assert lib.vxrt_cuda_launch_count(None) == -1
    assert lib.vxrt_cuda_destroy(None) == 0


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_create_fails_loudly_without_a_gpu():
    lib = abi.load_cuda()
    h = C.c_void_p()
    rc = lib.vxrt_cuda_create(C.byref(h), 0, None)
    assert rc < 0 and not h.value
    assert len(lib.vxrt_cuda_last_error()) > 0
