import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _native_libs():
    """Build the product library, the host library and the oracle once per session (no-op when fresh)."""
    from voxeltracing_b200 import build

    build.build_cuda()
    build.build_host()
    build.build_oracle()
    try:
        build.build_ref()
    except Exception as e:  # the reference build is optional test infrastructure
        print("oracle/_ref build skipped:", e)


def has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def plains0():
    from voxeltracing_b200 import host_api

    return host_api.gen_world("plains", 0)


@pytest.fixture(scope="session")
def plains0_oracle(plains0):
    from oracle import binding as ob

    return ob.OracleWorld(plains0)


SMALL_DIMS = (32, 16, 48)  # nx, ny, nz


def random_small_world(seed: int, density: float, dims=SMALL_DIMS) -> np.ndarray:
    nx, ny, nz = dims
    rng = np.random.default_rng(seed)
    return ((rng.random((nz, ny, nx)) < density) * rng.integers(1, 100, (nz, ny, nx))).astype(np.uint8)
