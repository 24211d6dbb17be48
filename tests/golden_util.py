"""Shared helpers for the golden-vector tests (tests/golden/*.npz were produced by
tests/golden/make_golden.py from oracle/_ref, i.e. from the reference's own shaders)."""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import make_golden as mg  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def df_golden():
    z = np.load(GOLD / "df_ref.npz")
    hashes = dict(s.split(":") for s in z["hashes"])
    return z, hashes


def trace_golden():
    return np.load(GOLD / "trace_ref.npz")


def worlds():
    from voxeltracing_b200 import host_api

    w = {"plains0": host_api.gen_world("plains", 0), "rooms2": host_api.gen_world("rooms", 2)}
    e = w["plains0"].copy()
    host_api.random_edits(e, 1024, 1234)
    w["plains0_edited1024"] = e
    return w


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


BLUE = np.random.default_rng(11).integers(0, 256, (256, 256, 4), dtype=np.uint8)
