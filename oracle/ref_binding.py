"""ctypes binding of oracle/_ref/libvxrt_ref.so — the reference's own shaders compiled for the CPU by
oracle/build_ref.py.  TEST INFRASTRUCTURE ONLY (tests/, bench.py --impl reference, cpu_baseline)."""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from voxeltracing_b200 import abi  # noqa: E402  (struct layouts only)

LIB_PATH = ROOT / "oracle" / "_ref" / "libvxrt_ref.so"
HAVE = {"df": 7, "initial": 8, "shadow": 16, "gbuffer": 32, "diffuse": 64, "reflection": 128, "color": 256}
_lib = None


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def lib():
    """The library, or None if it has not been built (the reference tree is only mounted in the build
    container; the prebuilt .so travels with the repository snapshot)."""
    global _lib
    if _lib is None and LIB_PATH.exists():
        L = C.CDLL(str(LIB_PATH))
        L.vxref_available.restype = C.c_int32
        _lib = L
    return _lib


def available(what: str) -> bool:
    L = lib()
    return bool(L) and (L.vxref_available() & HAVE[what]) == HAVE[what]


def distance_field(blocks: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(blocks, dtype=np.uint8)
    assert b.shape == (384, 128, 384), "the reference shaders hard-code WORLD_SIZE 384x128x384"
    out = np.zeros_like(b)
    lib().vxref_distance_field(_p(b), _p(out))
    return out


def initial_trace(blocks, df, params: abi.PrimaryParams):
    w, h = params.width, params.height
    out = {"t": np.zeros((h, w), np.float16), "normal": np.zeros((h, w), np.uint8), "block": np.zeros((h, w), np.uint8),
           "inv_t": np.zeros((h, w), np.float32), "t32": np.zeros((h, w), np.float32)}
    lib().vxref_initial_trace(_p(blocks), _p(df), C.byref(params), _p(out["t"]), _p(out["normal"]), _p(out["block"]),
                              _p(out["inv_t"]), _p(out["t32"]))
    return out


def shadow_trace(blocks, df, params: abi.ShadowParams, g_t, g_normal, blue_rgba):
    w, h = params.width, params.height
    gh, gw = g_t.shape
    g_t = np.ascontiguousarray(g_t, np.float16)
    g_normal = np.ascontiguousarray(g_normal, np.uint8)
    if blue_rgba is None:
        blue_rgba = np.zeros((1, 1, 4), np.uint8)
    blue_rgba = np.ascontiguousarray(blue_rgba, np.uint8)
    out = {"shadow": np.zeros((h, w), np.uint8), "transversal": np.zeros((h, w), np.float16)}
    lib().vxref_shadow_trace(_p(blocks), _p(df), C.byref(params), _p(g_t), _p(g_normal), gw, gh, _p(blue_rgba),
                             blue_rgba.shape[1], blue_rgba.shape[0], _p(out["shadow"]), _p(out["transversal"]))
    return out
