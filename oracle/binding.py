"""ctypes binding of the CPU oracle (oracle/libvxrt_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package voxeltracing_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from voxeltracing_b200 import abi  # noqa: E402  (struct layouts of include/vxrt_cuda.h only)

LIB_PATH = ROOT / "oracle" / "libvxrt_oracle.so"


class World(C.Structure):
    _fields_ = [("blocks", C.c_void_p), ("df", C.c_void_p), ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32)]


class Hit(C.Structure):
    _fields_ = [
        ("t", C.c_float), ("normal", C.c_float * 3), ("end", C.c_float * 3), ("block", C.c_int32),
        ("intersection", C.c_int32), ("min_idx", C.c_int32), ("iterations", C.c_int32), ("dda_steps", C.c_int32),
    ]


_lib = None


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        from voxeltracing_b200 import build

        build.build_oracle()
    L = C.CDLL(str(LIB_PATH))
    vp, i32, P = C.c_void_p, C.c_int32, C.POINTER
    L.vxo_distance_field.argtypes = [vp, i32, i32, i32, vp]
    L.vxo_distance_field_literal.argtypes = [vp, i32, i32, i32, vp]
    L.vxo_distance_field_brute.argtypes = [vp, i32, i32, i32, vp]
    L.vxo_step_table.argtypes = [vp]
    L.vxo_traverse.argtypes = [P(World), vp, vp, i32, P(Hit)]
    L.vxo_traverse.restype = C.c_float
    L.vxo_plain_dda.argtypes = [P(World), vp, vp, i32, vp]
    L.vxo_plain_dda.restype = i32
    L.vxo_initial_trace.argtypes = [P(World), P(abi.PrimaryParams), vp, vp, vp, vp, vp, P(abi.TraceStats)]
    L.vxo_shadow_trace.argtypes = [P(World), P(abi.ShadowParams), vp, vp, i32, i32, vp, i32, i32, vp, vp, P(abi.TraceStats)]
    L.vxo_float_to_half.argtypes = [C.c_float]
    L.vxo_float_to_half.restype = C.c_uint16
    L.vxo_half_to_float.argtypes = [C.c_uint16]
    L.vxo_half_to_float.restype = C.c_float
    L.vxo_float_to_unorm8.argtypes = [C.c_float]
    L.vxo_float_to_unorm8.restype = C.c_uint8
    L.vxo_set_threads.argtypes = [i32]
    L.vxo_get_threads.restype = i32
    _lib = L
    return L


def set_threads(n: int):
    lib().vxo_set_threads(n)


def get_threads() -> int:
    return int(lib().vxo_get_threads())


def distance_field(blocks: np.ndarray, variant: str = "fast") -> np.ndarray:
    b = np.ascontiguousarray(blocks, dtype=np.uint8)
    nz, ny, nx = b.shape
    out = np.empty_like(b)
    fn = {"fast": lib().vxo_distance_field, "literal": lib().vxo_distance_field_literal,
          "brute": lib().vxo_distance_field_brute}[variant]
    fn(_p(b), nx, ny, nz, _p(out))
    return out


def step_table() -> np.ndarray:
    t = np.zeros(256, dtype=np.int32)
    lib().vxo_step_table(_p(t))
    return t


class OracleWorld:
    def __init__(self, blocks: np.ndarray, df: np.ndarray | None = None):
        self.blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
        self.df = np.ascontiguousarray(df, dtype=np.uint8) if df is not None else distance_field(self.blocks)
        nz, ny, nx = self.blocks.shape
        self.c = World(self.blocks.ctypes.data, self.df.ctypes.data, nx, ny, nz)

    def traverse(self, origin, direction, max_iter: int = 350) -> Hit:
        o = np.asarray(origin, dtype=np.float32)
        d = np.asarray(direction, dtype=np.float32)
        h = Hit()
        lib().vxo_traverse(C.byref(self.c), _p(o), _p(d), max_iter, C.byref(h))
        return h

    def plain_dda(self, origin, direction, max_steps: int = 2000):
        o = np.asarray(origin, dtype=np.float32)
        d = np.asarray(direction, dtype=np.float32)
        v = np.zeros(3, dtype=np.int32)
        hit = lib().vxo_plain_dda(C.byref(self.c), _p(o), _p(d), max_steps, _p(v))
        return bool(hit), v

    def initial_trace(self, params: abi.PrimaryParams, want_stats: bool = False):
        w, h = params.width, params.height
        out = {
            "t": np.zeros((h, w), dtype=np.float16),
            "normal": np.zeros((h, w), dtype=np.uint8),
            "block": np.zeros((h, w), dtype=np.uint8),
            "inv_t": np.zeros((h, w), dtype=np.float32),
            "t32": np.zeros((h, w), dtype=np.float32),
        }
        st = abi.TraceStats()
        lib().vxo_initial_trace(C.byref(self.c), C.byref(params), _p(out["t"]), _p(out["normal"]), _p(out["block"]),
                                _p(out["inv_t"]), _p(out["t32"]), C.byref(st))
        if want_stats:
            out["stats"] = {"rays": st.rays, "iterations": st.iterations, "dda_steps": st.dda_steps, "hits": st.hits}
        return out

    def shadow_trace(self, params: abi.ShadowParams, g_t: np.ndarray, g_normal: np.ndarray, blue_rgba: np.ndarray | None,
                     want_stats: bool = False):
        w, h = params.width, params.height
        gh, gw = g_t.shape
        g_t = np.ascontiguousarray(g_t, dtype=np.float16)
        g_normal = np.ascontiguousarray(g_normal, dtype=np.uint8)
        if blue_rgba is None:
            blue_rgba = np.zeros((1, 1, 4), dtype=np.uint8)
        blue_rgba = np.ascontiguousarray(blue_rgba, dtype=np.uint8)
        out = {"shadow": np.zeros((h, w), dtype=np.uint8), "transversal": np.zeros((h, w), dtype=np.float16)}
        st = abi.TraceStats()
        lib().vxo_shadow_trace(C.byref(self.c), C.byref(params), _p(g_t), _p(g_normal), gw, gh, _p(blue_rgba),
                               blue_rgba.shape[1], blue_rgba.shape[0], _p(out["shadow"]), _p(out["transversal"]), C.byref(st))
        if want_stats:
            out["stats"] = {"rays": st.rays, "iterations": st.iterations, "dda_steps": st.dda_steps, "hits": st.hits}
        return out
