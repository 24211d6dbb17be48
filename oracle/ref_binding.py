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
HAVE = {"df": 7, "initial": 8, "shadow": 16, "gbuffer": 32, "diffuse": 64, "reflection": 128, "color": 256, "raycast": 512, "svgf_temporal": 1024, "svgf_variance": 2048, "svgf_spatial": 4096, "shadow_temporal": 8192, "shadow_filter": 16384, "specular_temporal": 32768, "reflection_denoise": 65536, "svgf_prespatial": 131072, "lpv_average": 262144}
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


def raycast_detect(blocks, positions, directions) -> np.ndarray:
    """World::RaycastDetect (the reference's own C++, lifted in place): int32 (n,4) = x, y, z, block; -2s when nothing is hit."""
    b = np.ascontiguousarray(blocks, np.uint8)
    assert b.shape == (384, 128, 384)
    o = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
    out = np.zeros((len(o), 4), np.int32)
    lib().vxref_raycast_detect(_p(b), _p(o), _p(d), len(o), _p(out))
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


# ---- scene resources + material / GI / reflection shaders -------------------------------------------
_keep = {}


def set_scene(blocks, df, table, blue_noise, textures, skymap):
    """textures: {kind: uint8[layers, size, size, 4]}"""
    L = lib()
    _keep["blocks"] = np.ascontiguousarray(blocks, np.uint8)
    _keep["df"] = np.ascontiguousarray(df, np.uint8)
    L.vxref_set_world(_p(_keep["blocks"]), _p(_keep["df"]))
    t = np.ascontiguousarray(table, np.int32)
    L.vxref_set_block_data(_p(t))
    b = np.ascontiguousarray(blue_noise, np.int32)
    L.vxref_set_blue_noise(_p(b), b.size)
    for kind, tex in textures.items():
        tex = np.ascontiguousarray(tex, np.uint8)
        L.vxref_set_texture_array(kind, tex.shape[0], tex.shape[2], tex.shape[1], _p(tex))
    f = np.ascontiguousarray(skymap, np.float32)
    L.vxref_set_skymap(f.shape[1], _p(f))


def set_lpv(level, block_type, avg512):
    """The propagation volume (two uint8 volumes) and BlockAverageColorData (128 x 4 float32) the reflection shader binds when
    params.lpv_gi is set."""
    _keep["lpv"] = (np.ascontiguousarray(level, np.uint8), np.ascontiguousarray(block_type, np.uint8), np.ascontiguousarray(avg512, np.float32))
    lib().vxref_set_lpv(*[_p(a) for a in _keep["lpv"]])


def lpv_average_colors() -> np.ndarray:
    """PrecomputeAverageBlockColor.comp on the scene of set_scene(): (128, 4) float32"""
    out = np.zeros((128, 4), dtype=np.float32)
    lib().vxref_lpv_average_colors(_p(out))
    return out


def generate_gbuffer(p: abi.GBufferParams, g_inv_t, g_normal, g_block):
    gh, gw = g_inv_t.shape
    w, h = p.width, p.height
    out = {"albedo": np.zeros((h, w, 3), np.float16), "normal": np.zeros((h, w, 3), np.float16),
           "pbr": np.zeros((h, w, 4), np.uint8), "texao": np.zeros((h, w), np.uint8)}
    lib().vxref_generate_gbuffer(C.byref(p), _p(np.ascontiguousarray(g_inv_t, np.float32)), _p(np.ascontiguousarray(g_normal, np.uint8)),
                                 _p(np.ascontiguousarray(g_block, np.uint8)), gw, gh, _p(out["albedo"]), _p(out["normal"]), _p(out["pbr"]), _p(out["texao"]))
    return out


def diffuse_trace(p: abi.GIParams, g_t, g_normal):
    gh, gw = g_t.shape
    w, h = p.width, p.height
    out = {"sh": np.zeros((h, w, 4), np.float16), "cocg": np.zeros((h, w, 2), np.float16), "utility": np.zeros((h, w), np.float16),
           "aosky": np.zeros((h, w, 2), np.uint8)}
    lib().vxref_diffuse_trace(C.byref(p), _p(np.ascontiguousarray(g_t, np.float16)), _p(np.ascontiguousarray(g_normal, np.uint8)), gw, gh,
                              _p(out["sh"]), _p(out["cocg"]), _p(out["utility"]), _p(out["aosky"]))
    return out


def shade_direct(p: abi.DirectParams, g_inv_t, gb, shadow):
    gh, gw = g_inv_t.shape
    mh, mw = gb["texao"].shape
    sh, sw = shadow.shape
    out = np.zeros((p.height, p.width, 3), np.float16)
    lib().vxref_shade_direct(C.byref(p), _p(np.ascontiguousarray(g_inv_t, np.float32)), gw, gh, _p(gb["albedo"]), _p(gb["normal"]), _p(gb["pbr"]),
                             _p(gb["texao"]), mw, mh, _p(np.ascontiguousarray(shadow, np.uint8)), sw, sh, _p(out))
    return out


def reflection_trace(p: abi.ReflectionParams, g_t, g_normal, gb, gi, shadow):
    from oracle.binding import ReflectionInputs

    gh, gw = g_t.shape
    mh, mw = gb["texao"].shape
    ih, iw = gi["utility"].shape
    sh, sw = shadow.shape
    keep = [np.ascontiguousarray(g_t, np.float16), np.ascontiguousarray(g_normal, np.uint8), np.ascontiguousarray(shadow, np.uint8)]
    ri = ReflectionInputs(keep[0].ctypes.data, keep[1].ctypes.data, gw, gh, gb["normal"].ctypes.data, gb["pbr"].ctypes.data, mw, mh,
                          gi["sh"].ctypes.data, gi["cocg"].ctypes.data, gi["aosky"].ctypes.data, iw, ih, keep[2].ctypes.data, sw, sh)
    w, h = p.width, p.height
    out = {"color": np.zeros((h, w, 4), np.float16), "hitdist": np.zeros((h, w), np.float16), "emissive": np.zeros((h, w), np.uint8)}
    lib().vxref_reflection_trace(C.byref(p), C.byref(ri), _p(out["color"]), _p(out["hitdist"]), _p(out["emissive"]))
    return out
