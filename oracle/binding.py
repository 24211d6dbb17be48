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
    L.vxo_traverse_batch.argtypes = [P(World), vp, vp, i32, i32, vp]
    L.vxo_traverse_batch.restype = None
    L.vxo_raycast_detect_batch.argtypes = [P(World), vp, vp, i32, vp]
    L.vxo_raycast_detect_batch.restype = None
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

    HIT_DTYPE = np.dtype([("t", "<f4"), ("normal", "<f4", 3), ("end", "<f4", 3), ("block", "<i4"), ("intersection", "<i4"),
                          ("min_idx", "<i4"), ("iterations", "<i4"), ("dda_steps", "<i4")])

    def traverse_batch(self, origins, directions, max_iter: int = 350) -> np.ndarray:
        """vxo_traverse over (n,3) float32 origins / directions; returns a structured array (HIT_DTYPE)."""
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 3)
        assert o.shape == d.shape and self.HIT_DTYPE.itemsize == C.sizeof(Hit)
        hits = np.zeros(len(o), dtype=self.HIT_DTYPE)
        lib().vxo_traverse_batch(C.byref(self.c), _p(o), _p(d), len(o), max_iter, _p(hits))
        return hits

    def raycast_detect(self, positions, directions) -> np.ndarray:
        """World::RaycastDetect over (n,3) rays; int32 (n,8): x, y, z, block, normal xyz, found."""
        o = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 3)
        out = np.zeros((len(o), 8), dtype=np.int32)
        lib().vxo_raycast_detect_batch(C.byref(self.c), _p(o), _p(d), len(o), _p(out))
        return out

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


# ---- scene (tables, texture arrays, sky map) + material / direct / GI / reflection passes ----------

class ReflectionInputs(C.Structure):
    _fields_ = [
        ("g_t_half", C.c_void_p), ("g_normal", C.c_void_p), ("gw", C.c_int32), ("gh", C.c_int32),
        ("gb_normal_h3", C.c_void_p), ("gb_pbr_u8x4", C.c_void_p), ("mw", C.c_int32), ("mh", C.c_int32),
        ("gi_sh_h4", C.c_void_p), ("gi_cocg_h2", C.c_void_p), ("gi_aosky_u8x2", C.c_void_p), ("iw", C.c_int32), ("ih", C.c_int32),
        ("shadow_u8", C.c_void_p), ("sw", C.c_int32), ("sh", C.c_int32),
    ]


def _scene_sigs(L):
    if getattr(L, "_scene_sigs_done", False):
        return
    vp, i32, P = C.c_void_p, C.c_int32, C.POINTER
    L.vxo_scene_create.restype = vp
    L.vxo_scene_create.argtypes = [P(World)]
    L.vxo_scene_destroy.argtypes = [vp]
    L.vxo_bind_alpha_scene.argtypes = [vp]
    L.vxo_scene_set_block_data.argtypes = [vp, vp]
    L.vxo_scene_set_blue_noise.argtypes = [vp, vp, i32]
    L.vxo_scene_set_texture_array.argtypes = [vp, i32, i32, i32, i32, vp]
    L.vxo_scene_set_skymap.argtypes = [vp, i32, vp]
    L.vxo_scene_texture_level.argtypes = [vp, i32, i32, vp, C.c_int64]
    L.vxo_scene_texture_level.restype = i32
    L.vxo_generate_gbuffer.argtypes = [vp, P(abi.GBufferParams), vp, vp, vp, i32, i32, vp, vp, vp, vp]
    L.vxo_shade_direct.argtypes = [P(abi.DirectParams), vp, i32, i32, vp, vp, vp, vp, i32, i32, vp, i32, i32, vp]
    L.vxo_diffuse_trace.argtypes = [vp, P(abi.GIParams), vp, vp, i32, i32, vp, vp, vp, vp, P(abi.TraceStats)]
    if hasattr(L, "vxo_reflection_trace"):
        L.vxo_reflection_trace.argtypes = [vp, P(abi.ReflectionParams), P(ReflectionInputs), vp, vp, vp, P(abi.TraceStats)]
    L._scene_sigs_done = True


def _stats(st):
    return {"rays": st.rays, "iterations": st.iterations, "dda_steps": st.dda_steps, "hits": st.hits}


class SvgfSet(C.Structure):
    _fields_ = [("sh", C.c_void_p), ("cocg", C.c_void_p), ("x", C.c_void_p), ("aosky", C.c_void_p)]


def svgf_set(d: dict) -> SvgfSet:
    """d: {"sh": f16 (h,w,4), "cocg": f16 (h,w,2), "x": f16 (h,w[,3]), "aosky": u8 (h,w,2)} (arrays must stay alive)."""
    return SvgfSet(d["sh"].ctypes.data, d["cocg"].ctypes.data, d["x"].ctypes.data, d["aosky"].ctypes.data)


def svgf_alloc(h: int, w: int, x_channels: int) -> dict:
    return {"sh": np.zeros((h, w, 4), np.float16), "cocg": np.zeros((h, w, 2), np.float16),
            "x": np.zeros((h, w, x_channels) if x_channels > 1 else (h, w), np.float16), "aosky": np.zeros((h, w, 2), np.uint8)}


def svgf_temporal(p: "abi.SvgfTemporalParams", cur: dict, hist: dict, g: dict, prev_g: dict, fn=None) -> dict:
    """TemporalFilter.glsl.  cur: raw GI set (x = luminance R16F), hist: previous temporal set (x = utility RGB16F),
    g / prev_g: {"t": f16, "normal": u8, "block": u8}.  fn: the entry point (oracle by default, oracle/_ref for pinning)."""
    out = svgf_alloc(p.height, p.width, 3)
    a, b, o = svgf_set(cur), svgf_set(hist), svgf_set(out)
    (fn or lib().vxo_svgf_temporal)(C.byref(p), C.byref(a), C.byref(b), _p(g["t"]), _p(g["normal"]), _p(g["block"]), _p(prev_g["t"]),
                                   _p(prev_g["normal"]), _p(prev_g["block"]), C.byref(o))
    return out


def svgf_prespatial(p: "abi.SvgfPreSpatialParams", raw: dict, g: dict, fn=None) -> dict:
    """Spatial3x3Initial.glsl.  raw: the GI trace set (x = utility R16F); g: {"t": f16, "normal": u8} of any size."""
    out = svgf_alloc(p.height, p.width, 1)
    a, o = svgf_set(raw), svgf_set(out)
    gh, gw = g["t"].shape
    (fn or lib().vxo_svgf_prespatial)(C.byref(p), C.byref(a), _p(g["t"]), _p(g["normal"]), gw, gh, C.byref(o))
    return out


def svgf_variance(p: "abi.SvgfVarianceParams", temporal: dict, g: dict, fn=None) -> dict:
    """VarianceEstimate.glsl.  temporal: this frame's temporal set (x = utility RGB16F).  Returns sh, cocg, x = variance R16F
    (aosky stays zero: VarianceFBO has no fourth attachment)."""
    out = svgf_alloc(p.height, p.width, 1)
    a, o = svgf_set(temporal), svgf_set(out)
    (fn or lib().vxo_svgf_variance)(C.byref(p), C.byref(a), _p(g["t"]), _p(g["normal"]), C.byref(o))
    return out


def svgf_spatial(p: "abi.SvgfSpatialParams", prev: dict, ao: np.ndarray, temporal_utility: np.ndarray, g: dict, fn=None) -> dict:
    """SpatialFilter.glsl, one a-trous iteration.  prev: sh / cocg / x = variance of the previous iteration (the variance
    pass's output first); ao: the RG8 image bound as u_AO; temporal_utility: RGB16F u_TemporalMoment."""
    out = svgf_alloc(p.height, p.width, 1)
    ao = np.ascontiguousarray(ao)
    a = SvgfSet(prev["sh"].ctypes.data, prev["cocg"].ctypes.data, prev["x"].ctypes.data, ao.ctypes.data)
    o = svgf_set(out)
    (fn or lib().vxo_svgf_spatial)(C.byref(p), C.byref(a), _p(temporal_utility), _p(g["t"]), _p(g["normal"]), C.byref(o))
    return out


def shadow_temporal(p: "abi.ShadowTemporalParams", raw: dict, hist: dict, g: dict, prev_t: np.ndarray, fn=None) -> dict:
    """ShadowTemporalFilter.glsl.  raw: {"shadow": u8 (sh, sw), "transversal": f16} of the shadow trace; hist: previous temporal
    set {"shadow": u8 (h, w), "frames": f16}; g: {"t": f16, "normal": u8}; prev_t: previous frame's hit distance."""
    out = {"shadow": np.zeros((p.height, p.width), np.uint8), "frames": np.zeros((p.height, p.width), np.float16)}
    sh, sw = raw["shadow"].shape
    gh, gw = g["t"].shape
    assert hist["shadow"].shape == (p.height, p.width) and prev_t.shape == (gh, gw)
    (fn or lib().vxo_shadow_temporal)(C.byref(p), _p(np.ascontiguousarray(raw["shadow"])), _p(np.ascontiguousarray(raw["transversal"])), sw, sh,
                                     _p(np.ascontiguousarray(hist["shadow"])), _p(np.ascontiguousarray(hist["frames"])), _p(np.ascontiguousarray(g["t"])),
                                     _p(np.ascontiguousarray(g["normal"])), _p(np.ascontiguousarray(prev_t)), gw, gh, _p(out["shadow"]), _p(out["frames"]))
    return out


def shadow_filter(p: "abi.ShadowFilterParams", temporal: dict, raw_transversal: np.ndarray, g: dict, fn=None) -> np.ndarray:
    """ShadowFilter.glsl.  temporal: this frame's temporal set; returns the filtered R8 image (p.height, p.width)."""
    out = np.zeros((p.height, p.width), np.uint8)
    ih, iw = temporal["shadow"].shape
    sh, sw = raw_transversal.shape
    gh, gw = g["t"].shape
    (fn or lib().vxo_shadow_filter)(C.byref(p), _p(np.ascontiguousarray(temporal["shadow"])), _p(np.ascontiguousarray(temporal["frames"])), iw, ih,
                                   _p(np.ascontiguousarray(raw_transversal)), sw, sh, _p(np.ascontiguousarray(g["t"])), _p(np.ascontiguousarray(g["normal"])),
                                   gw, gh, _p(out))
    return out


def specular_temporal(p: "abi.SpecularTemporalParams", cur: dict, prev_hitdist: np.ndarray, hist: dict, g: dict, prev_g: dict, pbr: np.ndarray,
                      fn=None) -> dict:
    """SpecularTemporalFilter.glsl.  cur: {"color": f16 (rh, rw, 4), "hitdist": f16, "mask": u8} of the reflection trace; prev_hitdist:
    the previous frame's hit distance (rh, rw); hist: previous temporal set {"color": f16 (h, w, 4), "frames": f16, "hitdist": f16};
    g / prev_g: {"t": f16, "normal": u8} of this / the previous frame; pbr: u8 (mh, mw, 4).  Returns this frame's temporal set."""
    out = {"color": np.zeros((p.height, p.width, 4), np.float16), "frames": np.zeros((p.height, p.width), np.float16),
           "hitdist": np.zeros((p.height, p.width), np.float16)}
    rh, rw = cur["hitdist"].shape
    gh, gw = g["t"].shape
    mh, mw = pbr.shape[:2]
    assert hist["color"].shape == (p.height, p.width, 4) and prev_hitdist.shape == (rh, rw) and prev_g["t"].shape == (gh, gw)
    c = np.ascontiguousarray
    f = fn
    if f is None:
        f = lib().vxo_specular_temporal
        f.restype = None
    f(C.byref(p), _p(c(cur["color"])), _p(c(cur["hitdist"])), _p(c(cur["mask"])), _p(c(prev_hitdist)), rw, rh, _p(c(hist["color"])), _p(c(hist["hitdist"])),
      _p(c(g["t"])), _p(c(g["normal"])), _p(c(prev_g["t"])), _p(c(prev_g["normal"])), gw, gh, _p(c(pbr)), mw, mh, _p(out["color"]), _p(out["frames"]),
      _p(out["hitdist"]))
    return out


def reflection_denoise(p: "abi.ReflectionDenoiseParams", in_color: np.ndarray, frames: np.ndarray, hitdist: np.ndarray, f: dict, fn=None,
                       time: float | None = None) -> np.ndarray:
    """ReflectionDenoiserNew.glsl, one direction.  in_color: f16 (ih, iw, 4); frames: f16 u_Frames; hitdist: f16 u_SpecularHitData; f: the frame
    {"g": {t, normal, block}, "gb_normal": f16 (mh, mw, 3), "pbr": u8 (mh, mw, 4)}.  Returns f16 (p.height, p.width, 4)."""
    out = np.zeros((p.height, p.width, 4), np.float16)
    ih, iw = in_color.shape[:2]
    th, tw = frames.shape
    hh, hw = hitdist.shape
    gh, gw = f["g"]["t"].shape
    mh, mw = f["pbr"].shape[:2]
    c = np.ascontiguousarray
    args = [C.byref(p), _p(c(in_color)), iw, ih, _p(c(frames)), _p(c(hitdist)), tw, th, hw, hh, _p(c(f["g"]["t"])), _p(c(f["g"]["normal"])),
            _p(c(f["g"]["block"])), gw, gh, _p(c(f["gb_normal"])), _p(c(f["pbr"])), mw, mh, _p(out)]
    if fn is None:
        g = lib().vxo_reflection_denoise
        g.restype = None
        g(*args)
    else:
        fn(*args, C.c_float(12.345 if time is None else time))   # the compiled shader also takes u_Time
    return out


class OracleScene:
    """The GL resources the material / GI / reflection shaders bind, on top of an OracleWorld."""

    def __init__(self, world: OracleWorld):
        self.world = world
        self.L = lib()
        _scene_sigs(self.L)
        self.h = self.L.vxo_scene_create(C.byref(world.c))
        self._keep = []

    def __del__(self):
        try:
            self.L.vxo_scene_destroy(self.h)
        except Exception:
            pass

    def set_block_data(self, table):
        t = np.ascontiguousarray(table, dtype=np.int32)
        self.L.vxo_scene_set_block_data(self.h, _p(t))

    # primary / shadow passes with this scene's block table and albedo array bound for the alpha-tested traversal
    # (params.alpha_test != 0; VoxelTraversalDF_AlphaTest)
    def initial_trace(self, params, want_stats: bool = False):
        self.L.vxo_bind_alpha_scene(self.h)
        try:
            return self.world.initial_trace(params, want_stats)
        finally:
            self.L.vxo_bind_alpha_scene(None)

    def shadow_trace(self, params, g_t, g_normal, blue_rgba, want_stats: bool = False):
        self.L.vxo_bind_alpha_scene(self.h)
        try:
            return self.world.shadow_trace(params, g_t, g_normal, blue_rgba, want_stats)
        finally:
            self.L.vxo_bind_alpha_scene(None)

    def set_blue_noise(self, data):
        d = np.ascontiguousarray(data, dtype=np.int32)
        self.L.vxo_scene_set_blue_noise(self.h, _p(d), d.size)

    def set_texture_array(self, kind, rgba):
        t = np.ascontiguousarray(rgba, dtype=np.uint8)
        self.L.vxo_scene_set_texture_array(self.h, kind, t.shape[0], t.shape[2], t.shape[1], _p(t))

    def set_skymap(self, faces):
        f = np.ascontiguousarray(faces, dtype=np.float32)
        self.L.vxo_scene_set_skymap(self.h, f.shape[1], _p(f))

    def set_lpv(self, level, block_type, avg512):
        """Light propagation volume + BlockAverageColorData for ApproximateGILPV (params.lpv_gi of the reflection pass)."""
        keep = (np.ascontiguousarray(level, np.uint8), np.ascontiguousarray(block_type, np.uint8), np.ascontiguousarray(avg512, np.float32))
        self._keep.append(keep)
        self.L.vxo_scene_set_lpv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.vxo_scene_set_lpv.restype = None
        self.L.vxo_scene_set_lpv(self.h, *[_p(a) for a in keep])

    def lpv_average_colors(self) -> np.ndarray:
        """BlockAverageColorData of PrecomputeAverageBlockColor.comp: (128, 4) float32"""
        out = np.zeros((128, 4), dtype=np.float32)
        self.L.vxo_scene_lpv_average_colors.argtypes = [C.c_void_p, C.c_void_p]
        self.L.vxo_scene_lpv_average_colors.restype = None
        self.L.vxo_scene_lpv_average_colors(self.h, _p(out))
        return out

    def texture_level(self, kind, level, layers, size):
        s = max(size >> level, 1)
        out = np.zeros((layers, s, s, 4), dtype=np.uint8)
        rc = self.L.vxo_scene_texture_level(self.h, kind, level, _p(out), out.nbytes)
        assert rc == 0, rc
        return out

    def generate_gbuffer(self, p: abi.GBufferParams, g_inv_t, g_normal, g_block):
        gh, gw = g_inv_t.shape
        w, h = p.width, p.height
        out = {"albedo": np.zeros((h, w, 3), np.float16), "normal": np.zeros((h, w, 3), np.float16),
               "pbr": np.zeros((h, w, 4), np.uint8), "texao": np.zeros((h, w), np.uint8)}
        self.L.vxo_generate_gbuffer(self.h, C.byref(p), _p(np.ascontiguousarray(g_inv_t, np.float32)), _p(np.ascontiguousarray(g_normal, np.uint8)),
                                    _p(np.ascontiguousarray(g_block, np.uint8)), gw, gh, _p(out["albedo"]), _p(out["normal"]), _p(out["pbr"]), _p(out["texao"]))
        return out

    def shade_direct(self, p: abi.DirectParams, g_inv_t, gb, shadow):
        gh, gw = g_inv_t.shape
        mh, mw = gb["texao"].shape
        sh, sw = shadow.shape
        out = np.zeros((p.height, p.width, 3), np.float16)
        self.L.vxo_shade_direct(C.byref(p), _p(np.ascontiguousarray(g_inv_t, np.float32)), gw, gh, _p(gb["albedo"]), _p(gb["normal"]), _p(gb["pbr"]),
                                _p(gb["texao"]), mw, mh, _p(np.ascontiguousarray(shadow, np.uint8)), sw, sh, _p(out))
        return out

    def diffuse_trace(self, p: abi.GIParams, g_t, g_normal):
        gh, gw = g_t.shape
        w, h = p.width, p.height
        out = {"sh": np.zeros((h, w, 4), np.float16), "cocg": np.zeros((h, w, 2), np.float16), "utility": np.zeros((h, w), np.float16),
               "aosky": np.zeros((h, w, 2), np.uint8)}
        st = abi.TraceStats()
        self.L.vxo_diffuse_trace(self.h, C.byref(p), _p(np.ascontiguousarray(g_t, np.float16)), _p(np.ascontiguousarray(g_normal, np.uint8)), gw, gh,
                                 _p(out["sh"]), _p(out["cocg"]), _p(out["utility"]), _p(out["aosky"]), C.byref(st))
        out["stats"] = _stats(st)
        return out

    def reflection_trace(self, p: abi.ReflectionParams, g_t, g_normal, gb, gi, shadow):
        gh, gw = g_t.shape
        mh, mw = gb["texao"].shape
        ih, iw = gi["utility"].shape
        sh, sw = shadow.shape
        keep = [np.ascontiguousarray(g_t, np.float16), np.ascontiguousarray(g_normal, np.uint8), np.ascontiguousarray(shadow, np.uint8)]
        ri = ReflectionInputs(keep[0].ctypes.data, keep[1].ctypes.data, gw, gh, gb["normal"].ctypes.data, gb["pbr"].ctypes.data, mw, mh,
                              gi["sh"].ctypes.data, gi["cocg"].ctypes.data, gi["aosky"].ctypes.data, iw, ih, keep[2].ctypes.data, sw, sh)
        w, h = p.width, p.height
        out = {"color": np.zeros((h, w, 4), np.float16), "hitdist": np.zeros((h, w), np.float16), "emissive": np.zeros((h, w), np.uint8)}
        st = abi.TraceStats()
        self.L.vxo_reflection_trace(self.h, C.byref(p), C.byref(ri), _p(out["color"]), _p(out["hitdist"]), _p(out["emissive"]), C.byref(st))
        out["stats"] = _stats(st)
        return out
