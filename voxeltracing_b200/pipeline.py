"""Frame orchestration: the order in which Core/Pipeline.cpp issues the hot-path passes, as calls into
the C ABI.  A `Frame` names the workload of one bench step (BASELINE.json configs 3/4/5).

Pipeline.cpp order inside one frame: primary trace (:2038-2094) -> GenerateGBuffer
(:2147-2229) -> diffuse GI (:2267-2374) -> shadow trace (:2885-2945) -> reflection trace (:3095-3257)
-> colour pass direct term (:3702-3918).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import abi, host_api
from .engine import Context


@dataclass
class FrameConfig:
    width: int = 1920
    height: int = 1080
    render_distance: int = 350           # Pipeline.cpp:58,4931
    shadow_iterations: int = 350         # ShadowRayTraceFrag.glsl:233
    soft_shadows: bool = True
    sun_tick: float = 50.0               # Pipeline.cpp:67
    gi_spp: int = 1                      # BASELINE config 4 (engine default 3, Pipeline.cpp:76)
    gi_checkerboard: bool = False
    refl_spp: int = 1                    # BASELINE config 4 (engine default 2, Pipeline.cpp:108)
    refl_reproject: bool = False
    refl_lpv_gi: bool = False            # ApproximateGILPV (engine default on, Pipeline.cpp:116); needs Context.lpv_repropagate + lpv_average_colors
    refl_decoupled_gi: bool = False      # Pipeline.cpp:117
    passes: tuple = ("primary", "shadow")


PASS_OUTPUTS = {
    "primary": (abi.ATT_INITIAL_T, abi.ATT_INITIAL_NORMAL, abi.ATT_INITIAL_BLOCK, abi.ATT_INITIAL_INVT),
    "gbuffer": (abi.ATT_GBUF_ALBEDO, abi.ATT_GBUF_NORMAL, abi.ATT_GBUF_PBR, abi.ATT_GBUF_TEXAO),
    "gi": (abi.ATT_GI_SH, abi.ATT_GI_COCG, abi.ATT_GI_UTILITY, abi.ATT_GI_AOSKY),
    "shadow": (abi.ATT_SHADOW, abi.ATT_SHADOW_TRANSVERSAL),
    "reflection": (abi.ATT_REFL_COLOR, abi.ATT_REFL_HITDIST, abi.ATT_REFL_EMISSIVE),
    "direct": (abi.ATT_DIRECT,),
}
# output bytes per ray-tracing pixel: the W of the algorithmic-bytes formula
PASS_OUTPUT_BYTES = {"primary": 8, "shadow": 3, "gi": 16, "reflection": 11}
PASS_KERNEL = {"primary": "initial_trace_kernel", "shadow": "shadow_trace_kernel", "gi": "diffuse_trace_kernel",
               "reflection": "reflection_trace_kernel", "gbuffer": "generate_gbuffer_kernel", "direct": "shade_direct_kernel"}


def orbit_camera(frame: int, aspect: float, n_frames: int = 64, radius: float = 120.0, height: float = 90.0,
                 centre=(192.0, 70.0, 192.0)) -> host_api.Camera:
    """Config-5 style camera path: orbit around the world centre, looking at it."""
    a = 2.0 * np.pi * (frame % n_frames) / n_frames
    pos = np.array([centre[0] + radius * np.cos(a), height, centre[2] + radius * np.sin(a)], dtype=np.float32)
    to = np.array(centre, dtype=np.float32) - pos
    yaw = float(np.degrees(np.arctan2(to[2], to[0])))
    pitch = float(np.degrees(np.arctan2(to[1], np.hypot(to[0], to[2]))))
    return host_api.camera(pos, yaw, pitch, aspect)


def rooms_camera(frame: int, aspect: float, dims=(384, 128, 384)) -> host_api.Camera:
    """Config-4 camera path for the `rooms` stand-in: visits the 6x6 rooms of the building (vxh_gen_rooms:
    24-voxel cells around the grid centre, floor at ny/2-12), standing at head height in the cell centre and
    turning 37 degrees per frame."""
    nx, ny, nz = dims
    room, floor_y = 24, ny // 2 - 12
    cell = (frame * 7) % 36
    rx, rz = cell % 6, cell // 6
    pos = np.array([nx / 2 - 3 * room + rx * room + 12.5, floor_y + 6.5, nz / 2 - 3 * room + rz * room + 12.5], dtype=np.float32)
    return host_api.camera(pos, float((frame * 37) % 360), -10.0, aspect)


def _fill(dst, src):
    for i, v in enumerate(np.asarray(src).ravel()):
        dst[i] = v.item()


class FrameRenderer:
    """Issues the passes of one frame on a Context in the order of Core/Pipeline.cpp.  `tile` = (row0, rows)
    restricts every pass to a band of rows (screen-tile sharding); (0, 0) renders the whole frame.
    `scene` supplies the block-database-derived uniforms (grass / cactus face props)."""

    def __init__(self, ctx: Context, cfg: FrameConfig, grass_props=None, cactus_props=None):
        self.ctx = ctx
        self.cfg = cfg
        self.sun, self.moon, self.light = host_api.sun_direction(cfg.sun_tick)
        self.grass = np.zeros(10, np.int32) if grass_props is None else np.asarray(grass_props, np.int32)
        self.cactus = np.zeros(10, np.int32) if cactus_props is None else np.asarray(cactus_props, np.int32)
        # Pipeline.cpp:1906
        self.sun_visibility = float(np.clip(np.float32(self.sun[1]) + np.float32(0.05), 0.0, 0.1) * np.float32(12.0))

    @property
    def outputs(self):
        out = []
        for name in self.cfg.passes:
            out.extend(PASS_OUTPUTS[name])
        return tuple(out)

    def _gbuffer_params(self, cam, tile):
        p = abi.GBufferParams()
        _fill(p.inv_view, cam.inv_view); _fill(p.inv_projection, cam.inv_projection)
        p.width, p.height = self.cfg.width, self.cfg.height
        _fill(p.grass_props, self.grass); _fill(p.cactus_props, self.cactus)
        abi.set_tile(p.tile, tile)
        return p

    def _direct_params(self, cam, tile):
        p = abi.DirectParams()
        _fill(p.inv_view, cam.inv_view); _fill(p.inv_projection, cam.inv_projection)
        p.width, p.height = self.cfg.width, self.cfg.height
        _fill(p.viewer_position, cam.position); _fill(p.sun_direction, self.sun); _fill(p.moon_direction, self.moon)
        c = np.float32(np.pi) * np.float32(2.2) * np.float32(0.85)  # constant sun radiance
        _fill(p.sun_color, np.array([c, c, c], np.float32)); _fill(p.moon_color, np.array([0.12, 0.14, 0.25], np.float32))
        p.texture_desat_amount, p.amplify_normal_map = 0.1, 0
        abi.set_tile(p.tile, tile)
        return p

    def _gi_params(self, cam, frame, tile):
        cfg = self.cfg
        p = abi.GIParams()
        _fill(p.inv_view, cam.inv_view); _fill(p.inv_projection, cam.inv_projection)
        p.width, p.height = cfg.width, cfg.height
        p.spp, p.checker_spp, p.checkerboard = cfg.gi_spp, (cfg.gi_spp + cfg.gi_spp % 2) // 2, int(cfg.gi_checkerboard)
        p.trace_length, p.shadow_trace_length = 48, 128           # Pipeline.cpp:78, DiffuseRayTraceFrag.glsl:1306
        p.current_frame, p.current_frame_mod128 = frame, frame % 128
        p.use_blue_noise, p.supersample = 1, 0
        _fill(p.sun_direction, self.sun); _fill(p.moon_direction, self.moon)
        p.sun_visibility, p.gi_sun_strength, p.gi_sky_strength, p.diffuse_light_intensity = self.sun_visibility, 1.0, 1.125, 1.25
        _fill(p.viewer_position, cam.position)
        abi.set_tile(p.tile, tile)
        return p

    def _reflection_params(self, cam, frame, tile):
        cfg = self.cfg
        p = abi.ReflectionParams()
        _fill(p.inv_view, cam.inv_view); _fill(p.inv_projection, cam.inv_projection)
        _fill(p.view, cam.view); _fill(p.projection, cam.projection)
        p.width, p.height = cfg.width, cfg.height
        p.spp, p.checkerboard, p.trace_length, p.shadow_trace_length = cfg.refl_spp, 0, 64, 150   # Pipeline.cpp:105
        p.current_frame, p.current_frame_mod128 = frame, frame % 128
        p.use_blue_noise, p.rough_reflections, p.roughness_bias, p.temporal = 1, 1, 0, 0
        p.reproject_to_screen_space, p.derive_from_diffuse_sh = int(cfg.refl_reproject), 0
        p.lpv_gi, p.use_decoupled_gi, p.screen_space_skylighting_valid = int(cfg.refl_lpv_gi), int(cfg.refl_decoupled_gi), 0
        _fill(p.sun_direction, self.sun); _fill(p.moon_direction, self.moon); _fill(p.stronger_light_direction, self.light)
        _fill(p.viewer_position, cam.position)
        p.sun_strength_modifier, p.moon_strength_modifier = 0.85, 1.0
        _fill(p.grass_props, self.grass)
        abi.set_tile(p.tile, tile)
        return p

    def _primary_params(self, cam, tile):
        p = abi.PrimaryParams()
        _fill(p.inv_view, cam.inv_view); _fill(p.inv_projection, cam.inv_projection)
        p.width, p.height, p.render_distance = self.cfg.width, self.cfg.height, self.cfg.render_distance
        abi.set_tile(p.tile, tile)
        return p

    def _shadow_params(self, cam, frame, tile):
        p = abi.ShadowParams()
        _fill(p.inv_view, cam.inv_view); _fill(p.inv_projection, cam.inv_projection)
        p.width, p.height = self.cfg.width, self.cfg.height
        _fill(p.light_direction, self.light)
        p.current_frame, p.soft_shadows, p.max_iterations = frame, int(self.cfg.soft_shadows), self.cfg.shadow_iterations
        abi.set_tile(p.tile, tile)
        return p

    def params_for(self, name: str, cam, frame: int = 0, tile=(0, 0)):
        """The parameter block of pass `name` (the uniform set the reference uploads before the draw)."""
        if name == "primary":
            return self._primary_params(cam, tile)
        if name == "gbuffer":
            return self._gbuffer_params(cam, tile)
        if name == "gi":
            return self._gi_params(cam, frame, tile)
        if name == "shadow":
            return self._shadow_params(cam, frame, tile)
        if name == "reflection":
            return self._reflection_params(cam, frame, tile)
        if name == "direct":
            return self._direct_params(cam, tile)
        raise ValueError(f"unknown pass {name!r}")

    def prepare(self, cam: host_api.Camera, frame: int = 0, tile=(0, 0)):
        """Marshal the parameter blocks of a frame once; `submit` then only makes the C-ABI calls."""
        lib, h = self.ctx._lib, self.ctx._h
        fns = {"primary": lib.vxrt_cuda_initial_trace, "gbuffer": lib.vxrt_cuda_generate_gbuffer, "gi": lib.vxrt_cuda_diffuse_trace,
               "shadow": lib.vxrt_cuda_shadow_trace, "reflection": lib.vxrt_cuda_reflection_trace, "direct": lib.vxrt_cuda_shade_direct}
        return [(name, fns[name], self.params_for(name, cam, frame, tile)) for name in self.cfg.passes]

    def submit(self, prepared, hook=None):
        """hook(name, 'begin'|'end') lets the bench bracket individual passes with CUDA events."""
        import ctypes as C

        h = self.ctx._h
        for name, fn, params in prepared:
            if hook:
                hook(name, "begin")
            self.ctx._check(fn(h, C.byref(params)))
            if hook:
                hook(name, "end")
        self.ctx.join_passes()   # set_option("pass_overlap", 1): the caller's stream sees the whole frame from here on

    def render(self, cam: host_api.Camera, frame: int = 0, tile=(0, 0), hook=None):
        self.submit(self.prepare(cam, frame, tile), hook)

    def output_bytes(self) -> int:
        total = 0
        for att in self.outputs:
            _, w, h, bpp = self.ctx.attachment_info(att)
            total += w * h * bpp
        return total


def band_rows(height: int, rank: int, world: int, band: int = 8):
    """Row bands for screen-tile sharding: contiguous strips, multiples of the 8-row CTA tile."""
    bands = (height + band - 1) // band
    per = bands // world
    extra = bands % world
    b0 = rank * per + min(rank, extra)
    nb = per + (1 if rank < extra else 0)
    row0 = b0 * band
    rows = min(height, (b0 + nb) * band) - row0
    return row0, max(rows, 0)


class SvgfChain:
    """The SVGF denoiser chain of the diffuse GI in the order of Core/Pipeline.cpp:2428-2700: temporal accumulation
    (temporal sets ping-ponged by frame parity, :2420-2426; with pre_spatial, the engine's default PreTemporalSpatialPass, the 3 x 3
    pass of Spatial3x3Initial.glsl runs first, :2381-2424, and the temporal filter reads its output, :2488-2520), variance estimate, five a-trous iterations with steps
    16, 8, 4, 2, 1 ping-ponging the two denoise sets (:2592-2700), then the G-buffer hand-over to the next frame.
    Consumes the attachments of `primary` and `gi` of the same frame; the denoised result is `final_set`."""

    STEPS = (16, 8, 4, 2, 1)                    # Pipeline.cpp:2583-2589 (WiderSVGF off)
    # bytes every stage reads + writes per pixel when each image is touched once
    STAGE_BYTES = {"prespatial": 35, "temporal": 64, "variance": 35, "spatial": 41}
    final_set = abi.ATT_SVGF_DENOISE_A          # iteration 4 writes DiffuseDenoiseFBO

    def __init__(self, ctx: Context, width: int, height: int, color_phi_bias: float = 2.8, resolution_scale: float = 0.25,
                 large_kernel: bool = False, aggressive: bool = True, pre_spatial: bool = False):
        self.ctx, self.width, self.height = ctx, width, height
        self.pre_spatial = pre_spatial
        self.color_phi_bias, self.resolution_scale = color_phi_bias, resolution_scale     # Pipeline.cpp:86, 112
        self.large_kernel, self.aggressive = large_kernel, aggressive
        self.prev_cam = None

    def prepare(self, cam: host_api.Camera, frame: int, time: float | None = None, tile=(0, 0)):
        lib = self.ctx._lib
        prev = self.prev_cam or cam
        self.prev_cam = cam
        cur_t, hist_t = (abi.ATT_SVGF_TEMPORAL_A, abi.ATT_SVGF_TEMPORAL_B) if frame % 2 == 0 else (abi.ATT_SVGF_TEMPORAL_B, abi.ATT_SVGF_TEMPORAL_A)
        out = []
        if self.pre_spatial:
            pp = abi.SvgfPreSpatialParams()
            _fill(pp.inv_view, cam.inv_view); _fill(pp.inv_projection, cam.inv_projection)
            pp.width, pp.height, pp.in_set = self.width, self.height, abi.ATT_GI_SH
            pp.time = frame / 60.0 if time is None else time
            abi.set_tile(pp.tile, tile)
            out.append(("prespatial", lib.vxrt_cuda_svgf_prespatial, pp))
        tp = abi.SvgfTemporalParams()
        _fill(tp.inv_view, cam.inv_view); _fill(tp.inv_projection, cam.inv_projection)
        _fill(tp.prev_view, prev.view); _fill(tp.prev_projection, prev.projection)
        tp.width, tp.height, tp.history_set, tp.out_set, tp.be_useful = self.width, self.height, hist_t, cur_t, 1
        tp.in_set = abi.ATT_SVGF_PRESPATIAL if self.pre_spatial else abi.ATT_GI_SH
        abi.set_tile(tp.tile, tile)
        out.append(("temporal", lib.vxrt_cuda_svgf_temporal, tp))
        vp = abi.SvgfVarianceParams()
        _fill(vp.inv_view, cam.inv_view); _fill(vp.inv_projection, cam.inv_projection)
        vp.width, vp.height, vp.in_set, vp.do_spatial, vp.aggressive_disocclusion = self.width, self.height, cur_t, 1, int(self.aggressive)
        abi.set_tile(vp.tile, tile)
        out.append(("variance", lib.vxrt_cuda_svgf_variance, vp))
        for i, step in enumerate(self.STEPS):
            cur = abi.ATT_SVGF_DENOISE_A if i % 2 == 0 else abi.ATT_SVGF_DENOISE_B
            prv = abi.ATT_SVGF_VARIANCE if i == 0 else (abi.ATT_SVGF_DENOISE_B if i % 2 == 0 else abi.ATT_SVGF_DENOISE_A)
            sp = abi.SvgfSpatialParams()
            _fill(sp.inv_view, cam.inv_view); _fill(sp.inv_projection, cam.inv_projection)
            sp.width, sp.height = self.width, self.height
            sp.in_set, sp.ao_set, sp.temporal_set, sp.out_set, sp.step = prv, (cur_t if i == 0 else prv), cur_t, cur, step
            sp.large_kernel, sp.do_spatial, sp.aggressive_disocclusion = int(self.large_kernel), 1, int(self.aggressive)
            sp.color_phi_bias, sp.resolution_scale = self.color_phi_bias, self.resolution_scale
            sp.time = frame / 60.0 if time is None else time
            abi.set_tile(sp.tile, tile)
            out.append((f"spatial{i}", lib.vxrt_cuda_svgf_spatial, sp))
        return out

    def submit(self, prepared, hook=None, end_frame: bool = True):
        import ctypes as C

        h = self.ctx._h
        for name, fn, params in prepared:
            if hook:
                hook(name, "begin")
            self.ctx._check(fn(h, C.byref(params)))
            if hook:
                hook(name, "end")
        if end_frame:
            self.ctx.svgf_end_frame()

    def run(self, cam: host_api.Camera, frame: int, time: float | None = None, tile=(0, 0), hook=None):
        self.submit(self.prepare(cam, frame, time, tile), hook)


class ShadowDenoiser:
    """The sun-shadow denoiser in the order of Core/Pipeline.cpp:2947-3044: temporal filter (temporal sets ping-ponged by
    frame parity, :1862-1863), then the spatial filter; `select=True` makes the result the shadow texture of the reflection
    and colour passes, as the engine binds it (:3231, :3838).  Consumes `primary` and `shadow` of the same frame."""

    STAGE_BYTES = {"temporal": 16, "filter": 9}   # bytes read + written per pixel when each image is touched once

    def __init__(self, ctx: Context, width: int, height: int, filter_scale: float = 1.0, select: bool = True):
        self.ctx, self.width, self.height, self.filter_scale, self.select = ctx, width, height, filter_scale, select
        self.prev_cam = None

    def prepare(self, cam: host_api.Camera, frame: int, tile=(0, 0)):
        lib = self.ctx._lib
        prev = self.prev_cam or cam
        self.prev_cam = cam
        hist, out = (abi.ATT_SHADOW_TEMPORAL_B, abi.ATT_SHADOW_TEMPORAL_A) if frame % 2 == 0 else (abi.ATT_SHADOW_TEMPORAL_A, abi.ATT_SHADOW_TEMPORAL_B)
        tp = abi.ShadowTemporalParams()
        _fill(tp.inv_view, cam.inv_view); _fill(tp.inv_projection, cam.inv_projection)
        _fill(tp.prev_view, prev.view); _fill(tp.prev_projection, prev.projection)
        tp.width, tp.height, tp.history_set, tp.out_set, tp.shadow_temporal = self.width, self.height, hist, out, 1
        abi.set_tile(tp.tile, tile)
        fp = abi.ShadowFilterParams()
        _fill(fp.inv_view, cam.inv_view); _fill(fp.inv_projection, cam.inv_projection)
        fp.width, fp.height, fp.in_set, fp.filter_scale = self.width, self.height, out, self.filter_scale
        abi.set_tile(fp.tile, tile)
        return [("temporal", lib.vxrt_cuda_shadow_temporal, tp), ("filter", lib.vxrt_cuda_shadow_filter, fp)]

    def submit(self, prepared, hook=None):
        import ctypes as C

        for name, fn, params in prepared:
            if hook:
                hook(name, "begin")
            self.ctx._check(fn(self.ctx._h, C.byref(params)))
            if hook:
                hook(name, "end")
        if self.select:
            self.ctx.select_shadow(abi.ATT_SHADOW_FILTERED)

    def run(self, cam: host_api.Camera, frame: int, tile=(0, 0), hook=None):
        self.submit(self.prepare(cam, frame, tile), hook)


class ReflectionTemporal:
    """The reflection temporal filter in the order of Core/Pipeline.cpp:3316-3400: temporal sets ping-ponged by frame parity
    (:1858-1859).  Consumes `primary`, `gbuffer` and `reflection` of the same frame; `ctx.end_frame()` hands this frame's
    G-buffer and reflection hit distance to the next one."""

    # bytes read + written per pixel when each image is touched once.  temporal: trace (8+2+1), history colour + hit distance (8+2),
    # previous hit distance 2, both G-buffers (2+1+2+1), PBR 4, outputs (8+2+2); denoise: colour in / out (8+8), frames 2, hit distance 2,
    # G-buffer (2+1), material normals 6, PBR 4
    STAGE_BYTES = {"temporal": 45, "denoise_x": 33, "denoise_y": 33}

    def __init__(self, ctx: Context, width: int, height: int, denoise: bool = True, resolution_scale: float = 0.25):
        self.ctx, self.width, self.height, self.denoise, self.resolution_scale = ctx, width, height, denoise, resolution_scale
        self.prev_cam = None

    def prepare(self, cam: host_api.Camera, frame: int, tile=(0, 0), **flags):
        lib = self.ctx._lib
        prev = self.prev_cam or cam
        self.prev_cam = cam
        hist, out = (abi.ATT_REFL_TEMPORAL_B, abi.ATT_REFL_TEMPORAL_A) if frame % 2 == 0 else (abi.ATT_REFL_TEMPORAL_A, abi.ATT_REFL_TEMPORAL_B)
        p = abi.SpecularTemporalParams()
        _fill(p.inv_view, cam.inv_view); _fill(p.inv_projection, cam.inv_projection)
        _fill(p.prev_view, prev.view); _fill(p.prev_projection, prev.projection)
        _fill(p.current_camera_pos, np.asarray(cam.inv_view, dtype=np.float32).reshape(-1)[12:15])
        _fill(p.prev_camera_pos, np.asarray(prev.inv_view, dtype=np.float32).reshape(-1)[12:15])
        p.width, p.height, p.history_set, p.out_set = self.width, self.height, hist, out
        for k, v in {"temporal_spec": 1, "firefly_rejection": 1, "aggressive_firefly_rejection": 1, "smart_clip": 1, "roughness_weight": 1,
                     "stabilize_hit_distance": 1, **flags}.items():
            setattr(p, k, int(v))
        abi.set_tile(p.tile, tile)
        self.out_set = out
        passes = [("temporal", lib.vxrt_cuda_specular_temporal, p)]
        if self.denoise:   # x then y pass of ReflectionDenoiserNew.glsl (Pipeline.cpp:3404-3560)
            stabilized = bool(p.temporal_spec and p.stabilize_hit_distance)
            for name, direction, src, dst in (("denoise_x", 1, out, abi.ATT_REFL_DENOISED_A), ("denoise_y", 0, abi.ATT_REFL_DENOISED_A, abi.ATT_REFL_DENOISED_B)):
                d = abi.ReflectionDenoiseParams()
                _fill(d.inv_view, cam.inv_view); _fill(d.inv_projection, cam.inv_projection); _fill(d.view, cam.view)
                d.width, d.height, d.in_attachment, d.out_attachment, d.temporal_set, d.dir = self.width, self.height, src, dst, out, direction
                d.hit_distance_attachment = out + 2 if stabilized else abi.ATT_REFL_HITDIST
                d.roughness_bias = d.normal_map_aware = d.handle_lobe_deviation = d.amplify_transversal_weight = 1
                d.temporal_weight, d.derive_from_diffuse_sh, d.radius_bias = int(bool(p.temporal_spec)), 0, 0
                d.normal_map_weight_strength, d.denoiser_scale, d.resolution_scale, d.roughness_normal_weight_bias_strength = 0.75, 1.0, self.resolution_scale, 1.075
                abi.set_tile(d.tile, tile)
                passes.append((name, lib.vxrt_cuda_reflection_denoise, d))
        return passes

    def submit(self, prepared, hook=None):
        import ctypes as C

        for name, fn, params in prepared:
            if hook:
                hook(name, "begin")
            self.ctx._check(fn(self.ctx._h, C.byref(params)))
            if hook:
                hook(name, "end")

    def run(self, cam: host_api.Camera, frame: int, tile=(0, 0), hook=None):
        self.submit(self.prepare(cam, frame, tile), hook)
