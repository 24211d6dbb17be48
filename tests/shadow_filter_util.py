"""Inputs for the sun-shadow denoiser tests: G-buffers and soft-shadow traces of a few nearby camera poses, produced by
the oracle (CPU) so the same arrays feed the oracle, oracle/_ref and the CUDA passes.  The raw trace runs at a lower
resolution than the temporal images, as in the engine (ShadowTraceResolution < ShadowSupersampleRes)."""
import numpy as np

import scene_util as su
from oracle import binding as ob
from voxeltracing_b200 import abi, host_api

W, H = 192, 108          # G-buffer and temporal / filtered images
SW, SH = 144, 81         # raw shadow trace (0.75 x)
POSES = [([192.0, 62.0, 192.0], 30.0, -15.0), ([192.3, 62.0, 191.8], 32.0, -14.5), ([192.5, 62.1, 191.6], 33.5, -14.0),
         ([192.6, 62.1, 191.5], 34.0, -14.0)]
BLUE = np.random.default_rng(11).integers(0, 256, (256, 256, 4), dtype=np.uint8)


def frames(world_blocks, light=None):
    """[{"cam", "g": {t, normal}, "raw": {shadow, transversal}} per frame]"""
    ow = ob.OracleWorld(world_blocks)
    light = host_api.sun_direction(50.0)[2] if light is None else light
    out = []
    for f, (pos, yaw, pitch) in enumerate(POSES):
        cam = host_api.camera(pos, yaw, pitch, W / H)
        p = abi.PrimaryParams()
        su.fill(p.inv_view, cam.inv_view); su.fill(p.inv_projection, cam.inv_projection)
        p.width, p.height, p.render_distance = W, H, 350
        g = ow.initial_trace(p)
        sp = abi.ShadowParams()
        su.fill(sp.inv_view, cam.inv_view); su.fill(sp.inv_projection, cam.inv_projection)
        sp.width, sp.height = SW, SH
        su.fill(sp.light_direction, light)
        sp.current_frame, sp.soft_shadows, sp.max_iterations = f, 1, 350
        s = ow.shadow_trace(sp, g["t"], g["normal"], BLUE)
        out.append({"cam": cam, "g": {"t": np.ascontiguousarray(g["t"]), "normal": np.ascontiguousarray(g["normal"])},
                    "raw": {"shadow": s["shadow"], "transversal": s["transversal"]}, "shadow_params": sp})
    return out


def temporal_params(cam, prev_cam, history_set, out_set, shadow_temporal=True) -> abi.ShadowTemporalParams:
    p = abi.ShadowTemporalParams()
    su.fill(p.inv_view, cam.inv_view); su.fill(p.inv_projection, cam.inv_projection)
    su.fill(p.prev_view, prev_cam.view); su.fill(p.prev_projection, prev_cam.projection)
    p.width, p.height, p.history_set, p.out_set, p.shadow_temporal = W, H, history_set, out_set, int(shadow_temporal)
    return p


def filter_params(cam, in_set, scale=1.0) -> abi.ShadowFilterParams:
    p = abi.ShadowFilterParams()
    su.fill(p.inv_view, cam.inv_view); su.fill(p.inv_projection, cam.inv_projection)
    p.width, p.height, p.in_set, p.filter_scale = W, H, in_set, scale
    return p


def sets_for(frame: int):
    """(history, out) temporal sets of a frame: ShadowTemporalFBO_1 / _2 by parity (Pipeline.cpp:1862-1863)"""
    return (abi.ATT_SHADOW_TEMPORAL_B, abi.ATT_SHADOW_TEMPORAL_A) if frame % 2 == 0 else (abi.ATT_SHADOW_TEMPORAL_A, abi.ATT_SHADOW_TEMPORAL_B)


def run_chain(seq, temporal_fn, filter_fn, shadow_temporal=True, scale=1.0):
    """Temporal pass over every frame (frame 0 against a zero history), spatial filter on each; [(temporal set, filtered)]."""
    hist = {"shadow": np.zeros((H, W), np.uint8), "frames": np.zeros((H, W), np.float16)}
    prev_t, prev_cam = np.zeros((H, W), np.float16), seq[0]["cam"]
    outs = []
    for k, f in enumerate(seq):
        hs, os_ = sets_for(k)
        t = temporal_fn(temporal_params(f["cam"], prev_cam, hs, os_, shadow_temporal), f["raw"], hist, f["g"], prev_t)
        flt = filter_fn(filter_params(f["cam"], os_, scale), t, f["raw"]["transversal"], f["g"])
        outs.append((t, flt))
        hist, prev_t, prev_cam = t, f["g"]["t"], f["cam"]
    return outs
