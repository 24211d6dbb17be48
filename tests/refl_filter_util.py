"""Inputs for the reflection temporal-filter tests: primary G-buffers of a few nearby camera poses from the oracle (so the
reprojection geometry is real) and seeded synthetic reflection-trace images (colour, hit distance incl. sky (< 0) and zero,
emissive mask, roughness / metalness planes that straddle every threshold of the shader) — the same arrays feed the oracle,
oracle/_ref and the CUDA pass.  The trace runs at a lower resolution than the temporal images, as in the engine."""
import numpy as np

import scene_util as su
from oracle import binding as ob
from voxeltracing_b200 import abi, host_api

W, H = 192, 108          # G-buffer, material G-buffer and temporal images
RW, RH = 96, 54          # reflection trace (0.5 x)
POSES = [([192.0, 62.0, 192.0], 30.0, -15.0), ([192.2, 62.0, 191.9], 30.3, -15.0), ([192.2, 62.0, 191.9], 30.3, -15.0),
         ([192.35, 62.05, 191.8], 30.5, -14.9), ([192.5, 62.1, 191.7], 30.8, -14.8)]   # frame 2 does not move (the clipping is skipped)


def _field(rng, h, w, cell, lo, hi):
    """low-frequency random field: coarse grid, bilinear-ish upsampling by repetition + a little per-pixel noise"""
    gh, gw = (h + cell - 1) // cell + 1, (w + cell - 1) // cell + 1
    coarse = rng.random((gh, gw))
    up = np.kron(coarse, np.ones((cell, cell)))[:h, :w]
    return lo + (hi - lo) * np.clip(up + 0.05 * rng.standard_normal((h, w)), 0.0, 1.0)


def frames(world_blocks):
    """[{"cam", "g": {t, normal}, "refl": {color, hitdist, mask}, "pbr"} per frame]"""
    ow = ob.OracleWorld(world_blocks)
    rng = np.random.default_rng(21)
    rough = _field(rng, H, W, 12, 0.0, 1.0)       # one material layout for the whole sequence
    metal = (_field(rng, H, W, 16, 0.0, 1.0) > 0.6) * 1.0
    out = []
    for f, (pos, yaw, pitch) in enumerate(POSES):
        cam = host_api.camera(pos, yaw, pitch, W / H)
        p = abi.PrimaryParams()
        su.fill(p.inv_view, cam.inv_view); su.fill(p.inv_projection, cam.inv_projection)
        p.width, p.height, p.render_distance = W, H, 350
        g = ow.initial_trace(p)
        color = np.stack([_field(rng, RH, RW, 6, 0.0, 1.5) for _ in range(3)] + [np.ones((RH, RW))], axis=2)
        mask = (_field(rng, RH, RW, 9, 0.0, 1.0) > 0.8).astype(np.uint8) * 255
        fire = (rng.random((RH, RW)) < 0.03) & (mask > 0)
        color[fire, :3] *= 12.0                    # fireflies where the trace hit an emitter
        hit = _field(rng, RH, RW, 8, 0.3, 30.0)
        sel = rng.random((RH, RW))
        hit[sel < 0.15] = -1.0                     # sky samples
        hit[(sel >= 0.15) & (sel < 0.18)] = 0.0
        pbr = np.zeros((H, W, 4), np.uint8)
        pbr[..., 0] = np.round(rough * 255); pbr[..., 1] = np.round(metal * 255); pbr[..., 2] = 128; pbr[..., 3] = 0
        # normal-mapped normals of the material G-buffer: the face normal, perturbed
        face = np.array([[0, 0, 1], [0, 0, -1], [0, 1, 0], [0, -1, 0], [-1, 0, 0], [1, 0, 0], [1, 1, 1]], np.float32)
        idx = np.minimum(np.round(g["normal"].astype(np.float32) / 255.0 * 10.0).astype(np.int32), 6)
        nm = face[idx] + 0.25 * np.stack([_field(rng, H, W, 3, -1.0, 1.0) for _ in range(3)], axis=2)
        nm /= np.linalg.norm(nm, axis=2, keepdims=True)
        out.append({"cam": cam, "pos": np.array(pos, np.float32),
                    "g": {"t": np.ascontiguousarray(g["t"]), "normal": np.ascontiguousarray(g["normal"]), "block": np.ascontiguousarray(g["block"])},
                    "refl": {"color": color.astype(np.float16), "hitdist": hit.astype(np.float16), "mask": mask}, "pbr": pbr,
                    "gb_normal": nm.astype(np.float16)})
    return out


def sets_for(frame: int):
    """(history, out) temporal sets of a frame: ReflectionTemporalFBO_1 / _2 by parity (Pipeline.cpp:1858-1859)"""
    return (abi.ATT_REFL_TEMPORAL_B, abi.ATT_REFL_TEMPORAL_A) if frame % 2 == 0 else (abi.ATT_REFL_TEMPORAL_A, abi.ATT_REFL_TEMPORAL_B)


DEFAULT_FLAGS = dict(temporal_spec=1, firefly_rejection=1, aggressive_firefly_rejection=1, smart_clip=1, roughness_weight=1, stabilize_hit_distance=1)


def params(f, prev, history_set, out_set, **flags) -> abi.SpecularTemporalParams:
    p = abi.SpecularTemporalParams()
    su.fill(p.inv_view, f["cam"].inv_view); su.fill(p.inv_projection, f["cam"].inv_projection)
    su.fill(p.prev_view, prev["cam"].view); su.fill(p.prev_projection, prev["cam"].projection)
    su.fill(p.current_camera_pos, f["pos"]); su.fill(p.prev_camera_pos, prev["pos"])
    p.width, p.height, p.history_set, p.out_set = W, H, history_set, out_set
    for k, v in {**DEFAULT_FLAGS, **flags}.items():
        setattr(p, k, int(v))
    return p


def zero_history():
    return {"color": np.zeros((H, W, 4), np.float16), "frames": np.zeros((H, W), np.float16), "hitdist": np.zeros((H, W), np.float16)}


def run_chain(seq, temporal_fn, **flags):
    """The temporal pass over every frame (frame 0 against zero-filled history images); [temporal set per frame]."""
    hist = zero_history()
    prev = seq[0]
    prev_g = {"t": np.zeros((H, W), np.float16), "normal": np.zeros((H, W), np.uint8)}
    prev_hit = np.zeros((RH, RW), np.float16)
    outs = []
    for k, f in enumerate(seq):
        hs, os_ = sets_for(k)
        t = temporal_fn(params(f, prev, hs, os_, **flags), f["refl"], prev_hit, hist, f["g"], prev_g, f["pbr"])
        outs.append(t)
        hist, prev, prev_g, prev_hit = t, f, f["g"], f["refl"]["hitdist"]
    return outs


DENOISE_DEFAULTS = dict(roughness_bias=1, normal_map_aware=1, handle_lobe_deviation=1, derive_from_diffuse_sh=0, amplify_transversal_weight=1,
                        temporal_weight=1, radius_bias=0, normal_map_weight_strength=0.75, denoiser_scale=1.0, resolution_scale=0.25,
                        roughness_normal_weight_bias_strength=1.075)


def denoise_params(f, direction: int, in_att: int, out_att: int, temporal_set: int, hit_att: int, **flags) -> abi.ReflectionDenoiseParams:
    p = abi.ReflectionDenoiseParams()
    su.fill(p.inv_view, f["cam"].inv_view); su.fill(p.inv_projection, f["cam"].inv_projection); su.fill(p.view, f["cam"].view)
    p.width, p.height, p.in_attachment, p.out_attachment, p.temporal_set, p.hit_distance_attachment, p.dir = W, H, in_att, out_att, temporal_set, hit_att, direction
    for k, v in {**DENOISE_DEFAULTS, **flags}.items():
        setattr(p, k, v)
    return p


def run_denoise(f, temporal: dict, temporal_set: int, denoise_fn, stabilized: bool = True, **flags):
    """x pass then y pass over one frame's temporal set (Pipeline.cpp:3404-3560); returns (x result, y result)."""
    hit = temporal["hitdist"] if stabilized else f["refl"]["hitdist"]
    hit_att = temporal_set + 2 if stabilized else abi.ATT_REFL_HITDIST
    px = denoise_params(f, 1, temporal_set, abi.ATT_REFL_DENOISED_A, temporal_set, hit_att, **flags)
    x = denoise_fn(px, temporal["color"], temporal["frames"], hit, f)
    py = denoise_params(f, 0, abi.ATT_REFL_DENOISED_A, abi.ATT_REFL_DENOISED_B, temporal_set, hit_att, **flags)
    y = denoise_fn(py, x, temporal["frames"], hit, f)
    return x, y
