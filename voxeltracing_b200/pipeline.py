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
    passes: tuple = ("primary", "shadow")
    # attachments read back by the end-to-end path / gathered to rank 0
    outputs: tuple = (abi.ATT_INITIAL_T, abi.ATT_INITIAL_NORMAL, abi.ATT_INITIAL_BLOCK, abi.ATT_INITIAL_INVT,
                      abi.ATT_SHADOW, abi.ATT_SHADOW_TRANSVERSAL)


def orbit_camera(frame: int, aspect: float, n_frames: int = 64, radius: float = 120.0, height: float = 90.0,
                 centre=(192.0, 70.0, 192.0)) -> host_api.Camera:
    """Config-5 style camera path: orbit around the world centre, looking at it."""
    a = 2.0 * np.pi * (frame % n_frames) / n_frames
    pos = np.array([centre[0] + radius * np.cos(a), height, centre[2] + radius * np.sin(a)], dtype=np.float32)
    to = np.array(centre, dtype=np.float32) - pos
    yaw = float(np.degrees(np.arctan2(to[2], to[0])))
    pitch = float(np.degrees(np.arctan2(to[1], np.hypot(to[0], to[2]))))
    return host_api.camera(pos, yaw, pitch, aspect)


class FrameRenderer:
    """Issues the passes of one frame on a Context.  `tile` = (row0, rows) restricts every pass to a
    band of rows (screen-tile sharding); (0, 0) renders the whole frame."""

    def __init__(self, ctx: Context, cfg: FrameConfig):
        self.ctx = ctx
        self.cfg = cfg
        self.light = host_api.sun_direction(cfg.sun_tick)[2]

    def render(self, cam: host_api.Camera, frame: int = 0, tile=(0, 0), hook=None):
        """hook(name, 'begin'|'end') lets the bench bracket individual passes with CUDA events."""
        c, cfg = self.ctx, self.cfg
        for name in cfg.passes:
            if hook:
                hook(name, "begin")
            if name == "primary":
                c.initial_trace(cam, cfg.width, cfg.height, cfg.render_distance, tile=tile)
            elif name == "shadow":
                c.shadow_trace(cam, cfg.width, cfg.height, self.light, frame=frame, soft=cfg.soft_shadows,
                               max_iterations=cfg.shadow_iterations, tile=tile)
            else:
                raise ValueError(f"unknown pass {name!r}")
            if hook:
                hook(name, "end")

    def output_bytes(self) -> int:
        total = 0
        for att in self.cfg.outputs:
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
