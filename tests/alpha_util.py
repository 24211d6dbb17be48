"""Inputs of the alpha-tested traversal tests (VoxelTraversalDF_AlphaTest, off by default in the engine)."""
import numpy as np

import scene_util as su
from voxeltracing_b200 import abi, host_api

from pathlib import Path

W, H = 320, 180
GOLDEN = Path(__file__).resolve().parent / "golden" / "alpha_ref.npz"
GOLDEN_FOV = 70.0
# camera poses inside / next to tree crowns of plains(seed=0): (position, yaw, pitch)
POSES = [([200.5, 40.5, 8.5], 75.0, -12.0), ([192.0, 60.0, 192.0], 200.0, -25.0), ([120.0, 48.0, 260.0], 310.0, -8.0)]


def alpha_inputs(tex_size=64) -> su.SceneInputs:
    """Scene inputs whose oak-leaves albedo layer has a cut-out alpha channel with soft edges (values 0, 120, 230, 250, 255),
    so StopRay's `Alpha > 0.975` is decided differently per texel and per mip level."""
    inp = su.SceneInputs(tex_size)
    table = inp.table.reshape(6, 128)
    assert table[4][7] == 1, "oak_leaves (id 7) must be flagged Transparent"
    layer = int(table[0][7])
    tex = inp.textures[0].copy()
    yy, xx = np.mgrid[0:tex_size, 0:tex_size]
    cell = ((xx // 4) * 7 + (yy // 4) * 13) % 11
    alpha = np.choose(np.minimum(cell, 4), [0, 120, 230, 250, 255]).astype(np.uint8)
    alpha[cell >= 5] = 255
    alpha[(xx + yy) % 16 < 3] = 0
    tex[layer, :, :, 3] = alpha
    inp.textures = dict(inp.textures)
    inp.textures[0] = tex
    return inp


def primary_params(cam, alpha=True, fov=90.0, rd=350) -> abi.PrimaryParams:
    p = abi.PrimaryParams()
    su.fill(p.inv_view, cam.inv_view); su.fill(p.inv_projection, cam.inv_projection)
    p.width, p.height = W, H
    p.render_distance = rd
    p.alpha_test = int(alpha)
    p.fov = fov
    return p


def shadow_params(cam, alpha=True, fov=90.0, soft=False, frame=0) -> abi.ShadowParams:
    p = abi.ShadowParams()
    su.fill(p.inv_view, cam.inv_view); su.fill(p.inv_projection, cam.inv_projection)
    p.width, p.height = W, H
    su.fill(p.light_direction, host_api.sun_direction(50.0)[2])
    p.current_frame = frame
    p.soft_shadows = int(soft)
    p.alpha_test = int(alpha)
    p.max_iterations = 350
    p.fov = fov
    return p
