"""Builds the inputs of the material / GI / reflection passes for tests and bench: block-data table from
the synthetic block database, synthetic texture arrays, the blue-noise tables fixture, a sky map."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from voxeltracing_b200 import abi, host_api  # noqa: E402

BLOCKDB = ROOT / "tests" / "data" / "blockdb_synth.txt"


def blue_noise_tables() -> np.ndarray:
    """sobol ++ scramble ++ ranking as int32[327680] (Core/BlueNoiseDataSSBO.cpp:19-25)."""
    fx = ROOT / "tests" / "golden" / "blue_noise_tables.npz"
    if fx.exists():
        z = np.load(fx)
        return np.concatenate([z["sobol_256spp_256d"], z["scramblingTile"], z["rankingTile"]]).astype(np.int32)
    rng = np.random.default_rng(5)  # "own tables": any bytes exercise the sampler identically
    return rng.integers(0, 256, 327680, dtype=np.int32)


class SceneInputs:
    def __init__(self, tex_size: int = 128, sky: str = "gradient"):
        self.db = host_api.BlockDatabase(BLOCKDB)
        self.table = self.db.table()
        self.grass = self.db.face_props("Grass")
        self.cactus = self.db.face_props("Cactus")
        self.blue = blue_noise_tables()
        self.textures = {k: host_api.gen_texture_array(k, len(self.db.layer_paths(k)), tex_size) for k in range(4)}
        self.sky = host_api.gradient_skymap(16) if sky == "gradient" else host_api.constant_skymap(16)

    def apply_to_context(self, ctx):
        ctx.set_block_data(self.table)
        ctx.set_blue_noise(self.blue)
        for k, t in self.textures.items():
            ctx.set_texture_array(k, t)
        ctx.set_skymap(self.sky)

    def apply_to_oracle(self, scene):
        scene.set_block_data(self.table)
        scene.set_blue_noise(self.blue)
        for k, t in self.textures.items():
            scene.set_texture_array(k, t)
        scene.set_skymap(self.sky)


def fill(dst, src):
    for i, v in enumerate(np.asarray(src).ravel()):
        dst[i] = v.item() if hasattr(v, "item") else v


def gbuffer_params(cam, w, h, inputs: SceneInputs, tile=(0, 0)) -> abi.GBufferParams:
    p = abi.GBufferParams()
    fill(p.inv_view, cam.inv_view); fill(p.inv_projection, cam.inv_projection)
    p.width, p.height = w, h
    fill(p.grass_props, inputs.grass); fill(p.cactus_props, inputs.cactus)
    abi.set_tile(p.tile, tile)
    return p


def direct_params(cam, w, h, sun_tick=50.0, tile=(0, 0)) -> abi.DirectParams:
    p = abi.DirectParams()
    fill(p.inv_view, cam.inv_view); fill(p.inv_projection, cam.inv_projection)
    p.width, p.height = w, h
    sun, moon, _ = host_api.sun_direction(sun_tick)
    fill(p.viewer_position, cam.position); fill(p.sun_direction, sun); fill(p.moon_direction, moon)
    c = np.float32(np.pi) * np.float32(2.2) * np.float32(0.85)   # (1,1,1)*pi*2.2*0.85
    fill(p.sun_color, [c, c, c]); fill(p.moon_color, [0.12, 0.14, 0.25])
    p.texture_desat_amount = 0.1
    p.amplify_normal_map = 0
    abi.set_tile(p.tile, tile)
    return p


def gi_params(cam, w, h, frame=0, spp=1, checkerboard=False, sun_tick=50.0, tile=(0, 0)) -> abi.GIParams:
    p = abi.GIParams()
    fill(p.inv_view, cam.inv_view); fill(p.inv_projection, cam.inv_projection)
    p.width, p.height = w, h
    p.spp, p.checker_spp, p.checkerboard = spp, (spp + spp % 2) // 2, int(checkerboard)
    p.trace_length, p.shadow_trace_length = 48, 128
    p.current_frame, p.current_frame_mod128 = frame, frame % 128
    p.use_blue_noise, p.supersample = 1, 0
    sun, moon, _ = host_api.sun_direction(sun_tick)
    fill(p.sun_direction, sun); fill(p.moon_direction, moon)
    sv = float(np.clip(np.float32(sun[1]) + np.float32(0.05), 0.0, 0.1) * np.float32(12.0))  # Pipeline.cpp:1906
    p.sun_visibility, p.gi_sun_strength, p.gi_sky_strength, p.diffuse_light_intensity = sv, 1.0, 1.125, 1.25
    fill(p.viewer_position, cam.position)
    p.apply_player_shadow = 0
    abi.set_tile(p.tile, tile)
    return p


def reflection_params(cam, w, h, frame=0, spp=1, sun_tick=50.0, tile=(0, 0), inputs: SceneInputs = None, reproject=False,
                      temporal=False, lpv_gi=False, decoupled=False, ss_sky_valid=False) -> abi.ReflectionParams:
    p = abi.ReflectionParams()
    fill(p.inv_view, cam.inv_view); fill(p.inv_projection, cam.inv_projection)
    fill(p.view, cam.view); fill(p.projection, cam.projection)
    p.width, p.height = w, h
    p.spp, p.checkerboard, p.trace_length, p.shadow_trace_length = spp, 0, 64, 150
    p.current_frame, p.current_frame_mod128 = frame, frame % 128
    p.use_blue_noise, p.rough_reflections, p.roughness_bias, p.temporal = 1, 1, 0, int(temporal)
    p.reproject_to_screen_space, p.derive_from_diffuse_sh = int(reproject), 0
    p.lpv_gi, p.use_decoupled_gi, p.screen_space_skylighting_valid = int(lpv_gi), int(decoupled), int(ss_sky_valid)
    sun, moon, strong = host_api.sun_direction(sun_tick)
    fill(p.sun_direction, sun); fill(p.moon_direction, moon); fill(p.stronger_light_direction, strong)
    fill(p.viewer_position, cam.position)
    p.sun_strength_modifier, p.moon_strength_modifier = 0.85, 1.0
    if inputs is not None:
        fill(p.grass_props, inputs.grass)
    abi.set_tile(p.tile, tile)
    return p
