"""Cases of tests/golden/shade_ref.npz (material fetch, Cook-Torrance direct term, diffuse GI, reflections).

One description of the frames, three executors: oracle/_ref (the reference's shaders, make_golden_shade.py), the oracle restatement
(tests/test_oracle_shade.py) and the CUDA path (tests/test_gpu_shade_golden.py).  Each pass of a case consumes the outputs of the
passes before it from the SAME executor (`run_case`), or — for the CUDA tests — from the fixture, so a pass is judged on identical
inputs.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import scene_util as su  # noqa: E402
from voxeltracing_b200 import abi, host_api  # noqa: E402

GOLD = ROOT / "tests" / "golden" / "shade_ref.npz"
GOLD_LPV = ROOT / "tests" / "golden" / "shade_lpv_ref.npz"
W, H = 160, 90
TEX = 64
WORLDS = ("rooms2", "plains1")

CASES = [
    dict(name="rooms_a", world="rooms2", pos=[200, 58, 200], yaw=30.0, pitch=-15.0, sun_ticks=(50.0,),
         gi=[dict(frame=0, spp=1, checkerboard=False)], refl=[dict(frame=3, spp=1)]),
    dict(name="plains_b", world="plains1", pos=[192, 80, 192], yaw=30.0, pitch=-15.0, sun_ticks=(50.0, 130.0),
         gi=[dict(frame=5, spp=2, checkerboard=False)], refl=[dict(frame=3, spp=2)]),
    dict(name="rooms_c", world="rooms2", pos=[150.5, 60.2, 221.3], yaw=200.0, pitch=5.0, sun_ticks=(),
         gi=[dict(frame=9, spp=3, checkerboard=True), dict(frame=2, spp=4, checkerboard=False)],
         refl=[dict(frame=9, spp=2, reproject=True, temporal=True)]),
]

# ApproximateGILPV inside the reflection pass (ReflectionTraceFrag.glsl:673-700,881-883; u_LPVGI is on by default in the engine): the same
# frame with the propagation volume of the world's lamps bound, in the plain and the decoupled form, with and without screen-space
# reprojection (the LPV term only replaces the ambient term where the reprojection fails)
LPV_CASE = dict(name="rooms_lpv", world="rooms2", pos=[200, 58, 200], yaw=30.0, pitch=-15.0, sun_ticks=(), lpv_limit=8,
                gi=[dict(frame=0, spp=1, checkerboard=False)],
                refl=[dict(frame=3, spp=1, lpv_gi=True), dict(frame=6, spp=2, lpv_gi=True, temporal=True, decoupled=True, ss_sky_valid=True),
                      dict(frame=2, spp=1, lpv_gi=True, decoupled=True), dict(frame=9, spp=2, lpv_gi=True, reproject=True, temporal=True)])

GB_KEYS = ("albedo", "normal", "pbr", "texao")
GI_KEYS = ("sh", "cocg", "utility", "aosky")
RF_KEYS = ("color", "hitdist", "emissive")


def world(name: str) -> np.ndarray:
    kind, seed = {"rooms2": ("rooms", 2), "plains1": ("plains", 1)}[name]
    return host_api.gen_world(kind, seed)


def inputs() -> su.SceneInputs:
    return su.SceneInputs(TEX)


def golden():
    return np.load(GOLD)


def golden_lpv():
    return np.load(GOLD_LPV)


def camera(case):
    return host_api.camera(case["pos"], case["yaw"], case["pitch"], W / H)


def primary_params(cam) -> abi.PrimaryParams:
    p = abi.PrimaryParams()
    su.fill(p.inv_view, cam.inv_view); su.fill(p.inv_projection, cam.inv_projection)
    p.width, p.height, p.render_distance = W, H, 350
    return p


def shadow_params(cam) -> abi.ShadowParams:
    s = abi.ShadowParams()
    su.fill(s.inv_view, cam.inv_view); su.fill(s.inv_projection, cam.inv_projection)
    s.width, s.height = W, H
    su.fill(s.light_direction, host_api.sun_direction(50.0)[2])
    s.current_frame, s.soft_shadows, s.max_iterations = 0, 0, 350
    return s


def lpv_inputs(blocks, inp, limit):
    """(light level, block type, BlockAverageColorData) of the world's lamps from the CPU restatements (pinned against the reference's own
    VolumetricFloodFill.cpp / PrecomputeAverageBlockColor.comp by tests/test_oracle_lpv.py)."""
    from oracle import binding as ob
    from oracle import world_binding as wb

    lights = wb.collect_lights(blocks, inp.table)
    level, btype = wb.lpv_repropagate(blocks, lights, limit)
    sc = ob.OracleScene(ob.OracleWorld(blocks, np.zeros_like(blocks)))
    inp.apply_to_oracle(sc)
    return level, btype, sc.lpv_average_colors(), lights


def run_case(be, case, inp) -> dict:
    """All passes of a case through one executor `be` (methods initial_trace, shadow_trace, generate_gbuffer, shade_direct,
    diffuse_trace, reflection_trace with the signatures of oracle.ref_binding).  Returns {key: array} without the case prefix."""
    cam = camera(case)
    g = be.initial_trace(primary_params(cam))
    sh = be.shadow_trace(shadow_params(cam), g["t"], g["normal"])
    out = {"g_t": g["t"], "g_normal": g["normal"], "g_block": g["block"], "shadow": sh["shadow"]}
    gb = be.generate_gbuffer(su.gbuffer_params(cam, W, H, inp), g["inv_t"], g["normal"], g["block"])
    for k in GB_KEYS:
        out[f"gb_{k}"] = gb[k]
    for tick in case["sun_ticks"]:
        out[f"direct_{int(tick)}"] = be.shade_direct(su.direct_params(cam, W, H, tick), g["inv_t"], gb, sh["shadow"])
    gis = []
    for i, kw in enumerate(case["gi"]):
        gi = be.diffuse_trace(su.gi_params(cam, W, H, **kw), g["t"], g["normal"])
        gis.append(gi)
        for k in GI_KEYS:
            out[f"gi{i}_{k}"] = gi[k]
    if "lpv_limit" in case:
        level, btype, avg, _ = lpv_inputs(be.blocks, inp, case["lpv_limit"])
        be.set_lpv(level, btype, avg)
    for i, kw in enumerate(case["refl"]):
        rf = be.reflection_trace(su.reflection_params(cam, W, H, inputs=inp, **kw), g["t"], g["normal"], gb, gis[0], sh["shadow"])
        for k in RF_KEYS:
            out[f"refl{i}_{k}"] = rf[k]
    return out


def same_bits(a: np.ndarray, b: np.ndarray) -> bool:
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))
