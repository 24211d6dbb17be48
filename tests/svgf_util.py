"""Two-frame inputs for the SVGF tests: G-buffers and raw GI outputs of two nearby camera poses in the rooms world,
produced by the oracle (CPU) so the same arrays feed the oracle, oracle/_ref and the CUDA passes."""
import numpy as np

import scene_util as su
from oracle import binding as ob
from voxeltracing_b200 import abi, host_api

W, H = 192, 108
POSES = [([192.0, 62.0, 192.0], 30.0, -15.0), ([192.4, 62.0, 191.7], 33.0, -14.0), ([192.7, 62.1, 191.5], 35.0, -14.0)]   # small camera motion


def fill(dst, src):
    for i, v in enumerate(np.asarray(src, np.float32).ravel()):
        dst[i] = float(v)


def frames(world_blocks, inputs=None):
    """Returns [{"cam", "g": {t, normal, block}, "raw": {sh, cocg, x (luminance), aosky}} per frame]."""
    inputs = inputs or su.SceneInputs(64)
    ow = ob.OracleWorld(world_blocks)
    sc = ob.OracleScene(ow)
    inputs.apply_to_oracle(sc)
    out = []
    for f, (pos, yaw, pitch) in enumerate(POSES):
        cam = host_api.camera(pos, yaw, pitch, W / H)
        p = abi.PrimaryParams()
        fill(p.inv_view, cam.inv_view); fill(p.inv_projection, cam.inv_projection)
        p.width, p.height, p.render_distance = W, H, 350
        g = ow.initial_trace(p)
        gi = sc.diffuse_trace(su.gi_params(cam, W, H, frame=f, spp=1), g["t"], g["normal"])
        raw = {"sh": np.ascontiguousarray(gi["sh"]), "cocg": np.ascontiguousarray(gi["cocg"]), "x": np.ascontiguousarray(gi["utility"]),
               "aosky": np.ascontiguousarray(gi["aosky"])}
        out.append({"cam": cam, "g": {k: np.ascontiguousarray(g[k]) for k in ("t", "normal", "block")}, "raw": raw})
    return out


def temporal_params(cam, prev_cam, in_set, history_set, out_set, be_useful=True) -> abi.SvgfTemporalParams:
    p = abi.SvgfTemporalParams()
    fill(p.inv_view, cam.inv_view); fill(p.inv_projection, cam.inv_projection)
    fill(p.prev_view, prev_cam.view); fill(p.prev_projection, prev_cam.projection)
    p.width, p.height = W, H
    p.in_set, p.history_set, p.out_set, p.be_useful = in_set, history_set, out_set, int(be_useful)
    return p


def zero_gbuf():
    return {"t": np.zeros((H, W), np.float16), "normal": np.zeros((H, W), np.uint8), "block": np.zeros((H, W), np.uint8)}


def prespatial_params(cam, in_set=abi.ATT_GI_SH, time=0.0) -> abi.SvgfPreSpatialParams:
    p = abi.SvgfPreSpatialParams()
    fill(p.inv_view, cam.inv_view); fill(p.inv_projection, cam.inv_projection)
    p.width, p.height, p.in_set, p.time = W, H, in_set, time
    return p


def variance_params(cam, in_set, do_spatial=True, aggressive=True) -> abi.SvgfVarianceParams:
    p = abi.SvgfVarianceParams()
    fill(p.inv_view, cam.inv_view); fill(p.inv_projection, cam.inv_projection)
    p.width, p.height, p.in_set, p.do_spatial, p.aggressive_disocclusion = W, H, in_set, int(do_spatial), int(aggressive)
    return p


def spatial_params(cam, in_set, ao_set, temporal_set, out_set, step, time=0.0, large=False, do_spatial=True, aggressive=True,
                   phi_bias=2.8, res_scale=0.25) -> abi.SvgfSpatialParams:
    p = abi.SvgfSpatialParams()
    fill(p.inv_view, cam.inv_view); fill(p.inv_projection, cam.inv_projection)
    p.width, p.height = W, H
    p.in_set, p.ao_set, p.temporal_set, p.out_set, p.step = in_set, ao_set, temporal_set, out_set, step
    p.large_kernel, p.do_spatial, p.aggressive_disocclusion = int(large), int(do_spatial), int(aggressive)
    p.color_phi_bias, p.time, p.resolution_scale = phi_bias, time, res_scale
    return p


STEPS = (16, 8, 4, 2, 1)   # Pipeline.cpp:2583-2589 (WiderSVGF off)


def spatial_chain(cam, temporal, variance, g, spatial_fn, time=1.25, **kw):
    """The five a-trous iterations of Pipeline.cpp:2592-2700: VARIANCE -> DENOISE_A -> DENOISE_B -> A -> B -> A.
    spatial_fn(params, prev_set, ao_image, temporal_utility, g) -> set.  Returns the list of the five outputs."""
    outs, prev, ao = [], variance, temporal["aosky"]
    for i, step in enumerate(STEPS):
        cur_id = abi.ATT_SVGF_DENOISE_A if i % 2 == 0 else abi.ATT_SVGF_DENOISE_B
        prev_id = abi.ATT_SVGF_VARIANCE if i == 0 else (abi.ATT_SVGF_DENOISE_B if i % 2 == 0 else abi.ATT_SVGF_DENOISE_A)
        p = spatial_params(cam, prev_id, abi.ATT_SVGF_TEMPORAL_A if i == 0 else prev_id, abi.ATT_SVGF_TEMPORAL_A, cur_id, step, time=time, **kw)
        out = spatial_fn(p, prev, ao, temporal["x"], g)
        outs.append(out)
        prev, ao = out, out["aosky"]
    return outs


def same_bits(a, b, keys=("sh", "cocg", "x", "aosky")):
    """bit equality with NaNs canonicalised (x86 and the GPU produce different quiet NaNs)"""
    for k in keys:
        x, y = a[k], b[k]
        if x.dtype == np.float16:
            nx, ny = np.isnan(x), np.isnan(y)
            if not (np.array_equal(nx, ny) and np.array_equal(x.view(np.uint16)[~nx], y.view(np.uint16)[~ny])):
                return False
        elif not np.array_equal(x, y):
            return False
    return True
