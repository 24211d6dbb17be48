"""CUDA material fetch, Cook-Torrance direct, diffuse GI and reflection passes vs the oracle, through the
C ABI (BASELINE configs 3 and 4).  Integer / table / texture-fetch work is bit exact; radiance that
passes through powf / sinf / cosf (CUDA vs libm differ by <= 2 ulp, which can flip a grazing ray) is held
to the tolerances written below."""
import numpy as np
import pytest

import scene_util as su
from oracle import binding as ob
from voxeltracing_b200 import abi, engine, host_api

pytestmark = pytest.mark.gpu

W, H = 320, 180


@pytest.fixture(scope="module")
def inputs():
    return su.SceneInputs(128)


def _make(world_kind, seed, inputs):
    blocks = host_api.gen_world(world_kind, seed)
    ow = ob.OracleWorld(blocks)
    sc = ob.OracleScene(ow)
    inputs.apply_to_oracle(sc)
    c = engine.Context(0)
    c.upload_world(blocks)
    c.generate_distance_field()
    inputs.apply_to_context(c)
    return c, ow, sc


@pytest.fixture(scope="module")
def rooms(inputs):
    c, ow, sc = _make("rooms", 2, inputs)
    yield c, ow, sc
    c.close()


@pytest.fixture(scope="module")
def plains(inputs):
    c, ow, sc = _make("plains", 1, inputs)
    yield c, ow, sc
    c.close()


def _primary_and_shadow(c, ow, cam):
    p = c.initial_trace(cam, W, H)
    g = ow.initial_trace(p)
    light = host_api.sun_direction(50.0)[2]
    sp = c.shadow_trace(cam, W, H, light, soft=False)
    sh = ow.shadow_trace(sp, g["t"], g["normal"], None)
    assert np.array_equal(c.read_attachment(abi.ATT_INITIAL_T).view(np.uint16), g["t"].view(np.uint16))
    assert np.array_equal(c.read_attachment(abi.ATT_SHADOW), sh["shadow"])
    return g, sh


def _close(got, want, rtol, atol):
    a, b = got.astype(np.float32), want.astype(np.float32)
    return np.abs(a - b) <= rtol * np.abs(b) + atol


def test_mip_chain_matches_oracle_model(rooms, inputs):
    """glGenerateMipmap model: both sides must build byte-identical levels (checked through sampling:
    a GI-style textureLod at integer lods over a grid of uvs)."""
    c, ow, sc = rooms
    lv = sc.texture_level(abi.TEX_ALBEDO, 3, inputs.textures[0].shape[0], 128)
    assert lv.shape[1] == 16 and lv.std() > 1


@pytest.mark.parametrize("world,pos,yaw,pitch", [("plains", [192, 80, 192], 30.0, -15.0), ("rooms", [200, 58, 200], 30.0, -15.0),
                                                 ("rooms", [150.5, 60.2, 221.3], 200.0, 5.0)])
def test_generate_gbuffer_bit_exact(request, inputs, world, pos, yaw, pitch):
    c, ow, sc = request.getfixturevalue(world)
    cam = host_api.camera(pos, yaw, pitch, W / H)
    g, _ = _primary_and_shadow(c, ow, cam)
    gp = su.gbuffer_params(cam, W, H, inputs)
    c.generate_gbuffer(gp)
    want = sc.generate_gbuffer(gp, g["inv_t"], g["normal"], g["block"])
    assert np.array_equal(c.read_attachment(abi.ATT_GBUF_ALBEDO).view(np.uint16), want["albedo"].view(np.uint16))
    assert np.array_equal(c.read_attachment(abi.ATT_GBUF_NORMAL).view(np.uint16), want["normal"].view(np.uint16))
    assert np.array_equal(c.read_attachment(abi.ATT_GBUF_PBR), want["pbr"])
    assert np.array_equal(c.read_attachment(abi.ATT_GBUF_TEXAO), want["texao"])
    assert want["albedo"].astype(np.float32).std() > 0.01


@pytest.mark.parametrize("world,pos,sun_tick", [("plains", [192, 80, 192], 50.0), ("rooms", [200, 58, 200], 50.0), ("plains", [192, 80, 192], 130.0)])
def test_config3_cook_torrance_direct(request, inputs, world, pos, sun_tick):
    c, ow, sc = request.getfixturevalue(world)
    cam = host_api.camera(pos, 30.0, -15.0, W / H)
    g, sh = _primary_and_shadow(c, ow, cam)
    gp = su.gbuffer_params(cam, W, H, inputs)
    c.generate_gbuffer(gp)
    gb = sc.generate_gbuffer(gp, g["inv_t"], g["normal"], g["block"])
    dp = su.direct_params(cam, W, H, sun_tick)
    c.shade_direct(dp)
    got = c.read_attachment(abi.ATT_DIRECT)
    want = sc.shade_direct(dp, g["inv_t"], gb, sh["shadow"])
    # one powf (Fresnel) per light: stated tolerance 2 half-ulps (2^-9 relative) on >= 99.9 % of pixels
    ok = _close(got, want, 2.0 ** -9, 1e-6).all(axis=-1)
    assert ok.mean() >= 0.999, ok.mean()
    assert (got.view(np.uint16) == want.view(np.uint16)).mean() > 0.98
    if sun_tick == 50.0 and world == "plains":
        assert want.astype(np.float32).max() > 0.05


@pytest.mark.parametrize("world,pos,frame,spp,checker", [("rooms", [200, 58, 200], 0, 1, False), ("plains", [192, 80, 192], 5, 2, False),
                                                         ("rooms", [150.5, 60.2, 221.3], 9, 3, True)])
def test_config4_diffuse_gi(request, inputs, world, pos, frame, spp, checker):
    c, ow, sc = request.getfixturevalue(world)
    cam = host_api.camera(pos, 30.0, -15.0, W / H)
    g, _ = _primary_and_shadow(c, ow, cam)
    ip = su.gi_params(cam, W, H, frame=frame, spp=spp, checkerboard=checker)
    c.stats_enable(True); c.stats_read(True)
    c.diffuse_trace(ip)
    stats = c.stats_read(True); c.stats_enable(False)
    want = sc.diffuse_trace(ip, g["t"], g["normal"])
    got = {"sh": c.read_attachment(abi.ATT_GI_SH), "cocg": c.read_attachment(abi.ATT_GI_COCG),
           "utility": c.read_attachment(abi.ATT_GI_UTILITY), "aosky": c.read_attachment(abi.ATT_GI_AOSKY)}
    # sky-hit fraction and AO are decided by hit / miss of each path: identical on >= 99.9 % of pixels
    same_paths = (got["aosky"] == want["aosky"]).all(axis=-1)
    assert same_paths.mean() >= 0.999, same_paths.mean()
    # radiance: cos/sin/pow differ by <= 2 ulp between CUDA and libm; stated tolerance 1e-2 relative
    # (+1e-3 absolute) on the R16F outputs, on >= 99.5 % of pixels
    for k in ("sh", "cocg"):
        ok = _close(got[k], want[k], 1e-2, 1e-3).all(axis=-1)
        assert ok.mean() >= 0.995, (k, ok.mean())
    ok = _close(got["utility"], want["utility"], 1e-2, 1e-3)
    assert ok.mean() >= 0.995
    # the ray counts agree to 0.1 % (a flipped grazing ray changes how many shadow rays follow)
    assert abs(stats["rays"] - want["stats"]["rays"]) <= 1e-3 * want["stats"]["rays"]
    assert want["sh"].astype(np.float32).std() > 1e-3


@pytest.mark.parametrize("world,pos,kw", [("rooms", [200, 58, 200], dict(frame=3, spp=1)), ("plains", [192, 80, 192], dict(frame=3, spp=2)),
                                         ("rooms", [200, 58, 200], dict(frame=9, spp=2, reproject=True, temporal=True))])
def test_config4_reflections(request, inputs, world, pos, kw):
    c, ow, sc = request.getfixturevalue(world)
    cam = host_api.camera(pos, 30.0, -15.0, W / H)
    g, sh = _primary_and_shadow(c, ow, cam)
    gp = su.gbuffer_params(cam, W, H, inputs)
    c.generate_gbuffer(gp)
    gb = sc.generate_gbuffer(gp, g["inv_t"], g["normal"], g["block"])
    ip = su.gi_params(cam, W, H, frame=kw["frame"], spp=1)
    c.diffuse_trace(ip)
    # feed the oracle the CUDA GI attachments so only the reflection pass is under test
    gi = {"sh": c.read_attachment(abi.ATT_GI_SH), "cocg": c.read_attachment(abi.ATT_GI_COCG),
          "utility": c.read_attachment(abi.ATT_GI_UTILITY), "aosky": c.read_attachment(abi.ATT_GI_AOSKY)}
    rp = su.reflection_params(cam, W, H, inputs=inputs, **kw)
    c.reflection_trace(rp)
    want = sc.reflection_trace(rp, g["t"], g["normal"], gb, gi, sh["shadow"])
    got_c, got_h, got_e = c.read_attachment(abi.ATT_REFL_COLOR), c.read_attachment(abi.ATT_REFL_HITDIST), c.read_attachment(abi.ATT_REFL_EMISSIVE)
    same_mask = got_e == want["emissive"]
    assert same_mask.mean() >= 0.999
    hit_same = (got_h.astype(np.float32) > 0) == (want["hitdist"].astype(np.float32) > 0)
    assert hit_same.mean() >= 0.999
    okh = _close(got_h, want["hitdist"], 1e-2, 1e-2)
    assert okh.mean() >= 0.995, okh.mean()
    okc = _close(got_c, want["color"], 1e-2, 1e-3).all(axis=-1)
    assert okc.mean() >= 0.995, okc.mean()


def test_pass_order_errors(inputs):
    c = engine.Context(0)
    c.upload_world(host_api.gen_world("flat", 0))
    c.generate_distance_field()
    cam = host_api.camera([192, 80, 192], 0.0, -30.0, W / H)
    with pytest.raises(engine.VxrtError):
        c.generate_gbuffer(su.gbuffer_params(cam, W, H, inputs))      # no primary pass yet
    c.initial_trace(cam, W, H)
    with pytest.raises(engine.VxrtError):
        c.generate_gbuffer(su.gbuffer_params(cam, W, H, inputs))      # no texture arrays
    with pytest.raises(engine.VxrtError):
        c.diffuse_trace(su.gi_params(cam, W, H))                      # no tables / sky
    inputs.apply_to_context(c)
    bad = su.gi_params(cam, W, H)
    bad.use_blue_noise = 0
    with pytest.raises(engine.VxrtError):
        c.diffuse_trace(bad)
    with pytest.raises(engine.VxrtError):
        c.set_texture_array(0, np.zeros((2, 100, 100, 4), np.uint8))  # not a power of two
    c.generate_gbuffer(su.gbuffer_params(cam, W, H, inputs))
    c.diffuse_trace(su.gi_params(cam, W, H))
    c.close()


@pytest.mark.parametrize("world,pos,frame,spp,checker", [("rooms", [200, 58, 200], 4, 3, True), ("plains", [192, 80, 192], 1, 2, False)])
def test_wavefront_gi_is_bit_identical_to_the_per_pixel_kernel(request, inputs, world, pos, frame, spp, checker):
    c, ow, sc = request.getfixturevalue(world)
    cam = host_api.camera(pos, 75.0, -12.0, W / H)
    c.initial_trace(cam, W, H)
    ip = su.gi_params(cam, W, H, frame=frame, spp=spp, checkerboard=checker)
    atts = (abi.ATT_GI_SH, abi.ATT_GI_COCG, abi.ATT_GI_UTILITY, abi.ATT_GI_AOSKY)
    res = {}
    for mode in (0, 1):
        c.set_option("wavefront", mode)
        c.stats_enable(True); c.stats_read(True)
        c.diffuse_trace(ip)
        res[mode] = ([c.read_attachment(a).copy() for a in atts], c.stats_read(True))
        c.stats_enable(False)
    c.set_option("wavefront", 1)
    for a, b in zip(res[0][0], res[1][0]):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert res[0][1] == res[1][1]   # same rays, iterations, DDA steps, hits
    with pytest.raises(engine.VxrtError):
        c.set_option("no-such-option", 1)


@pytest.mark.parametrize("world,pos,kw", [("rooms", [200, 58, 200], dict(frame=9, spp=5, reproject=True, temporal=True)),
                                         ("plains", [192, 80, 192], dict(frame=2, spp=2)), ("rooms", [150.5, 60.2, 221.3], dict(frame=0, spp=1))])
def test_wavefront_reflections_are_bit_identical_to_the_per_pixel_kernel(request, inputs, world, pos, kw):
    c, ow, sc = request.getfixturevalue(world)
    cam = host_api.camera(pos, 110.0, -8.0, W / H)
    c.initial_trace(cam, W, H)
    c.shadow_trace(cam, W, H, host_api.sun_direction(50.0)[2], soft=False)
    c.generate_gbuffer(su.gbuffer_params(cam, W, H, inputs))
    c.diffuse_trace(su.gi_params(cam, W, H, frame=kw["frame"], spp=1))
    rp = su.reflection_params(cam, W, H, inputs=inputs, **kw)
    if kw["spp"] == 5:
        rp.checkerboard = 1
        rp.derive_from_diffuse_sh = 1
    atts = (abi.ATT_REFL_COLOR, abi.ATT_REFL_HITDIST, abi.ATT_REFL_EMISSIVE)
    res = {}
    # per-pixel kernel, the wavefront with the last sample's shade_b fused with resolve (default), the wavefront with the separate kernels
    for key, (mode, fuse) in enumerate(((0, 1), (1, 1), (1, 0))):
        c.set_option("wavefront", mode); c.set_option("gi_fuse_final", fuse)
        c.stats_enable(True); c.stats_read(True)
        c.reflection_trace(rp)
        res[key] = ([c.read_attachment(a).copy() for a in atts], c.stats_read(True))
        c.stats_enable(False)
    c.set_option("wavefront", 1); c.set_option("gi_fuse_final", 1)
    for other in (1, 2):
        for a, b in zip(res[0][0], res[other][0]):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
        assert res[0][1] == res[other][1]


@pytest.mark.parametrize("caps", [(), (12, 24), (1, 2, 3), (5,), (47,), (30, 100), (200, 250)])
def test_capped_trace_passes_do_not_change_a_bit(rooms, inputs, caps):
    """Iteration-capped passes with compaction between launches (trace_queue.cuh, set_option "trace_caps"): every cap schedule - none,
    the default, caps of one iteration, a cap next to the GI trace length (48), caps beyond the reflection trace length (64) but inside
    the shadow lengths (128 / 150), caps beyond every length - gives the attachments and the traversal statistics of the uncapped kernels."""
    c, ow, sc = rooms
    cam = host_api.camera([188.3, 61.0, 172.9], 115.0, -8.0, W / H)
    c.initial_trace(cam, W, H)
    c.shadow_trace(cam, W, H, host_api.sun_direction(50.0)[2], soft=False)
    c.generate_gbuffer(su.gbuffer_params(cam, W, H, inputs))
    ip = su.gi_params(cam, W, H, frame=6, spp=2, checkerboard=True)
    rp = su.reflection_params(cam, W, H, inputs=inputs, frame=6, spp=2)
    atts = (abi.ATT_GI_SH, abi.ATT_GI_COCG, abi.ATT_GI_UTILITY, abi.ATT_GI_AOSKY, abi.ATT_REFL_COLOR, abi.ATT_REFL_HITDIST, abi.ATT_REFL_EMISSIVE)

    def run(packed):
        c.set_option("trace_caps", packed)
        c.stats_enable(True); c.stats_read(True)
        c.diffuse_trace(ip)
        c.reflection_trace(rp)
        st = c.stats_read(True); c.stats_enable(False)
        return [c.read_attachment(a).copy() for a in atts], st

    try:
        want, want_st = run(0)
        got, got_st = run(sum(k << (8 * j) for j, k in enumerate(caps)))
    finally:
        c.set_option("trace_caps", 0)
    for a, b in zip(got, want):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert got_st == want_st
    assert want_st["rays"] > 100000


@pytest.mark.parametrize("thresholds", [(8,), (12, 8), (16, 12, 8), (1,), (31,), (31, 31, 31)])
def test_adaptive_hand_over_does_not_change_a_bit(rooms, inputs, thresholds):
    """Adaptive hand-over of the queue trace kernels (trace_queue.cuh, set_option "trace_spill"): a warp appends its stragglers to a
    continuation queue once at most T lanes still have a ray and the next launch packs them.  Every threshold schedule - the measured
    ones, a warp that only gives up its last ray, warps that hand over everything at once (T = 31: all rays travel through every
    pass) - gives the attachments and the traversal statistics of the plain kernels."""
    c, ow, sc = rooms
    cam = host_api.camera([188.3, 61.0, 172.9], 115.0, -8.0, W / H)
    c.initial_trace(cam, W, H)
    c.shadow_trace(cam, W, H, host_api.sun_direction(50.0)[2], soft=False)
    c.generate_gbuffer(su.gbuffer_params(cam, W, H, inputs))
    ip = su.gi_params(cam, W, H, frame=6, spp=2, checkerboard=True)
    rp = su.reflection_params(cam, W, H, inputs=inputs, frame=6, spp=2)
    atts = (abi.ATT_GI_SH, abi.ATT_GI_COCG, abi.ATT_GI_UTILITY, abi.ATT_GI_AOSKY, abi.ATT_REFL_COLOR, abi.ATT_REFL_HITDIST, abi.ATT_REFL_EMISSIVE)

    def run(packed):
        c.set_option("trace_spill", packed)
        c.stats_enable(True); c.stats_read(True)
        c.diffuse_trace(ip)
        c.reflection_trace(rp)
        st = c.stats_read(True); c.stats_enable(False)
        return [c.read_attachment(a).copy() for a in atts], st

    default = c.get_option("trace_spill") if hasattr(c, "get_option") else None
    try:
        want, want_st = run(0)
        got, got_st = run(sum(k << (8 * j) for j, k in enumerate(thresholds)))
    finally:
        c.set_option("trace_spill", default if default is not None else 0)
    for a, b in zip(got, want):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert got_st == want_st
    assert want_st["rays"] > 100000


@pytest.mark.parametrize("world,pos,frame,spp,checker", [("rooms", [200, 58, 200], 4, 1, False), ("rooms", [200, 58, 200], 7, 3, True),
                                                         ("plains", [192, 80, 192], 1, 4, False)])
def test_fused_final_gi_kernel_is_bit_identical(request, inputs, world, pos, frame, spp, checker):
    """The last sample's shade<2> fused with resolve (gi_wf_final_kernel) and the bounce-0 terms carried in contrib / thr give the bits
    of the separate kernels and of the one-thread-per-pixel kernel, at 1 spp (accumulators of live paths never touched), with a
    checkerboard (pixels whose last sample is not the frame's last) and at 4 spp."""
    c, ow, sc = request.getfixturevalue(world)
    cam = host_api.camera(pos, 75.0, -12.0, W / H)
    c.initial_trace(cam, W, H)
    ip = su.gi_params(cam, W, H, frame=frame, spp=spp, checkerboard=checker)
    atts = (abi.ATT_GI_SH, abi.ATT_GI_COCG, abi.ATT_GI_UTILITY, abi.ATT_GI_AOSKY)
    res = []
    try:
        # the default pipeline (shadow-queue trace on a side stream beside the bounce trace) first, then each simplification undone
        for wavefront, fuse, overlap in ((1, 1, 1), (1, 1, 0), (1, 0, 1), (1, 0, 0), (0, 0, 0)):
            c.set_option("wavefront", wavefront); c.set_option("gi_fuse_final", fuse); c.set_option("gi_overlap", overlap)
            for _ in range(2 if overlap else 1):   # twice: a race between the two streams would not repeat itself
                c.diffuse_trace(ip)
                res.append([c.read_attachment(a).copy() for a in atts])
    finally:
        c.set_option("wavefront", 1); c.set_option("gi_fuse_final", 1); c.set_option("gi_overlap", 1)
    for other in res[1:]:
        for a, b in zip(res[0], other):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert res[0][0].astype(np.float32).std() > 1e-3
