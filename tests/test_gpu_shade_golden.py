"""CUDA material fetch, Cook-Torrance direct term, diffuse GI and reflections vs tests/golden/shade_ref.npz — the outputs of the
reference's own shaders (GenerateGBuffer.glsl:522-578, ColorPassFrag.glsl:419-451, DiffuseRayTraceFrag.glsl:535-664,
ReflectionTraceFrag.glsl:717-1038) compiled by oracle/build_ref.py.  Through the C ABI.

Every pass is judged on the reference's own inputs: primary / hard-shadow / material outputs are bit-identical to the fixture
(asserted), and the reflection pass reads the fixture's GI attachments written into the context (vxrt_cuda_write_attachment), so
no CUDA-side deviation of an earlier pass leaks into a later one.

Tolerances (north_star: "shaded radiance within a stated tolerance when given identical blue-noise samples"): integer / table /
texture-fetch work is bit exact.  Radiance passes through powf / sinf / cosf, which differ by <= 2 ulp between CUDA and glibc and
can flip a grazing bounce ray, so: hit / miss decided identically on >= 99.9 % of pixels, R16F radiance within 1e-2 relative
(+1e-3 absolute) on >= 99.5 % of pixels, direct term within 2 half-ulps (2^-9 relative) on >= 99.9 % of pixels."""
import numpy as np
import pytest

import scene_util as su
import shade_golden_util as sg
from voxeltracing_b200 import abi, engine, host_api

pytestmark = pytest.mark.gpu
W, H = sg.W, sg.H


@pytest.fixture(scope="module")
def gold():
    return sg.golden()


@pytest.fixture(scope="module")
def inputs():
    return sg.inputs()


@pytest.fixture(scope="module")
def contexts(inputs):
    cs = {}
    for wname in sg.WORLDS:
        c = engine.Context(0)
        c.upload_world(sg.world(wname))
        c.generate_distance_field()
        inputs.apply_to_context(c)
        cs[wname] = c
    yield cs
    for c in cs.values():
        c.close()


def _close(got, want, rtol, atol):
    a, b = got.astype(np.float32), want.astype(np.float32)
    return np.abs(a - b) <= rtol * np.abs(b) + atol


def _front(c, case, gold, inputs):
    """primary + hard shadow + material fetch on the GPU; all bit-identical to the reference frame"""
    n = case["name"]
    cam = sg.camera(case)
    c.initial_trace(cam, W, H)
    c.shadow_trace(cam, W, H, host_api.sun_direction(50.0)[2], soft=False)
    assert sg.same_bits(c.read_attachment(abi.ATT_INITIAL_T), gold[f"{n}_g_t"])
    assert sg.same_bits(c.read_attachment(abi.ATT_INITIAL_NORMAL), gold[f"{n}_g_normal"])
    assert sg.same_bits(c.read_attachment(abi.ATT_INITIAL_BLOCK), gold[f"{n}_g_block"])
    assert sg.same_bits(c.read_attachment(abi.ATT_SHADOW), gold[f"{n}_shadow"])
    c.generate_gbuffer(su.gbuffer_params(cam, W, H, inputs))
    return cam


@pytest.mark.parametrize("case", sg.CASES, ids=lambda c: c["name"])
def test_material_fetch_bit_exact_vs_reference(case, contexts, gold, inputs):
    c = contexts[case["world"]]
    _front(c, case, gold, inputs)
    n = case["name"]
    for att, k in ((abi.ATT_GBUF_ALBEDO, "albedo"), (abi.ATT_GBUF_NORMAL, "normal"), (abi.ATT_GBUF_PBR, "pbr"), (abi.ATT_GBUF_TEXAO, "texao")):
        assert sg.same_bits(c.read_attachment(att), gold[f"{n}_gb_{k}"]), (n, k)


@pytest.mark.parametrize("case", [c for c in sg.CASES if c["sun_ticks"]], ids=lambda c: c["name"])
def test_direct_term_vs_reference(case, contexts, gold, inputs):
    c = contexts[case["world"]]
    cam = _front(c, case, gold, inputs)
    for tick in case["sun_ticks"]:
        c.shade_direct(su.direct_params(cam, W, H, tick))
        got, want = c.read_attachment(abi.ATT_DIRECT), gold[f"{case['name']}_direct_{int(tick)}"]
        ok = _close(got, want, 2.0 ** -9, 1e-6).all(axis=-1)
        assert ok.mean() >= 0.999, (tick, ok.mean())
        assert (got.view(np.uint16) == want.view(np.uint16)).mean() > 0.98


@pytest.mark.parametrize("case", sg.CASES, ids=lambda c: c["name"])
def test_diffuse_gi_vs_reference(case, contexts, gold, inputs):
    c = contexts[case["world"]]
    cam = _front(c, case, gold, inputs)
    for i, kw in enumerate(case["gi"]):
        c.diffuse_trace(su.gi_params(cam, W, H, **kw))
        got = {"sh": c.read_attachment(abi.ATT_GI_SH), "cocg": c.read_attachment(abi.ATT_GI_COCG),
               "utility": c.read_attachment(abi.ATT_GI_UTILITY), "aosky": c.read_attachment(abi.ATT_GI_AOSKY)}
        want = {k: gold[f"{case['name']}_gi{i}_{k}"] for k in sg.GI_KEYS}
        same_paths = (got["aosky"] == want["aosky"]).all(axis=-1)
        assert same_paths.mean() >= 0.999, (kw, same_paths.mean())
        for k in ("sh", "cocg", "utility"):
            ok = _close(got[k], want[k], 1e-2, 1e-3)
            ok = ok.all(axis=-1) if ok.ndim == 3 else ok
            assert ok.mean() >= 0.995, (kw, k, ok.mean())
        # most values come out with the very same bits
        assert (got["sh"].view(np.uint16) == want["sh"].view(np.uint16)).mean() > 0.9, kw


@pytest.mark.parametrize("case", sg.CASES, ids=lambda c: c["name"])
def test_reflections_vs_reference(case, contexts, gold, inputs):
    c = contexts[case["world"]]
    cam = _front(c, case, gold, inputs)
    n = case["name"]
    # size the GI attachments, then replace their contents by the reference's
    c.diffuse_trace(su.gi_params(cam, W, H, **case["gi"][0]))
    for att, k in ((abi.ATT_GI_SH, "sh"), (abi.ATT_GI_COCG, "cocg"), (abi.ATT_GI_UTILITY, "utility"), (abi.ATT_GI_AOSKY, "aosky")):
        c.write_attachment(att, gold[f"{n}_gi0_{k}"])
    for i, kw in enumerate(case["refl"]):
        c.reflection_trace(su.reflection_params(cam, W, H, inputs=inputs, **kw))
        got_c, got_h, got_e = c.read_attachment(abi.ATT_REFL_COLOR), c.read_attachment(abi.ATT_REFL_HITDIST), c.read_attachment(abi.ATT_REFL_EMISSIVE)
        want_c, want_h, want_e = (gold[f"{n}_refl{i}_{k}"] for k in sg.RF_KEYS)
        assert (got_e == want_e).mean() >= 0.999
        assert ((got_h.astype(np.float32) > 0) == (want_h.astype(np.float32) > 0)).mean() >= 0.999
        assert _close(got_h, want_h, 1e-2, 1e-2).mean() >= 0.995
        okc = _close(got_c, want_c, 1e-2, 1e-3).all(axis=-1)
        assert okc.mean() >= 0.995, (kw, okc.mean())
        assert (got_c.view(np.uint16) == want_c.view(np.uint16)).mean() > 0.9, kw


def test_reflections_with_lpv_gi_vs_reference(contexts, gold, inputs):
    """ApproximateGILPV inside the reflection shading (ReflectionTraceFrag.glsl:673-700,881-883): the CUDA pass with the propagation
    volume bound against the compiled shader's outputs with u_LPVGI on (plain / decoupled, with / without reprojection).  The volume
    and the average colours are uploaded from the CPU restatements, so the pass is judged on the fixture's exact inputs; the wavefront
    and the per-pixel kernels must agree bit for bit."""
    g = sg.golden_lpv()
    case = sg.LPV_CASE
    c = contexts[case["world"]]
    cam = _front(c, sg.CASES[0], gold, inputs)        # same camera as rooms_a
    n = case["name"]
    c.diffuse_trace(su.gi_params(cam, W, H, **case["gi"][0]))
    for att, k in ((abi.ATT_GI_SH, "sh"), (abi.ATT_GI_COCG, "cocg"), (abi.ATT_GI_UTILITY, "utility"), (abi.ATT_GI_AOSKY, "aosky")):
        c.write_attachment(att, gold[f"rooms_a_gi0_{k}"])
    rp0 = su.reflection_params(cam, W, H, inputs=inputs, **case["refl"][0])
    with pytest.raises(engine.VxrtError):
        c.reflection_trace(rp0)                        # lpv_gi without a volume
    level, btype, avg, _ = sg.lpv_inputs(sg.world(case["world"]), inputs, case["lpv_limit"])
    c.lpv_upload(level, btype)
    c.lpv_set_average_colors(avg)
    for i, kw in enumerate(case["refl"]):
        rp = su.reflection_params(cam, W, H, inputs=inputs, **kw)
        outs = {}
        for mode in (1, 0):
            c.set_option("wavefront", mode)
            c.reflection_trace(rp)
            outs[mode] = [c.read_attachment(a).copy() for a in (abi.ATT_REFL_COLOR, abi.ATT_REFL_HITDIST, abi.ATT_REFL_EMISSIVE)]
        c.set_option("wavefront", 1)
        for a, b in zip(outs[0], outs[1]):
            assert sg.same_bits(a, b), kw
        got_c, got_h, got_e = outs[1]
        want_c, want_h, want_e = (g[f"{n}_refl{i}_{k}"] for k in sg.RF_KEYS)
        assert (got_e == want_e).mean() >= 0.999
        assert ((got_h.astype(np.float32) > 0) == (want_h.astype(np.float32) > 0)).mean() >= 0.999
        okc = _close(got_c, want_c, 1e-2, 1e-3).all(axis=-1)
        assert okc.mean() >= 0.995, (kw, okc.mean())
        assert (got_c.view(np.uint16) == want_c.view(np.uint16)).mean() > 0.9, kw
