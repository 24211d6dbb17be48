"""SVGF chain of the diffuse GI: oracle self-checks and oracle == the reference's own shaders compiled
for the CPU (oracle/_ref) on a two-frame sequence.  CPU only."""
import ctypes as C

import numpy as np
import pytest

import svgf_util as sv
from oracle import binding as ob
from oracle import ref_binding as rb
from voxeltracing_b200 import abi, host_api


@pytest.fixture(scope="module")
def seq():
    return sv.frames(host_api.gen_world("plains", 0))


def _same(a, b):
    return all(np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)) for k in ("sh", "cocg", "x", "aosky"))


def _temporal_frames(seq, fn=None):
    """frame 0 against an all-zero history, frame k against frame k-1's output"""
    hist = ob.svgf_alloc(sv.H, sv.W, 3)
    prev_g, prev_cam = sv.zero_gbuf(), seq[0]["cam"]
    outs = []
    for f in seq:
        p = sv.temporal_params(f["cam"], prev_cam, abi.ATT_GI_SH, abi.ATT_SVGF_TEMPORAL_B, abi.ATT_SVGF_TEMPORAL_A)
        out = ob.svgf_temporal(p, f["raw"], hist, f["g"], prev_g, fn)
        outs.append(out)
        hist, prev_g, prev_cam = out, f["g"], f["cam"]
    return outs


def test_temporal_accumulates_where_the_history_is_valid(seq):
    o0, o1, o2 = _temporal_frames(seq)
    # frame 0: an all-zero history reconstructs every previous position at the camera, so only surfaces within the
    # position tolerance of the eye can match it: the rest passes through with zero accumulated frames
    fresh = o0["x"][..., 0] == 0
    assert fresh.mean() > 0.9
    # (up to the bilinear weights of texture() at a pixel centre, which are not exactly 0 / 1 in float)
    assert np.allclose(o0["sh"].astype(np.float32)[fresh], seq[0]["raw"]["sh"].astype(np.float32)[fresh], atol=2e-3)
    assert (np.abs(seq[0]["raw"]["sh"].astype(np.float32)) > 0.01).mean() > 0.3
    # frame 1: pixels that reproject onto frame 0 count one accumulated frame (blend factor still 1); frame 2: two, and
    # the output is a blend of history and current sample
    acc1, acc2 = o1["x"][..., 0].astype(np.float32), o2["x"][..., 0].astype(np.float32)
    assert (acc1 >= 0.99).mean() > 0.25 and (acc1 == 0).mean() > 0.005
    assert (acc2 >= 1.9).mean() > 0.2
    moved = np.abs(o2["sh"].astype(np.float32) - seq[2]["raw"]["sh"].astype(np.float32)).max(-1) > 1e-3
    assert moved[acc2 >= 1.9].mean() > 0.5 and moved[acc2 == 0].mean() < 0.05
    # second moment and luminance history are carried
    assert np.isfinite(o2["x"].astype(np.float32)).all() and (o2["x"][..., 1].astype(np.float32) >= 0).all()


@pytest.mark.skipif(not rb.available("svgf_temporal"), reason="oracle/_ref not built on this box")
@pytest.mark.parametrize("be_useful", [True, False])
def test_oracle_temporal_equals_compiled_reference_shader(seq, be_useful):
    L = rb.lib()
    a = _temporal_frames(seq)
    b = _temporal_frames(seq, L.vxref_svgf_temporal)
    assert all(_same(x, y) for x, y in zip(a, b)) and len(a) == 3
    if not be_useful:
        hist, prev_g = a[0], seq[0]["g"]
        p = sv.temporal_params(seq[1]["cam"], seq[0]["cam"], abi.ATT_GI_SH, abi.ATT_SVGF_TEMPORAL_B, abi.ATT_SVGF_TEMPORAL_A, be_useful=False)
        assert _same(ob.svgf_temporal(p, seq[1]["raw"], hist, seq[1]["g"], prev_g), ob.svgf_temporal(p, seq[1]["raw"], hist, seq[1]["g"], prev_g, L.vxref_svgf_temporal))


@pytest.fixture(scope="module")
def temporal(seq):
    return _temporal_frames(seq)


def test_variance_and_spatial_self_checks(seq, temporal):
    f, t = seq[2], temporal[2]
    v = ob.svgf_variance(sv.variance_params(f["cam"], abi.ATT_SVGF_TEMPORAL_A), t, f["g"])
    # do_spatial off: the inputs pass through and variance = moment - luminance^2
    v0 = ob.svgf_variance(sv.variance_params(f["cam"], abi.ATT_SVGF_TEMPORAL_A, do_spatial=False), t, f["g"])
    # (bilinear weights at a pixel centre are not exactly 0 / 1 in float, so "pass through" is to within a half ulp or so)
    assert np.allclose(v0["cocg"].astype(np.float32), t["cocg"].astype(np.float32), atol=2e-3, rtol=2e-3)
    lum = np.maximum(0, 3.544905 * t["sh"][..., 3].astype(np.float32))
    assert np.allclose(v0["x"].astype(np.float32), t["x"][..., 1].astype(np.float32) - lum * lum, atol=5e-2, rtol=1e-2)
    acc = t["x"][..., 0].astype(np.float32)
    # pixels without history: THRESH / 0 makes the variance inf (clamped to 50) or NaN (0 * inf)
    fresh = acc == 0
    vx = v["x"].astype(np.float32)
    assert fresh.any() and (np.isnan(vx[fresh]) | (vx[fresh] == 50) | (vx[fresh] == -1)).all()
    assert np.isfinite(vx[~fresh]).all() and (vx[~fresh] >= -1).all() and (vx[~fresh] <= 50).all()
    # the 9x9 bilateral pre-filter smooths the SH luminance band
    lum_in, lum_out = t["sh"][..., 3].astype(np.float32), v["sh"][..., 3].astype(np.float32)
    assert np.abs(np.diff(lum_out, axis=1)).mean() < 0.7 * np.abs(np.diff(lum_in, axis=1)).mean()
    outs = sv.spatial_chain(f["cam"], t, v, f["g"], ob.svgf_spatial)
    assert len(outs) == 5
    last = outs[-1]
    assert np.isfinite(last["sh"].astype(np.float32)).all()
    # AO is only filtered by the iterations with step <= 4 (the first two pass it through)
    assert np.array_equal(outs[0]["aosky"][..., 0], t["aosky"][..., 0]) and np.array_equal(outs[1]["aosky"][..., 0], t["aosky"][..., 0])
    assert not np.array_equal(outs[2]["aosky"][..., 0], outs[1]["aosky"][..., 0])
    # do_spatial off: identity on sh / cocg / variance / ao
    ident = ob.svgf_spatial(sv.spatial_params(f["cam"], abi.ATT_SVGF_VARIANCE, abi.ATT_SVGF_TEMPORAL_A, abi.ATT_SVGF_TEMPORAL_A, abi.ATT_SVGF_DENOISE_A, 16, do_spatial=False),
                            v, t["aosky"], t["x"], f["g"])
    ok = ~np.isnan(v["x"])
    assert np.allclose(ident["sh"].astype(np.float32), v["sh"].astype(np.float32), atol=2e-3, rtol=2e-3)
    assert np.allclose(ident["x"].astype(np.float32)[ok], v["x"].astype(np.float32)[ok], atol=2e-3, rtol=2e-3)
    assert np.abs(ident["aosky"].astype(np.int32) - t["aosky"].astype(np.int32)).max() <= 1


@pytest.mark.skipif(not rb.available("svgf_variance"), reason="oracle/_ref not built on this box")
@pytest.mark.parametrize("do_spatial,aggressive", [(True, True), (True, False), (False, True)])
def test_oracle_variance_equals_compiled_reference_shader(seq, temporal, do_spatial, aggressive):
    L = rb.lib()
    for k in (0, 2):
        p = sv.variance_params(seq[k]["cam"], abi.ATT_SVGF_TEMPORAL_A, do_spatial, aggressive)
        a = ob.svgf_variance(p, temporal[k], seq[k]["g"])
        b = ob.svgf_variance(p, temporal[k], seq[k]["g"], L.vxref_svgf_variance)
        assert sv.same_bits(a, b, ("sh", "cocg", "x"))


@pytest.mark.skipif(not rb.available("svgf_spatial"), reason="oracle/_ref not built on this box")
@pytest.mark.parametrize("kw", [dict(), dict(large=True, time=7.3), dict(aggressive=False, phi_bias=0.05, res_scale=1.0), dict(do_spatial=False)])
def test_oracle_spatial_chain_equals_compiled_reference_shader(seq, temporal, kw):
    L = rb.lib()
    f, t = seq[2], temporal[2]
    v = ob.svgf_variance(sv.variance_params(f["cam"], abi.ATT_SVGF_TEMPORAL_A), t, f["g"])
    a = sv.spatial_chain(f["cam"], t, v, f["g"], ob.svgf_spatial, **kw)
    b = sv.spatial_chain(f["cam"], t, v, f["g"], lambda *args: ob.svgf_spatial(*args, fn=L.vxref_svgf_spatial), **kw)
    assert all(sv.same_bits(x, y) for x, y in zip(a, b))


@pytest.mark.skipif(not rb.available("svgf_prespatial"), reason="oracle/_ref not built on this box")
def test_oracle_prespatial_equals_compiled_reference_shader(seq):
    L = rb.lib()
    for k, f in enumerate(seq):
        p = sv.prespatial_params(f["cam"], time=0.4 + k)
        a = ob.svgf_prespatial(p, f["raw"], f["g"])
        assert sv.same_bits(a, ob.svgf_prespatial(p, f["raw"], f["g"], L.vxref_svgf_prespatial)), k
        # G-buffer at another resolution than the GI images (InitialTraceFBO is full size, the GI trace is not)
        g2 = {"t": np.ascontiguousarray(f["g"]["t"][::2, ::2]), "normal": np.ascontiguousarray(f["g"]["normal"][::2, ::2])}
        assert sv.same_bits(ob.svgf_prespatial(p, f["raw"], g2), ob.svgf_prespatial(p, f["raw"], g2, L.vxref_svgf_prespatial)), k


def test_oracle_prespatial_behaviour_and_golden(seq):
    """the 3 x 3 pass smooths the raw trace inside surfaces, leaves isolated pixels alone, and equals the compiled shader's output
    stored in tests/golden/svgf_ref.npz"""
    f = seq[0]
    out = ob.svgf_prespatial(sv.prespatial_params(f["cam"], time=1.0), f["raw"], f["g"])
    lum_in, lum_out = f["raw"]["sh"][..., 3].astype(np.float32), out["sh"][..., 3].astype(np.float32)
    hit = f["g"]["t"].astype(np.float32) > 0
    assert np.abs(np.diff(lum_out, axis=1))[hit[:, 1:]].mean() < 0.9 * np.abs(np.diff(lum_in, axis=1))[hit[:, 1:]].mean()
    assert 0.1 < (out["sh"].view(np.uint16) != f["raw"]["sh"].view(np.uint16)).any(axis=-1).mean() < 0.9
    z = np.load(str(sv.__file__).rsplit("/", 1)[0] + "/golden/svgf_ref.npz")
    for k in ("sh", "cocg", "x"):
        assert np.array_equal(out[k].view(np.uint16), z[f"prespatial0_{k}"].view(np.uint16)), k
    assert np.array_equal(out["aosky"], z["prespatial0_aosky"])


def test_oracle_chain_equals_golden_fixture(seq):
    """tests/golden/svgf_ref.npz was produced by the reference's own shaders (tests/golden/make_golden_svgf.py)."""
    import sys
    sys.path.insert(0, str(sv.__file__).rsplit("/", 1)[0] + "/golden")
    import make_golden_svgf as mg

    z = np.load(mg.OUT / "svgf_ref.npz")
    assert str(z["input_sha256"]) == mg.input_hash(seq), "the oracle no longer regenerates the fixture's inputs"
    got = mg.run_chain(seq, ob.svgf_temporal, ob.svgf_variance, ob.svgf_spatial)
    for name, a in got.items():
        assert sv.same_bits({"v": a}, {"v": z[name]}, ("v",)), name
