"""CUDA sun-shadow denoiser vs the oracle and the golden fixture, through the C ABI.  Both passes weight
with exp() / pow() (CUDA vs libm: <= 2 ulp) before rounding to R8 / R16F: every pass is fed the CUDA output of the pass
before it, and held to: R8 within one code with >= 99.5 % of the pixels identical, frame counter within 2 half ulps with
>= 99.5 % identical."""
import numpy as np
import pytest

import shadow_filter_util as sf
from oracle import binding as ob
from voxeltracing_b200 import abi, engine, host_api

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def seq():
    return sf.frames(host_api.gen_world("plains", 1))


def _close_u8(got, want, what):
    assert np.abs(got.astype(np.int32) - want.astype(np.int32)).max() <= 1, what
    assert (got == want).mean() >= 0.995, (what, (got == want).mean())


def _close_f16(got, want, what):
    g, w = got.astype(np.float32), want.astype(np.float32)
    assert (np.abs(g - w) <= 2.0 ** -10 * np.abs(w) + 1e-6).all(), (what, np.abs(g - w).max())
    assert (got.view(np.uint16) == want.view(np.uint16)).mean() >= 0.995, what


def _load_frame(c, f):
    c.write_attachment(abi.ATT_INITIAL_T, f["g"]["t"]); c.write_attachment(abi.ATT_INITIAL_NORMAL, f["g"]["normal"])
    c.write_attachment(abi.ATT_INITIAL_BLOCK, np.zeros_like(f["g"]["normal"]))
    c.write_attachment(abi.ATT_SHADOW, f["raw"]["shadow"]); c.write_attachment(abi.ATT_SHADOW_TRANSVERSAL, f["raw"]["transversal"])


def _run(c, seq, shadow_temporal=True, scale=1.0):
    """yields per frame (k, temporal params, history fed, prev t, CUDA temporal set, filter params, CUDA filtered)"""
    hist = {"shadow": np.zeros((sf.H, sf.W), np.uint8), "frames": np.zeros((sf.H, sf.W), np.float16)}
    prev_t, prev_cam = np.zeros((sf.H, sf.W), np.float16), seq[0]["cam"]
    for k, f in enumerate(seq):
        hs, os_ = sf.sets_for(k)
        _load_frame(c, f)
        tp = sf.temporal_params(f["cam"], prev_cam, hs, os_, shadow_temporal)
        c.shadow_temporal(tp)
        t = {"shadow": c.read_attachment(os_), "frames": c.read_attachment(os_ + 1)}
        fp = sf.filter_params(f["cam"], os_, scale)
        c.shadow_filter(fp)
        flt = c.read_attachment(abi.ATT_SHADOW_FILTERED)
        yield k, tp, hist, prev_t, t, fp, flt
        c.end_frame()
        hist, prev_t, prev_cam = t, f["g"]["t"], f["cam"]


@pytest.mark.parametrize("shadow_temporal,scale", [(True, 1.0), (False, 1.0), (True, 2.5)])
def test_passes_match_the_oracle(seq, shadow_temporal, scale):
    c = engine.Context(0)   # fresh context: frame 0 runs against the zero-filled history the library creates
    try:
        for k, tp, hist, prev_t, t, fp, flt in _run(c, seq, shadow_temporal, scale):
            want = ob.shadow_temporal(tp, seq[k]["raw"], hist, seq[k]["g"], prev_t)
            _close_u8(t["shadow"], want["shadow"], ("temporal", k))
            _close_f16(t["frames"], want["frames"], ("frames", k))
            _close_u8(flt, ob.shadow_filter(fp, t, seq[k]["raw"]["transversal"], seq[k]["g"]), ("filter", k))
        assert t["frames"].astype(np.float32).max() >= 3.0
    finally:
        c.close()


def test_cuda_chain_against_golden_fixture(seq):
    """CUDA vs the committed output of the reference's own shaders: each pass fed the fixture's previous stage."""
    import sys
    sys.path.insert(0, str(sf.__file__).rsplit("/", 1)[0] + "/golden")
    import make_golden_shadow_filter as mg

    z = np.load(mg.OUT / "shadow_filter_ref.npz")
    assert str(z["input_sha256"]) == mg.input_hash(seq)
    c = engine.Context(0)
    try:
        prev_cam = seq[0]["cam"]
        for k, f in enumerate(seq):
            hs, os_ = sf.sets_for(k)
            _load_frame(c, f)
            if k > 0:   # history and previous G-buffer from the fixture / the inputs
                c.write_attachment(hs, z[f"temporal{k - 1}_shadow"]); c.write_attachment(hs + 1, z[f"temporal{k - 1}_frames"])
                c.write_attachment(abi.ATT_PREV_INITIAL_T, seq[k - 1]["g"]["t"])
            c.shadow_temporal(sf.temporal_params(f["cam"], prev_cam, hs, os_))
            _close_u8(c.read_attachment(os_), z[f"temporal{k}_shadow"], ("temporal", k))
            _close_f16(c.read_attachment(os_ + 1), z[f"temporal{k}_frames"], ("frames", k))
            c.write_attachment(os_, z[f"temporal{k}_shadow"]); c.write_attachment(os_ + 1, z[f"temporal{k}_frames"])
            c.shadow_filter(sf.filter_params(f["cam"], os_))
            _close_u8(c.read_attachment(abi.ATT_SHADOW_FILTERED), z[f"filtered{k}"], ("filter", k))
            prev_cam = f["cam"]
    finally:
        c.close()


def test_chain_on_gpu_traced_shadows_and_tile_sharding():
    """primary -> soft shadow trace (0.75 x resolution) -> temporal -> filter for three frames entirely on the GPU; the last
    frame is compared with the oracle fed the CUDA attachments, and a row-band sharded run reproduces the full-frame one."""
    W, H, SW, SH = 256, 144, 192, 108
    blocks = host_api.gen_world("plains", 1)
    light = host_api.sun_direction(50.0)[2]

    def render(bands):
        c = engine.Context(0)
        c.upload_world(blocks); c.generate_distance_field(); c.set_blue_noise_texture(sf.BLUE)
        prev_cam, snap = None, {}
        for k, (pos, yaw, pitch) in enumerate(sf.POSES[:3]):
            cam = host_api.camera(pos, yaw, pitch, W / H)
            prev_cam = prev_cam or cam
            c.initial_trace(cam, W, H)
            c.shadow_trace(cam, SW, SH, light, frame=k, soft=True)
            hs, os_ = sf.sets_for(k)
            tps, fps = [], []
            for row0, rows in bands:
                tp = sf.temporal_params(cam, prev_cam, hs, os_); tp.width, tp.height = W, H
                abi.set_tile(tp.tile, (row0, rows))
                c.shadow_temporal(tp); tps.append(tp)
            for row0, rows in bands:
                fp = sf.filter_params(cam, os_); fp.width, fp.height = W, H
                abi.set_tile(fp.tile, (row0, rows))
                c.shadow_filter(fp); fps.append(fp)
            if k == 2:
                snap = {"tp": tps[0], "fp": fps[0], "t": {"shadow": c.read_attachment(os_), "frames": c.read_attachment(os_ + 1)},
                        "flt": c.read_attachment(abi.ATT_SHADOW_FILTERED), "hist": {"shadow": c.read_attachment(hs), "frames": c.read_attachment(hs + 1)},
                        "g": {"t": c.read_attachment(abi.ATT_INITIAL_T), "normal": c.read_attachment(abi.ATT_INITIAL_NORMAL)},
                        "prev_t": c.read_attachment(abi.ATT_PREV_INITIAL_T),
                        "raw": {"shadow": c.read_attachment(abi.ATT_SHADOW), "transversal": c.read_attachment(abi.ATT_SHADOW_TRANSVERSAL)}}
            c.end_frame()
            prev_cam = cam
        c.close()
        return snap

    full = render([(0, 0)])
    tp, fp = full["tp"], full["fp"]
    tp.tile.row0 = tp.tile.rows = fp.tile.row0 = fp.tile.rows = 0
    _close_u8(full["t"]["shadow"], ob.shadow_temporal(tp, full["raw"], full["hist"], full["g"], full["prev_t"])["shadow"], "temporal")
    _close_u8(full["flt"], ob.shadow_filter(fp, full["t"], full["raw"]["transversal"], full["g"]), "filter")
    assert ((full["flt"] > 8) & (full["flt"] < 247)).mean() > 0.002 and full["t"]["frames"].astype(np.float32).max() >= 2.0
    banded = render([(0, 48), (48, 56), (104, 40)])
    assert np.array_equal(full["t"]["shadow"], banded["t"]["shadow"]) and np.array_equal(full["flt"], banded["flt"])
    assert np.array_equal(full["t"]["frames"].view(np.uint16), banded["t"]["frames"].view(np.uint16))


def test_error_paths(seq):
    c = engine.Context(0)
    try:
        f = seq[0]
        tp = sf.temporal_params(f["cam"], f["cam"], abi.ATT_SHADOW_TEMPORAL_B, abi.ATT_SHADOW_TEMPORAL_A)
        with pytest.raises(engine.VxrtError):
            c.shadow_temporal(tp)                                  # no shadow trace / G-buffer yet
        _load_frame(c, f)
        with pytest.raises(engine.VxrtError):
            c.shadow_filter(sf.filter_params(f["cam"], abi.ATT_SHADOW_TEMPORAL_A))   # temporal set not written
        bad = sf.temporal_params(f["cam"], f["cam"], abi.ATT_SHADOW_TEMPORAL_A, abi.ATT_SHADOW_TEMPORAL_A)
        with pytest.raises(engine.VxrtError):
            c.shadow_temporal(bad)                                 # out_set == history_set
        with pytest.raises(engine.VxrtError):
            c.shadow_filter(sf.filter_params(f["cam"], abi.ATT_SHADOW))   # not a temporal set
        c.shadow_temporal(tp)
        c.shadow_filter(sf.filter_params(f["cam"], abi.ATT_SHADOW_TEMPORAL_A))
    finally:
        c.close()


def test_select_shadow_feeds_the_direct_term():
    """vxrt_cuda_select_shadow: the colour pass samples the denoised image instead of the raw trace (Pipeline.cpp:3838);
    the direct term then equals the oracle's evaluated on that image."""
    import scene_util as su
    from voxeltracing_b200 import pipeline

    W, H = 256, 144
    blocks = host_api.gen_world("plains", 1)
    inputs = su.SceneInputs(64)
    ow = ob.OracleWorld(blocks); sc = ob.OracleScene(ow); inputs.apply_to_oracle(sc)
    c = engine.Context(0)
    try:
        c.upload_world(blocks); c.generate_distance_field(); inputs.apply_to_context(c); c.set_blue_noise_texture(sf.BLUE)
        cam = host_api.camera([192.0, 62.0, 192.0], 30.0, -15.0, W / H)
        fr = pipeline.FrameRenderer(c, pipeline.FrameConfig(width=W, height=H, passes=("primary", "gbuffer", "shadow")), inputs.grass, inputs.cactus)
        den = pipeline.ShadowDenoiser(c, W, H)
        for k in range(3):
            fr.render(cam, k)
            den.run(cam, k)
            c.end_frame()
        dp = fr.params_for("direct", cam)
        c.shade_direct(dp)
        with_filtered = c.read_attachment(abi.ATT_DIRECT)
        c.select_shadow(abi.ATT_SHADOW)
        c.shade_direct(dp)
        with_raw = c.read_attachment(abi.ATT_DIRECT)
        assert not np.array_equal(with_filtered.view(np.uint16), with_raw.view(np.uint16))
        g = {"inv_t": c.read_attachment(abi.ATT_INITIAL_INVT)}
        gb = {"albedo": c.read_attachment(abi.ATT_GBUF_ALBEDO), "normal": c.read_attachment(abi.ATT_GBUF_NORMAL), "pbr": c.read_attachment(abi.ATT_GBUF_PBR),
              "texao": c.read_attachment(abi.ATT_GBUF_TEXAO)}
        want = sc.shade_direct(dp, g["inv_t"], gb, c.read_attachment(abi.ATT_SHADOW_FILTERED))
        a, b = with_filtered.astype(np.float32), want.astype(np.float32)
        assert (np.abs(a - b) <= 1e-2 * np.abs(b) + 1e-3).all(axis=-1).mean() >= 0.995
        with pytest.raises(engine.VxrtError):
            c.select_shadow(abi.ATT_GI_SH)
    finally:
        c.close()
