"""CUDA SVGF chain of the diffuse GI vs the oracle, through the C ABI.  The temporal pass has no
transcendental on its path: bit exact.  The variance and spatial passes weight their taps with exp() / pow(), where CUDA
and libm differ by <= 2 ulp: every stage is fed the CUDA output of the stage before it (so only that stage is under
test) and held to the tolerance written in `_close_sets`."""
import numpy as np
import pytest

import svgf_util as sv
from oracle import binding as ob
from voxeltracing_b200 import abi, engine, host_api, pipeline

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def seq():
    return sv.frames(host_api.gen_world("plains", 0))


@pytest.fixture(scope="module")
def ctx():
    c = engine.Context(0)
    yield c
    c.close()


def _load_frame(c, f):
    c.write_attachment(abi.ATT_INITIAL_T, f["g"]["t"]); c.write_attachment(abi.ATT_INITIAL_NORMAL, f["g"]["normal"])
    c.write_attachment(abi.ATT_INITIAL_BLOCK, f["g"]["block"])
    c.write_set(abi.ATT_GI_SH, f["raw"])


def _close_sets(got, want, keys, min_identical=0.97):
    """R16F outputs: NaN pattern identical; >= min_identical of the values bit-identical; every value within 4 half ulps
    (2^-9 relative) + 1e-4 absolute.  RG8 outputs: within one code."""
    for k in keys:
        g, w = got[k], want[k]
        if g.dtype == np.uint8:
            assert np.abs(g.astype(np.int32) - w.astype(np.int32)).max() <= 1, k
            assert (g == w).mean() >= min_identical, (k, (g == w).mean())
            continue
        ng, nw = np.isnan(g), np.isnan(w)
        assert np.array_equal(ng, nw), k
        gf, wf = g.astype(np.float32)[~ng], w.astype(np.float32)[~nw]
        same = g.view(np.uint16)[~ng] == w.view(np.uint16)[~nw]
        assert same.mean() >= min_identical, (k, same.mean())
        assert (np.abs(gf - wf) <= 2.0 ** -9 * np.abs(wf) + 1e-4).all(), (k, np.abs(gf - wf).max())


def _run_temporal(c, seq, be_useful=True):
    """three frames through the CUDA temporal pass; yields (frame index, params, history set, prev g, out set)"""
    prev_cam = seq[0]["cam"]
    hist = ob.svgf_alloc(sv.H, sv.W, 3)
    prev_g = sv.zero_gbuf()
    for k, f in enumerate(seq):
        cur_id, hist_id = (abi.ATT_SVGF_TEMPORAL_A, abi.ATT_SVGF_TEMPORAL_B) if k % 2 == 0 else (abi.ATT_SVGF_TEMPORAL_B, abi.ATT_SVGF_TEMPORAL_A)
        _load_frame(c, f)
        p = sv.temporal_params(f["cam"], prev_cam, abi.ATT_GI_SH, hist_id, cur_id, be_useful)
        c.svgf_temporal(p)
        out = c.read_set(cur_id)
        yield k, p, hist, prev_g, out
        c.svgf_end_frame()
        hist, prev_g, prev_cam = out, f["g"], f["cam"]


@pytest.mark.parametrize("be_useful", [True, False])
def test_temporal_bit_exact(seq, be_useful):
    c = engine.Context(0)   # fresh context: the first frame runs against the zero-filled history the library creates
    try:
        for k, p, hist, prev_g, out in _run_temporal(c, seq, be_useful):
            want = ob.svgf_temporal(p, seq[k]["raw"], hist, seq[k]["g"], prev_g)
            assert sv.same_bits(out, want), k
        # the accumulated-frame counter advanced on re-projected pixels
        if be_useful:
            assert (out["x"][..., 0].astype(np.float32) >= 1.9).mean() > 0.2
        # end_frame handed the G-buffer over
        assert np.array_equal(c.read_attachment(abi.ATT_PREV_INITIAL_T).view(np.uint16), seq[-1]["g"]["t"].view(np.uint16))
        assert np.array_equal(c.read_attachment(abi.ATT_PREV_INITIAL_BLOCK), seq[-1]["g"]["block"])
    finally:
        c.close()


def test_prespatial_pass(ctx, seq):
    """Spatial3x3Initial.glsl (one expf per tap): against the oracle with the svgf tolerance, against the compiled shader's output in
    the golden fixture, at a G-buffer resolution different from the GI images, and feeding the temporal filter."""
    z = np.load(str(sv.__file__).rsplit("/", 1)[0] + "/golden/svgf_ref.npz")
    for k, f in enumerate(seq):
        _load_frame(ctx, f)
        p = sv.prespatial_params(f["cam"], time=1.0)
        ctx.svgf_prespatial(p)
        got = ctx.read_set(abi.ATT_SVGF_PRESPATIAL)
        _close_sets(got, ob.svgf_prespatial(p, f["raw"], f["g"]), ("sh", "cocg", "x", "aosky"))
        assert np.array_equal(got["x"].view(np.uint16), ob.svgf_prespatial(p, f["raw"], f["g"])["x"].view(np.uint16))   # utility: no transcendental
        if k == 0:
            _close_sets(got, {n: z[f"prespatial0_{n}"] for n in ("sh", "cocg", "x", "aosky")}, ("sh", "cocg", "x", "aosky"))
    # half-resolution G-buffer
    g2 = {"t": np.ascontiguousarray(f["g"]["t"][::2, ::2]), "normal": np.ascontiguousarray(f["g"]["normal"][::2, ::2]),
          "block": np.ascontiguousarray(f["g"]["block"][::2, ::2])}
    c = engine.Context(0)
    try:
        _load_frame(c, {"g": g2, "raw": f["raw"]})
        c.svgf_prespatial(p)
        _close_sets(c.read_set(abi.ATT_SVGF_PRESPATIAL), ob.svgf_prespatial(p, f["raw"], g2), ("sh", "cocg", "x", "aosky"))
        # the temporal filter takes the pre-filtered set as its input (Pipeline.cpp:2488-2520); full-size G-buffer again
        _load_frame(c, f)
        c.svgf_prespatial(p)
        pre = c.read_set(abi.ATT_SVGF_PRESPATIAL)
        tp = sv.temporal_params(f["cam"], f["cam"], abi.ATT_SVGF_PRESPATIAL, abi.ATT_SVGF_TEMPORAL_B, abi.ATT_SVGF_TEMPORAL_A)
        c.svgf_temporal(tp)
        want = ob.svgf_temporal(tp, pre, ob.svgf_alloc(sv.H, sv.W, 3), f["g"], sv.zero_gbuf())
        assert sv.same_bits(c.read_set(abi.ATT_SVGF_TEMPORAL_A), want)
        with pytest.raises(engine.VxrtError):
            bad = sv.prespatial_params(f["cam"], in_set=abi.ATT_SVGF_TEMPORAL_A)
            c.svgf_prespatial(bad)
    finally:
        c.close()
    with pytest.raises(engine.VxrtError):
        c2 = engine.Context(0)
        try:
            c2.svgf_prespatial(p)   # nothing traced yet
        finally:
            c2.close()


@pytest.fixture(scope="module")
def after_temporal(ctx, seq):
    """context holding frame 2's G-buffer and temporal set (TEMPORAL_A)"""
    for k, p, hist, prev_g, out in _run_temporal(ctx, seq):
        pass
    return out


@pytest.mark.parametrize("do_spatial,aggressive", [(True, True), (True, False), (False, True)])
def test_variance_estimate(ctx, seq, after_temporal, do_spatial, aggressive):
    f = seq[2]
    p = sv.variance_params(f["cam"], abi.ATT_SVGF_TEMPORAL_A, do_spatial, aggressive)
    ctx.svgf_variance(p)
    got = ctx.read_set(abi.ATT_SVGF_VARIANCE, with_ao=False)
    want = ob.svgf_variance(p, after_temporal, f["g"])
    _close_sets(got, want, ("sh", "cocg", "x"))


@pytest.mark.parametrize("kw", [dict(), dict(large=True, time=7.3), dict(aggressive=False, phi_bias=0.05, res_scale=1.0), dict(do_spatial=False)])
def test_spatial_chain(ctx, seq, after_temporal, kw):
    f = seq[2]
    ctx.svgf_variance(sv.variance_params(f["cam"], abi.ATT_SVGF_TEMPORAL_A))
    prev = ctx.read_set(abi.ATT_SVGF_VARIANCE, with_ao=False)
    ao = after_temporal["aosky"]
    for i, step in enumerate(sv.STEPS):
        cur_id = abi.ATT_SVGF_DENOISE_A if i % 2 == 0 else abi.ATT_SVGF_DENOISE_B
        prev_id = abi.ATT_SVGF_VARIANCE if i == 0 else (abi.ATT_SVGF_DENOISE_B if i % 2 == 0 else abi.ATT_SVGF_DENOISE_A)
        p = sv.spatial_params(f["cam"], prev_id, abi.ATT_SVGF_TEMPORAL_A if i == 0 else prev_id, abi.ATT_SVGF_TEMPORAL_A, cur_id, step, **kw)
        ctx.svgf_spatial(p)
        got = ctx.read_set(cur_id)
        want = ob.svgf_spatial(p, prev, ao, after_temporal["x"], f["g"])   # fed the CUDA output of the previous iteration
        _close_sets(got, want, ("sh", "cocg", "x", "aosky"))
        prev, ao = got, got["aosky"]
    if kw.get("do_spatial", True):
        lum_in, lum_out = after_temporal["sh"][..., 3].astype(np.float32), got["sh"][..., 3].astype(np.float32)
        assert np.abs(np.diff(lum_out, axis=1)).mean() < 0.5 * np.abs(np.diff(lum_in, axis=1)).mean()   # it denoises


def test_chain_on_gpu_rendered_frames_and_tile_sharding():
    """primary -> GI -> SVGF for three frames entirely on the GPU (pipeline.SvgfChain); the last frame's stages are
    compared with the oracle fed the CUDA attachments, and a row-band sharded run reproduces the full-frame one."""
    import scene_util as su

    W, H = 256, 144
    blocks = host_api.gen_world("rooms", 2)
    inputs = su.SceneInputs(64)

    def render(bands):
        c = engine.Context(0)
        c.upload_world(blocks); c.generate_distance_field(); inputs.apply_to_context(c)
        chain = pipeline.SvgfChain(c, W, H)
        snaps = {}
        for k, (pos, yaw, pitch) in enumerate([([200.0, 58.0, 200.0], 30.0, -15.0), ([200.3, 58.0, 199.8], 32.0, -15.0), ([200.6, 58.1, 199.6], 34.0, -14.0)]):
            cam = host_api.camera(pos, yaw, pitch, W / H)
            for row0, rows in bands:
                c.initial_trace(cam, W, H, tile=(row0, rows))
            for row0, rows in bands:
                gp = su.gi_params(cam, W, H, frame=k, spp=1)
                abi.set_tile(gp.tile, (row0, rows))
                c.diffuse_trace(gp)
            prev_cam = chain.prev_cam or cam
            stages = []
            for b in bands:
                chain.prev_cam = prev_cam
                stages.append(chain.prepare(cam, k, tile=b))
            for s in range(len(stages[0])):          # every band runs stage s before any band runs stage s + 1
                for prepared in stages:
                    chain.submit([prepared[s]], end_frame=False)
                if k == 2:
                    snaps[stages[0][s][0]] = (stages[0][s][2], c.read_set(stages[0][s][2].out_set if hasattr(stages[0][s][2], "out_set") else abi.ATT_SVGF_VARIANCE,
                                                                        with_ao=stages[0][s][0] != "variance"))
            if k == 2:
                snaps["g"] = {"t": c.read_attachment(abi.ATT_INITIAL_T), "normal": c.read_attachment(abi.ATT_INITIAL_NORMAL), "block": c.read_attachment(abi.ATT_INITIAL_BLOCK)}
                snaps["prev_g"] = {"t": c.read_attachment(abi.ATT_PREV_INITIAL_T), "normal": c.read_attachment(abi.ATT_PREV_INITIAL_NORMAL),
                                   "block": c.read_attachment(abi.ATT_PREV_INITIAL_BLOCK)}
                snaps["raw"] = c.read_set(abi.ATT_GI_SH)
                snaps["hist"] = c.read_set(abi.ATT_SVGF_TEMPORAL_B)
                snaps["prev_cam"] = prev_cam
            c.svgf_end_frame()
        c.close()
        return snaps

    full = render([(0, 0)])
    g = full["g"]
    tp, t_out = full["temporal"]
    assert sv.same_bits(t_out, ob.svgf_temporal(tp, full["raw"], full["hist"], g, full["prev_g"]))
    assert (t_out["x"][..., 0].astype(np.float32) >= 1.9).mean() > 0.3
    vp, v_out = full["variance"]
    _close_sets(v_out, ob.svgf_variance(vp, t_out, g), ("sh", "cocg", "x"))
    prev, ao = v_out, t_out["aosky"]
    for i in range(5):
        sp, s_out = full[f"spatial{i}"]
        _close_sets(s_out, ob.svgf_spatial(sp, prev, ao, t_out["x"], g), ("sh", "cocg", "x", "aosky"))
        prev, ao = s_out, s_out["aosky"]
    banded = render([(0, 40), (40, 64), (104, 40)])
    for name in ["temporal", "variance"] + [f"spatial{i}" for i in range(5)]:
        assert sv.same_bits(full[name][1], banded[name][1], tuple(full[name][1].keys())), name


def test_error_paths(ctx, seq):
    c = engine.Context(0)
    try:
        p = sv.temporal_params(seq[0]["cam"], seq[0]["cam"], abi.ATT_GI_SH, abi.ATT_SVGF_TEMPORAL_B, abi.ATT_SVGF_TEMPORAL_A)
        with pytest.raises(engine.VxrtError):
            c.svgf_temporal(p)                      # no GI output / G-buffer yet
        with pytest.raises(engine.VxrtError):
            c.svgf_end_frame()
        _load_frame(c, seq[0])
        bad = sv.temporal_params(seq[0]["cam"], seq[0]["cam"], abi.ATT_GI_SH, abi.ATT_SVGF_TEMPORAL_A, abi.ATT_SVGF_TEMPORAL_A)
        with pytest.raises(engine.VxrtError):
            c.svgf_temporal(bad)                    # out_set aliases history_set
        with pytest.raises(engine.VxrtError):
            c.svgf_variance(sv.variance_params(seq[0]["cam"], abi.ATT_SVGF_TEMPORAL_A))   # temporal set not written
        c.svgf_temporal(p)
        with pytest.raises(engine.VxrtError):
            c.svgf_spatial(sv.spatial_params(seq[0]["cam"], abi.ATT_SVGF_VARIANCE, abi.ATT_SVGF_TEMPORAL_A, abi.ATT_SVGF_TEMPORAL_A, abi.ATT_SVGF_DENOISE_A, 16))  # no variance yet
        c.svgf_variance(sv.variance_params(seq[0]["cam"], abi.ATT_SVGF_TEMPORAL_A))
        with pytest.raises(engine.VxrtError):
            c.svgf_spatial(sv.spatial_params(seq[0]["cam"], abi.ATT_SVGF_VARIANCE, abi.ATT_SVGF_TEMPORAL_A, abi.ATT_SVGF_TEMPORAL_A, abi.ATT_SVGF_VARIANCE, 16))   # bad out_set
    finally:
        c.close()


def test_cuda_chain_against_golden_fixture(seq):
    """CUDA vs the committed output of the reference's own SVGF shaders (tests/golden/svgf_ref.npz): the temporal pass bit
    for bit over the three-frame sequence, the variance and spatial passes each fed the fixture's previous stage."""
    import sys
    sys.path.insert(0, str(sv.__file__).rsplit("/", 1)[0] + "/golden")
    import make_golden_svgf as mg

    z = np.load(mg.OUT / "svgf_ref.npz")
    assert str(z["input_sha256"]) == mg.input_hash(seq)
    gold = lambda stage, keys=("sh", "cocg", "x", "aosky"): {k: z[f"{stage}_{k}"] for k in keys}
    c = engine.Context(0)
    try:
        for k, p, hist, prev_g, out in _run_temporal(c, seq):
            assert sv.same_bits(out, gold(f"temporal{k}")), k
        f = seq[2]
        c.svgf_variance(sv.variance_params(f["cam"], abi.ATT_SVGF_TEMPORAL_A))
        _close_sets(c.read_set(abi.ATT_SVGF_VARIANCE, with_ao=False), gold("variance", ("sh", "cocg", "x")), ("sh", "cocg", "x"))
        c.write_set(abi.ATT_SVGF_VARIANCE, gold("variance", ("sh", "cocg", "x")))
        for i, step in enumerate(sv.STEPS):
            cur_id = abi.ATT_SVGF_DENOISE_A if i % 2 == 0 else abi.ATT_SVGF_DENOISE_B
            prev_id = abi.ATT_SVGF_VARIANCE if i == 0 else (abi.ATT_SVGF_DENOISE_B if i % 2 == 0 else abi.ATT_SVGF_DENOISE_A)
            c.svgf_spatial(sv.spatial_params(f["cam"], prev_id, abi.ATT_SVGF_TEMPORAL_A if i == 0 else prev_id, abi.ATT_SVGF_TEMPORAL_A, cur_id, step, time=mg.TIME))
            _close_sets(c.read_set(cur_id), gold(f"spatial{i}"), ("sh", "cocg", "x", "aosky"))
            c.write_set(cur_id, gold(f"spatial{i}"))
    finally:
        c.close()


def test_filter_snap_tolerance_mode_stays_within_one_percent():
    """set_option("filter_snap", 256): bilinear weights below 1/256 snap to 0 and a tap with both weights 0 is one texel load
    (filter_sampler.cuh).  Tolerance of the mode, stated here: every value of the SVGF chain's output within 1e-2 relative (+1e-2
    absolute) of the bit-faithful mode on >= 99.9 % of the pixels; snap 0 restores the bit-faithful outputs exactly."""
    import scene_util as su
    W2, H2 = 320, 180
    inputs = su.SceneInputs(64, sky="constant")
    c = engine.Context(0)
    try:
        c.upload_world(host_api.gen_world("rooms", 2)); c.generate_distance_field(); inputs.apply_to_context(c)
        fr = pipeline.FrameRenderer(c, pipeline.FrameConfig(width=W2, height=H2, passes=("primary", "gi")), inputs.grass, inputs.cactus)

        def run(snap):
            c.set_option("filter_snap", snap)
            chain = pipeline.SvgfChain(c, W2, H2, pre_spatial=True)
            for k in range(4):
                cam = pipeline.rooms_camera(k // 2, W2 / H2)
                fr.render(cam, k)
                chain.run(cam, k)
            return [c.read_attachment(a).astype(np.float32) for a in (abi.ATT_SVGF_DENOISE_A, abi.ATT_SVGF_DENOISE_A + 1, abi.ATT_SVGF_DENOISE_B)]

        exact, snapped, again = run(0), run(256), run(0)
        for a, b in zip(exact, again):
            assert np.array_equal(a, b, equal_nan=True)
        for a, b in zip(exact, snapped):
            ok = np.isfinite(a) & np.isfinite(b)
            close = np.abs(a - b)[ok] <= 1e-2 * np.abs(a)[ok] + 1e-2
            assert close.mean() >= 0.999, close.mean()
        with pytest.raises(engine.VxrtError):
            c.set_option("filter_snap", 1 << 20)
    finally:
        c.close()
