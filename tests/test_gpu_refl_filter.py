"""CUDA reflection temporal filter vs the oracle and the golden fixture, through the C ABI.  The pass evaluates
one exp() per pixel (CUDA vs libm: <= 2 ulp) before rounding to R16F: every frame is fed the CUDA output of the frame before
it and held to 2 half-ulps with >= 99.5 % of the R16F values identical."""
from pathlib import Path

import numpy as np
import pytest

import refl_filter_util as rf
from oracle import binding as ob
from voxeltracing_b200 import abi, engine, host_api

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "refl_filter_ref.npz"


@pytest.fixture(scope="module")
def seq():
    return rf.frames(host_api.gen_world("plains", 0))


def _close_f16(got, want, what):
    g, w = got.astype(np.float32), want.astype(np.float32)
    assert (np.abs(g - w) <= 2.0 ** -10 * np.abs(w) + 1e-6).all(), (what, float(np.abs(g - w).max()))
    assert (got.view(np.uint16) == want.view(np.uint16)).mean() >= 0.995, (what, float((got.view(np.uint16) == want.view(np.uint16)).mean()))


def _load_frame(c, f):
    c.write_attachment(abi.ATT_INITIAL_T, f["g"]["t"]); c.write_attachment(abi.ATT_INITIAL_NORMAL, f["g"]["normal"])
    c.write_attachment(abi.ATT_INITIAL_BLOCK, np.zeros_like(f["g"]["normal"]))
    c.write_attachment(abi.ATT_REFL_COLOR, f["refl"]["color"]); c.write_attachment(abi.ATT_REFL_HITDIST, f["refl"]["hitdist"])
    c.write_attachment(abi.ATT_REFL_EMISSIVE, f["refl"]["mask"]); c.write_attachment(abi.ATT_GBUF_PBR, f["pbr"])


def _run(c, seq, **flags):
    """yields per frame (k, params, history fed, previous frame, CUDA temporal set)"""
    hist, prev = rf.zero_history(), None
    for k, f in enumerate(seq):
        hs, os_ = rf.sets_for(k)
        _load_frame(c, f)
        p = rf.params(f, prev or f, hs, os_, **flags)
        c.specular_temporal(p)
        t = {"color": c.read_attachment(os_), "frames": c.read_attachment(os_ + 1), "hitdist": c.read_attachment(os_ + 2)}
        yield k, p, hist, prev, t
        c.end_frame()
        hist, prev = t, f


@pytest.mark.parametrize("flags", [{}, {"temporal_spec": 0}, {"firefly_rejection": 0, "smart_clip": 0}, {"aggressive_firefly_rejection": 0, "roughness_weight": 0},
                                   {"stabilize_hit_distance": 0}])
def test_pass_matches_the_oracle(seq, flags):
    c = engine.Context(0)   # fresh context: frame 0 runs against the zero-filled history the library creates
    try:
        for k, p, hist, prev, t in _run(c, seq, **flags):
            prev_g = prev["g"] if prev else {"t": np.zeros((rf.H, rf.W), np.float16), "normal": np.zeros((rf.H, rf.W), np.uint8)}
            prev_hit = prev["refl"]["hitdist"] if prev else np.zeros((rf.RH, rf.RW), np.float16)
            want = ob.specular_temporal(p, seq[k]["refl"], prev_hit, hist, seq[k]["g"], prev_g, seq[k]["pbr"])
            for name in ("color", "frames", "hitdist"):
                _close_f16(t[name], want[name], (flags, k, name))
    finally:
        c.close()


def test_sequence_matches_golden(seq):
    z = np.load(GOLD)
    c = engine.Context(0)
    try:
        for k, p, hist, prev, t in _run(c, seq):
            for name in ("color", "frames", "hitdist"):
                _close_f16(t[name], z[f"{name}{k}"].view(np.float16), ("golden", k, name))
    finally:
        c.close()


def test_row_band_equals_full_frame_and_errors(seq):
    c = engine.Context(0)
    try:
        f = seq[1]
        with pytest.raises(engine.VxrtError):
            c.specular_temporal(rf.params(f, seq[0], *rf.sets_for(1)))   # no reflection trace yet
        _load_frame(c, f)
        hs, os_ = rf.sets_for(1)
        p = rf.params(f, seq[0], hs, os_)
        c.specular_temporal(p)
        full = c.read_attachment(os_)
        c.write_attachment(os_, np.zeros_like(full))
        for row0, rows in ((0, 40), (40, 68)):
            abi.set_tile(p.tile, (row0, rows))
            c.specular_temporal(p)
        assert np.array_equal(c.read_attachment(os_).view(np.uint16), full.view(np.uint16))
        bad = rf.params(f, seq[0], abi.ATT_REFL_TEMPORAL_A, abi.ATT_REFL_TEMPORAL_A)
        with pytest.raises(engine.VxrtError):
            c.specular_temporal(bad)
        # end_frame hands the reflection hit distance over
        c.end_frame()
        assert np.array_equal(c.read_attachment(abi.ATT_PREV_REFL_HITDIST).view(np.uint16), f["refl"]["hitdist"].view(np.uint16))
    finally:
        c.close()


# ---- spatial pass: ReflectionDenoiserNew.glsl (x then y) ----
def _close_denoised(got, want, what):
    """The pass chains up to seven powf() per tap over up to 33 taps (CUDA vs libm: <= 2 ulp each): 8 half-ulps, >= 95 % identical."""
    g, w = got.astype(np.float32), want.astype(np.float32)
    assert (np.abs(g - w) <= 2.0 ** -8 * np.abs(w) + 1e-5).all(), (what, float(np.abs(g - w).max()))
    same = float((got.view(np.uint16) == want.view(np.uint16)).mean())
    assert same >= 0.95, (what, same)


def _load_denoise_inputs(c, f, temporal, temporal_set):
    _load_frame(c, f)
    c.write_attachment(abi.ATT_GBUF_NORMAL, f["gb_normal"])
    c.write_attachment(temporal_set, temporal["color"]); c.write_attachment(temporal_set + 1, temporal["frames"])
    c.write_attachment(temporal_set + 2, temporal["hitdist"])


@pytest.mark.parametrize("flags", [{}, {"temporal_weight": 0, "normal_map_aware": 0}, {"roughness_bias": 0, "handle_lobe_deviation": 0, "amplify_transversal_weight": 0},
                                   {"derive_from_diffuse_sh": 1, "radius_bias": 1, "denoiser_scale": 2.5}, {"resolution_scale": 1.0, "normal_map_weight_strength": 0.3}])
def test_denoiser_matches_the_oracle(seq, flags):
    sets = rf.run_chain(seq, ob.specular_temporal)
    c = engine.Context(0)
    try:
        for k, stabilized in ((1, True), (3, False), (4, True)):
            ts = rf.sets_for(k)[1]
            _load_denoise_inputs(c, seq[k], sets[k], ts)

            def on_gpu(p, in_color, frames, hitdist, f):
                if p.dir == 0:   # the y pass is fed the oracle's x result, so each pass is compared on identical inputs
                    c.write_attachment(abi.ATT_REFL_DENOISED_A, in_color)
                c.reflection_denoise(p)
                return c.read_attachment(p.out_attachment)
            want = rf.run_denoise(seq[k], sets[k], ts, ob.reflection_denoise, stabilized, **flags)
            hit = sets[k]["hitdist"] if stabilized else seq[k]["refl"]["hitdist"]
            hit_att = ts + 2 if stabilized else abi.ATT_REFL_HITDIST
            px = rf.denoise_params(seq[k], 1, ts, abi.ATT_REFL_DENOISED_A, ts, hit_att, **flags)
            _close_denoised(on_gpu(px, sets[k]["color"], sets[k]["frames"], hit, seq[k]), want[0], (flags, k, "x"))
            py = rf.denoise_params(seq[k], 0, abi.ATT_REFL_DENOISED_A, abi.ATT_REFL_DENOISED_B, ts, hit_att, **flags)
            _close_denoised(on_gpu(py, want[0], sets[k]["frames"], hit, seq[k]), want[1], (flags, k, "y"))
    finally:
        c.close()


def test_denoiser_chain_matches_golden(seq):
    """temporal -> x -> y entirely on the GPU against the compiled shaders' outputs"""
    z = np.load(GOLD)
    c = engine.Context(0)
    try:
        for k, p, hist, prev, t in _run(c, seq):
            if k not in (1, 3, 4):
                continue
            ts = rf.sets_for(k)[1]
            c.write_attachment(abi.ATT_GBUF_NORMAL, seq[k]["gb_normal"])
            c.reflection_denoise(rf.denoise_params(seq[k], 1, ts, abi.ATT_REFL_DENOISED_A, ts, ts + 2))
            c.reflection_denoise(rf.denoise_params(seq[k], 0, abi.ATT_REFL_DENOISED_A, abi.ATT_REFL_DENOISED_B, ts, ts + 2))
            g, w = c.read_attachment(abi.ATT_REFL_DENOISED_B).astype(np.float32), z[f"denoise_y{k}"].view(np.float16).astype(np.float32)
            assert (np.abs(g - w) <= 2.0 ** -7 * np.abs(w) + 1e-4).mean() >= 0.999, (k, float(np.abs(g - w).max()))
    finally:
        c.close()


def test_denoiser_row_band_and_errors(seq):
    sets = rf.run_chain(seq, ob.specular_temporal)
    c = engine.Context(0)
    try:
        k, ts = 3, rf.sets_for(3)[1]
        p = rf.denoise_params(seq[k], 1, ts, abi.ATT_REFL_DENOISED_A, ts, ts + 2)
        with pytest.raises(engine.VxrtError):
            c.reflection_denoise(p)                     # nothing bound yet
        _load_denoise_inputs(c, seq[k], sets[k], ts)
        c.reflection_denoise(p)
        full = c.read_attachment(abi.ATT_REFL_DENOISED_A)
        c.write_attachment(abi.ATT_REFL_DENOISED_A, np.zeros_like(full))
        for row0, rows in ((0, 50), (50, 58)):
            abi.set_tile(p.tile, (row0, rows))
            c.reflection_denoise(p)
        assert np.array_equal(c.read_attachment(abi.ATT_REFL_DENOISED_A).view(np.uint16), full.view(np.uint16))
        for bad in (dict(out_att=abi.ATT_REFL_COLOR), dict(in_att=abi.ATT_REFL_DENOISED_A), dict(temporal_set=abi.ATT_GI_SH)):
            q = rf.denoise_params(seq[k], 1, bad.get("in_att", ts), bad.get("out_att", abi.ATT_REFL_DENOISED_A), bad.get("temporal_set", ts), ts + 2)
            with pytest.raises(engine.VxrtError):
                c.reflection_denoise(q)
    finally:
        c.close()
