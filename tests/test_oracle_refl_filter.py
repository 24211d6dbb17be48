"""Reflection temporal filter, CPU side: the oracle restatement (oracle/vxrt_oracle_refl_filter.cpp) against
SpecularTemporalFilter.glsl itself compiled through the GLSL shim (oracle/_ref; only where it was built) — bit-identical R16F
outputs over a 5-frame sequence for every flag combination the engine can set — and against the golden outputs of that build."""
import numpy as np
import pytest

import refl_filter_util as rf
from oracle import binding as ob
from oracle import ref_binding as rb

from pathlib import Path

GOLD = Path(__file__).parent / "golden" / "refl_filter_ref.npz"
FLAG_SETS = [{}, {"temporal_spec": 0}, {"firefly_rejection": 0, "smart_clip": 0}, {"aggressive_firefly_rejection": 0, "roughness_weight": 0},
             {"stabilize_hit_distance": 0}]


@pytest.fixture(scope="module")
def seq(plains0):
    return rf.frames(plains0)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint16)


def test_sequence_exercises_the_shader(seq):
    outs = rf.run_chain(seq, ob.specular_temporal)
    fr = [np.asarray(o["frames"], np.float32) for o in outs]
    assert (fr[0] == 0).all()                                  # zero history: no previous normal matches... except misses
    assert (fr[1] > 0).mean() > 0.3 and (fr[3] > 0).mean() > 0.3   # history is accepted on later frames
    assert 0 < (fr[1] == 0).mean() < 0.7                       # ... and rejected somewhere (disocclusion / screen edge / sky)
    for o in outs:
        assert np.isfinite(np.asarray(o["color"], np.float32)).all()
    # the hit-distance stabilisation changes the third image where history was accepted
    plain = rf.run_chain(seq, ob.specular_temporal, stabilize_hit_distance=0)
    assert (bits(plain[3]["hitdist"]) != bits(outs[3]["hitdist"])).mean() > 0.05
    # clipping is skipped on the frame whose camera did not move
    noclip = rf.run_chain(seq, ob.specular_temporal, smart_clip=0)
    assert (bits(noclip[1]["color"]) != bits(outs[1]["color"])).any()
    off = rf.run_chain(seq, ob.specular_temporal, temporal_spec=0)
    assert (np.asarray(off[2]["frames"], np.float32) == -1).all()


@pytest.mark.skipif(not rb.available("specular_temporal"), reason="oracle/_ref not built on this box")
@pytest.mark.parametrize("flags", FLAG_SETS)
def test_oracle_equals_compiled_reference_shader(seq, flags):
    L = rb.lib()
    L.vxref_specular_temporal.restype = None
    a = rf.run_chain(seq, ob.specular_temporal, **flags)
    b = rf.run_chain(seq, lambda *x: ob.specular_temporal(*x, fn=L.vxref_specular_temporal), **flags)
    for k, (x, y) in enumerate(zip(a, b)):
        for name in ("color", "frames", "hitdist"):
            assert np.array_equal(bits(x[name]), bits(y[name])), (flags, k, name, int((bits(x[name]) != bits(y[name])).sum()))


def test_oracle_matches_golden(seq):
    z = np.load(GOLD)
    for k, o in enumerate(rf.run_chain(seq, ob.specular_temporal)):
        for name in ("color", "frames", "hitdist"):
            assert np.array_equal(bits(o[name]), z[f"{name}{k}"]), (k, name)


# ---- spatial pass: ReflectionDenoiserNew.glsl ----
DENOISE_FLAG_SETS = [{}, {"temporal_weight": 0, "normal_map_aware": 0}, {"roughness_bias": 0, "handle_lobe_deviation": 0, "amplify_transversal_weight": 0},
                     {"derive_from_diffuse_sh": 1, "radius_bias": 1, "denoiser_scale": 2.5}, {"resolution_scale": 1.0, "normal_map_weight_strength": 0.3}]


@pytest.fixture(scope="module")
def temporal_sets(seq):
    return rf.run_chain(seq, ob.specular_temporal)


def test_denoiser_filters(seq, temporal_sets):
    k = 3
    x, y = rf.run_denoise(seq[k], temporal_sets[k], rf.sets_for(k)[1], ob.reflection_denoise)
    src = np.asarray(temporal_sets[k]["color"], np.float32)
    xf, yf = np.asarray(x, np.float32), np.asarray(y, np.float32)
    assert np.isfinite(yf).all() and (bits(x) != bits(temporal_sets[k]["color"])).mean() > 0.5 and (bits(y) != bits(x)).mean() > 0.5
    # a weighted average never leaves the range of its inputs
    assert yf.min() >= src.min() - 1e-3 and yf.max() <= src.max() + 1e-3
    # smoothing: the filtered image has less high-frequency energy than its input
    hf = lambda a: float(np.abs(np.diff(a[..., 0], axis=1)).mean())
    assert hf(xf) < hf(src) and hf(yf) <= hf(xf) * 1.05


@pytest.mark.skipif(not rb.available("reflection_denoise"), reason="oracle/_ref not built on this box")
@pytest.mark.parametrize("flags", DENOISE_FLAG_SETS)
def test_denoiser_oracle_equals_compiled_reference_shader(seq, temporal_sets, flags):
    L = rb.lib()
    L.vxref_reflection_denoise.restype = None
    for k, stabilized in ((1, True), (3, False), (4, True)):
        a = rf.run_denoise(seq[k], temporal_sets[k], rf.sets_for(k)[1], ob.reflection_denoise, stabilized, **flags)
        b = rf.run_denoise(seq[k], temporal_sets[k], rf.sets_for(k)[1], lambda *x: ob.reflection_denoise(*x, fn=L.vxref_reflection_denoise, time=3.7 * k),
                           stabilized, **flags)
        for name, x, y in (("x", a[0], b[0]), ("y", a[1], b[1])):
            assert np.array_equal(bits(x), bits(y)), (flags, k, name, int((bits(x) != bits(y)).sum()))


def test_denoiser_matches_golden(seq, temporal_sets):
    z = np.load(GOLD)
    for k in (1, 3, 4):
        x, y = rf.run_denoise(seq[k], temporal_sets[k], rf.sets_for(k)[1], ob.reflection_denoise)
        assert np.array_equal(bits(x), z[f"denoise_x{k}"]) and np.array_equal(bits(y), z[f"denoise_y{k}"]), k
