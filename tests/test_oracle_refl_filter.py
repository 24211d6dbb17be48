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
