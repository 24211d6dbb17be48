"""Sun-shadow denoiser: oracle self-checks and oracle == the reference's own ShadowTemporalFilter.glsl /
ShadowFilter.glsl compiled for the CPU (oracle/_ref) over a four-frame sequence.  CPU only."""
import numpy as np
import pytest

import shadow_filter_util as sf
from oracle import binding as ob
from oracle import ref_binding as rb
from voxeltracing_b200 import host_api


@pytest.fixture(scope="module")
def seq():
    return sf.frames(host_api.gen_world("plains", 1))


def test_denoiser_behaviour(seq):
    outs = sf.run_chain(seq, ob.shadow_temporal, ob.shadow_filter)
    raw = seq[-1]["raw"]["shadow"]
    assert 0.02 < (raw == 0).mean() < 0.98, "the scene needs lit and shadowed pixels"
    t, flt = outs[-1]
    # soft shadows: the raw trace is binary, the accumulated result has penumbra values
    assert set(np.unique(raw)) <= {0, 255}
    assert ((t["shadow"] > 8) & (t["shadow"] < 247)).mean() > 0.002
    # the frame counter accumulates where reprojection succeeds and is reset elsewhere / on the sky
    fr = [o[0]["frames"].astype(np.float32) for o in outs]
    assert (fr[0] <= 1.0).all() and fr[-1].max() >= 3.0 and (fr[-1] >= 0).all() and (fr[-1] <= 256).all()
    sky = seq[-1]["g"]["t"].astype(np.float32) < 0
    assert sky.any() and (fr[-1][sky] == 0).all()
    # sky and sharp-shadow pixels pass through the spatial filter
    assert np.array_equal(flt[sky], t["shadow"][sky])
    assert not np.array_equal(flt, t["shadow"])
    # u_ShadowTemporal off: plain accumulation of the raw trace
    plain = sf.run_chain(seq, ob.shadow_temporal, ob.shadow_filter, shadow_temporal=False)
    assert not np.array_equal(plain[-1][0]["shadow"], t["shadow"])


@pytest.mark.skipif(not (rb.available("shadow_temporal") and rb.available("shadow_filter")), reason="oracle/_ref not built on this box")
@pytest.mark.parametrize("shadow_temporal,scale", [(True, 1.0), (False, 1.0), (True, 2.5)])
def test_oracle_equals_compiled_reference_shaders(seq, shadow_temporal, scale):
    L = rb.lib()
    a = sf.run_chain(seq, ob.shadow_temporal, ob.shadow_filter, shadow_temporal, scale)
    b = sf.run_chain(seq, lambda *x: ob.shadow_temporal(*x, fn=L.vxref_shadow_temporal), lambda *x: ob.shadow_filter(*x, fn=L.vxref_shadow_filter),
                     shadow_temporal, scale)
    for k, ((ta, fa), (tb, fb)) in enumerate(zip(a, b)):
        assert np.array_equal(ta["shadow"], tb["shadow"]), k
        assert np.array_equal(ta["frames"].view(np.uint16), tb["frames"].view(np.uint16)), k
        assert np.array_equal(fa, fb), k


def test_oracle_chain_equals_golden_fixture(seq):
    """tests/golden/shadow_filter_ref.npz was produced by the reference's own shaders (make_golden_shadow_filter.py)."""
    import sys
    sys.path.insert(0, str(sf.__file__).rsplit("/", 1)[0] + "/golden")
    import make_golden_shadow_filter as mg

    z = np.load(mg.OUT / "shadow_filter_ref.npz")
    assert str(z["input_sha256"]) == mg.input_hash(seq), "the oracle no longer regenerates the fixture's inputs"
    for k, (t, flt) in enumerate(sf.run_chain(seq, ob.shadow_temporal, ob.shadow_filter)):
        assert np.array_equal(t["shadow"], z[f"temporal{k}_shadow"]) and np.array_equal(flt, z[f"filtered{k}"]), k
        assert np.array_equal(t["frames"].view(np.uint16), z[f"temporal{k}_frames"].view(np.uint16)), k
