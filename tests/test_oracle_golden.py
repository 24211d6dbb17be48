"""Pin the oracle: (1) against the committed golden vectors made from the reference's own shaders,
(2) live against oracle/_ref when that build is present.  CPU only."""
import numpy as np
import pytest

import golden_util as gu
from oracle import binding as ob
from oracle import ref_binding as rb
from voxeltracing_b200 import host_api


@pytest.fixture(scope="module")
def worlds():
    return gu.worlds()


def test_oracle_df_matches_reference_golden(worlds):
    z, hashes = gu.df_golden()
    for name, w in worlds.items():
        df = ob.distance_field(w)
        assert gu.sha(df) == hashes[name], name
    assert np.array_equal(ob.distance_field(worlds["plains0"])[192], z["plains0_z192"])
    assert np.array_equal(ob.distance_field(worlds["rooms2"])[190], z["rooms2_z190"])


def test_oracle_primary_and_shadow_match_reference_golden(worlds):
    g = gu.trace_golden()
    for wname in ("plains0", "rooms2"):
        ow = ob.OracleWorld(worlds[wname])
        for pi, (pos, yaw, pitch) in enumerate(gu.mg.POSES):
            cam = host_api.camera(pos, yaw, pitch, gu.mg.W / gu.mg.H)
            jitter = host_api.taa_jitter(5) if pi == 1 else None
            out = ow.initial_trace(gu.mg.primary_params(cam, jitter))
            key = f"{wname}_pose{pi}"
            for k in ("t", "normal", "block", "inv_t"):
                assert np.array_equal(out[k].view(np.uint8), g[f"{key}_{k}"].view(np.uint8)), (key, k)
            for soft, frame in ((0, 0), (1, 7)):
                s = ow.shadow_trace(gu.mg.shadow_params(cam, frame, soft), out["t"], out["normal"], gu.BLUE)
                assert np.array_equal(s["shadow"], g[f"{key}_shadow{soft}"]), (key, soft)
                assert np.array_equal(s["transversal"].view(np.uint16), g[f"{key}_transversal{soft}"].view(np.uint16)), (key, soft)
    # the fixtures are not degenerate
    assert (g["plains0_pose0_block"] > 0).mean() > 0.2 and 0.05 < (g["plains0_pose0_shadow1"] == 255).mean() < 0.95


@pytest.mark.skipif(not rb.available("df"), reason="oracle/_ref not built on this box")
def test_oracle_equals_compiled_reference_shaders_live(worlds):
    """Bigger than the fixtures: full 640x360 frames and the whole field, oracle vs oracle/_ref."""
    w = host_api.gen_world("town", 3)
    df = ob.distance_field(w)
    assert np.array_equal(rb.distance_field(w), df)
    ow = ob.OracleWorld(w, df)
    old_w, old_h = gu.mg.W, gu.mg.H
    gu.mg.W, gu.mg.H = 640, 360
    try:
        for pos, yaw, pitch in [([192, 90, 192], 10.0, -30.0), ([60, 70, 330], 250.0, -5.0)]:
            cam = host_api.camera(pos, yaw, pitch, 640 / 360)
            p = gu.mg.primary_params(cam)
            a, b = ow.initial_trace(p), rb.initial_trace(w, df, p)
            for k in ("t", "normal", "block", "inv_t", "t32"):
                assert np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)), k
            sp = gu.mg.shadow_params(cam, 3, 1)
            sa, sb = ow.shadow_trace(sp, a["t"], a["normal"], gu.BLUE), rb.shadow_trace(w, df, sp, a["t"], a["normal"], gu.BLUE)
            assert np.array_equal(sa["shadow"], sb["shadow"])
            assert np.array_equal(sa["transversal"].view(np.uint16), sb["transversal"].view(np.uint16))
    finally:
        gu.mg.W, gu.mg.H = old_w, old_h
