"""The alpha-tested traversal: oracle self-checks and oracle == the reference's own shaders
compiled for the CPU (oracle/_ref) with u_ShouldAlphaTest on.  CPU only."""
import numpy as np
import pytest

import alpha_util as au
from oracle import binding as ob
from oracle import ref_binding as rb
from voxeltracing_b200 import host_api

BLUE = np.random.default_rng(11).integers(0, 256, (256, 256, 4), dtype=np.uint8)


@pytest.fixture(scope="module")
def scene(plains0, plains0_oracle):
    inp = au.alpha_inputs()
    sc = ob.OracleScene(plains0_oracle)
    inp.apply_to_oracle(sc)
    return inp, sc


def test_alpha_test_looks_through_cut_out_leaves(scene, plains0_oracle):
    inp, sc = scene
    seen_through = 0
    for pos, yaw, pitch in au.POSES:
        cam = host_api.camera(pos, yaw, pitch, au.W / au.H)
        opaque = plains0_oracle.initial_trace(au.primary_params(cam, alpha=False))
        alpha = sc.initial_trace(au.primary_params(cam, alpha=True))
        leaves = opaque["block"] == 7
        assert leaves.mean() > 0.02, "pose does not look at leaves"
        # pixels that are not leaves without the test are unchanged by it
        same = ~leaves
        assert np.array_equal(opaque["block"][same], alpha["block"][same])
        assert np.array_equal(opaque["t"].view(np.uint16)[same], alpha["t"].view(np.uint16)[same])
        seen_through += int((alpha["block"][leaves] != 7).sum())
        # a ray stopped by a leaf texel is never closer than the opaque hit
        both = leaves & (alpha["t32"] > 0)
        assert (alpha["t32"][both] >= opaque["t32"][both] - 1e-3).all()
    assert seen_through > 500


def test_alpha_params_off_is_the_plain_traversal(scene, plains0_oracle):
    _, sc = scene
    cam = host_api.camera(*au.POSES[0], au.W / au.H)
    a = plains0_oracle.initial_trace(au.primary_params(cam, alpha=False))
    b = sc.initial_trace(au.primary_params(cam, alpha=False))
    for k in ("t", "normal", "block", "inv_t"):
        assert np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8))


@pytest.mark.skipif(not (rb.available("initial") and rb.available("shadow")), reason="oracle/_ref not built on this box")
@pytest.mark.parametrize("fov", [90.0, 70.0])
def test_oracle_alpha_equals_compiled_reference_shaders(scene, plains0, plains0_oracle, fov):
    inp, sc = scene
    rb.set_scene(plains0, plains0_oracle.df, inp.table, inp.blue, inp.textures, inp.sky)
    for pos, yaw, pitch in au.POSES:
        cam = host_api.camera(pos, yaw, pitch, au.W / au.H, fov)
        p = au.primary_params(cam, alpha=True, fov=fov)
        a, b = sc.initial_trace(p), rb.initial_trace(plains0, plains0_oracle.df, p)
        for k in ("t", "normal", "block", "inv_t", "t32"):
            assert np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)), (k, pos)
        for soft in (False, True):
            sp = au.shadow_params(cam, alpha=True, fov=fov, soft=soft, frame=3)
            sa = sc.shadow_trace(sp, a["t"], a["normal"], BLUE)
            sb = rb.shadow_trace(plains0, plains0_oracle.df, sp, a["t"], a["normal"], BLUE)
            assert np.array_equal(sa["shadow"], sb["shadow"]), (pos, soft)
            assert np.array_equal(sa["transversal"].view(np.uint16), sb["transversal"].view(np.uint16)), (pos, soft)


def test_oracle_alpha_matches_reference_golden(scene):
    """tests/golden/alpha_ref.npz was produced by the reference's own shaders (tests/golden/make_golden_alpha.py)."""
    _, sc = scene
    g = np.load(au.GOLDEN)
    for pi, (pos, yaw, pitch) in enumerate(au.POSES):
        cam = host_api.camera(pos, yaw, pitch, au.W / au.H, au.GOLDEN_FOV)
        a = sc.initial_trace(au.primary_params(cam, alpha=True, fov=au.GOLDEN_FOV))
        for k in ("t", "normal", "block", "inv_t"):
            assert np.array_equal(a[k].view(np.uint8), g[f"pose{pi}_{k}"].view(np.uint8)), (pi, k)
        s = sc.shadow_trace(au.shadow_params(cam, alpha=True, fov=au.GOLDEN_FOV, soft=False, frame=3), a["t"], a["normal"], BLUE)
        assert np.array_equal(s["shadow"], g[f"pose{pi}_shadow"]) and np.array_equal(s["transversal"].view(np.uint16), g[f"pose{pi}_transversal"].view(np.uint16))
    assert (g["pose0_block"] == 7).any() and (g["pose0_block"] != 7).any()
