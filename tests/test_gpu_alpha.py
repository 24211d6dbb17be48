"""Alpha-tested traversal through the C ABI vs the oracle: primary and shadow passes."""
import numpy as np
import pytest

import alpha_util as au
from oracle import binding as ob
from voxeltracing_b200 import abi, engine, host_api

pytestmark = pytest.mark.gpu
BLUE = np.random.default_rng(11).integers(0, 256, (256, 256, 4), dtype=np.uint8)


@pytest.fixture(scope="module")
def scene(plains0, plains0_oracle):
    inp = au.alpha_inputs()
    sc = ob.OracleScene(plains0_oracle)
    inp.apply_to_oracle(sc)
    c = engine.Context(0)
    c.upload_world(plains0)
    c.generate_distance_field()
    c.set_blue_noise_texture(BLUE)
    inp.apply_to_context(c)
    yield c, sc
    c.close()


@pytest.mark.parametrize("pose", range(len(au.POSES)))
@pytest.mark.parametrize("fov", [90.0, 70.0])
def test_alpha_tested_primary_and_shadow_bit_exact(scene, pose, fov):
    c, sc = scene
    pos, yaw, pitch = au.POSES[pose]
    cam = host_api.camera(pos, yaw, pitch, au.W / au.H, fov)
    c.stats_enable(True)
    c.stats_read(reset=True)
    p = c.initial_trace(cam, au.W, au.H, alpha_test=True, fov=fov)
    st = c.stats_read(reset=True)
    want = sc.initial_trace(p, want_stats=True)
    got = {"t": c.read_attachment(abi.ATT_INITIAL_T), "normal": c.read_attachment(abi.ATT_INITIAL_NORMAL),
           "block": c.read_attachment(abi.ATT_INITIAL_BLOCK), "inv_t": c.read_attachment(abi.ATT_INITIAL_INVT)}
    for k in got:
        assert np.array_equal(got[k].view(np.uint8), want[k].view(np.uint8)), k
    assert st == want["stats"]
    # the test is not vacuous: the alpha test changes what this pose sees
    opaque = sc.world.initial_trace(au.primary_params(cam, alpha=False, fov=fov))
    assert (opaque["block"] != want["block"]).sum() > 100
    for soft in (False, True):
        sp = c.shadow_trace(cam, au.W, au.H, host_api.sun_direction(50.0)[2], frame=3, soft=soft, alpha_test=True, fov=fov)
        ws = sc.shadow_trace(sp, got["t"], got["normal"], BLUE)
        gs, gt = c.read_attachment(abi.ATT_SHADOW), c.read_attachment(abi.ATT_SHADOW_TRANSVERSAL)
        if soft:  # sinf / cosf of the cone sample differ by ulps between CUDA and libm
            assert (gs == ws["shadow"]).mean() >= 0.999
        else:
            assert np.array_equal(gs, ws["shadow"])
            assert np.array_equal(gt.view(np.uint16), ws["transversal"].view(np.uint16))
    c.stats_enable(False)


def test_alpha_matches_reference_golden(scene):
    """CUDA vs the committed output of the reference's own shaders with u_ShouldAlphaTest on (tests/golden/alpha_ref.npz)."""
    c, _ = scene
    g = np.load(au.GOLDEN)
    for pi, (pos, yaw, pitch) in enumerate(au.POSES):
        cam = host_api.camera(pos, yaw, pitch, au.W / au.H, au.GOLDEN_FOV)
        c.initial_trace(cam, au.W, au.H, alpha_test=True, fov=au.GOLDEN_FOV)
        for k, att in (("t", abi.ATT_INITIAL_T), ("normal", abi.ATT_INITIAL_NORMAL), ("block", abi.ATT_INITIAL_BLOCK), ("inv_t", abi.ATT_INITIAL_INVT)):
            assert np.array_equal(c.read_attachment(att).view(np.uint8), g[f"pose{pi}_{k}"].view(np.uint8)), (pi, k)
        c.shadow_trace(cam, au.W, au.H, host_api.sun_direction(50.0)[2], frame=3, soft=False, alpha_test=True, fov=au.GOLDEN_FOV)
        assert np.array_equal(c.read_attachment(abi.ATT_SHADOW), g[f"pose{pi}_shadow"])
        assert np.array_equal(c.read_attachment(abi.ATT_SHADOW_TRANSVERSAL).view(np.uint16), g[f"pose{pi}_transversal"].view(np.uint16))


def test_alpha_needs_its_resources(plains0):
    c = engine.Context(0)
    c.upload_world(plains0)
    c.generate_distance_field()
    cam = host_api.camera([192, 75, 192], 0.0, -20.0, 16 / 9)
    with pytest.raises(engine.VxrtError):
        c.initial_trace(cam, 64, 36, alpha_test=True)      # no albedo array bound
    c.initial_trace(cam, 64, 36)                            # the default path does not need it
    c.close()
