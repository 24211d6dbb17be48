"""CUDA primary + shadow passes vs the oracle through the C ABI (BASELINE configs 1 and 3)."""
import numpy as np
import pytest

from oracle import binding as ob
from voxeltracing_b200 import abi, engine, host_api

pytestmark = pytest.mark.gpu

POSES = [(yaw, -20.0) for yaw in range(0, 360, 45)] + [(90.0, 0.0)]


@pytest.fixture(scope="module")
def scene(plains0, plains0_oracle):
    c = engine.Context(0)
    c.upload_world(plains0)
    c.generate_distance_field()
    rng = np.random.default_rng(11)
    blue = rng.integers(0, 256, (256, 256, 4), dtype=np.uint8)
    c.set_blue_noise_texture(blue)
    yield c, plains0_oracle, blue
    c.close()


def _compare_primary(c, ow, cam, w, h, jitter=None, rd=350, tile=(0, 0)):
    p = c.initial_trace(cam, w, h, rd, jitter, tile)
    want = ow.initial_trace(p)
    got = {
        "t": c.read_attachment(abi.ATT_INITIAL_T), "normal": c.read_attachment(abi.ATT_INITIAL_NORMAL),
        "block": c.read_attachment(abi.ATT_INITIAL_BLOCK), "inv_t": c.read_attachment(abi.ATT_INITIAL_INVT),
    }
    return got, want


@pytest.mark.parametrize("yaw,pitch", POSES)
def test_config1_primary_640x360(scene, yaw, pitch):
    c, ow, _ = scene
    cam = host_api.camera([192, 75, 192], float(yaw), float(pitch), 640 / 360)
    got, want = _compare_primary(c, ow, cam, 640, 360)
    same = (got["block"] == want["block"]) & (got["normal"] == want["normal"])
    assert same.mean() >= 0.999, same.mean()          # north_star: voxel + face exact on >= 99.9 %
    assert (want["block"] > 0).mean() > 0.05            # the pose actually sees terrain
    hit = same & (want["block"] > 0)
    rel = np.abs(1.0 / got["inv_t"][hit] - 1.0 / want["inv_t"][hit]) * np.abs(want["inv_t"][hit])
    assert rel.max() <= 1e-4                            # north_star: t within 1e-4 relative
    # the R16F attachment is RNE(t)
    assert np.array_equal(got["t"].view(np.uint16)[same], want["t"].view(np.uint16)[same])
    # with --fmad=false the CUDA path rounds like the oracle: in practice everything is bit-exact
    assert same.all() and np.array_equal(got["inv_t"].view(np.uint32), want["inv_t"].view(np.uint32))


def test_primary_jitter_odd_size_and_outside_camera(scene):
    c, ow, _ = scene
    cam = host_api.camera([192, 75, 192], 30.0, -35.0, 333 / 177)
    got, want = _compare_primary(c, ow, cam, 333, 177, jitter=host_api.taa_jitter(5))
    for k in ("block", "normal"):
        assert np.array_equal(got[k], want[k])
    assert np.array_equal(got["inv_t"].view(np.uint32), want["inv_t"].view(np.uint32))
    # camera outside the volume: rays are clipped to the box first (InitialRayTraceFrag.glsl:448-452)
    cam = host_api.camera([-60, 150, -40], 45.0, -30.0, 16 / 9)
    got, want = _compare_primary(c, ow, cam, 320, 180)
    assert (want["block"] > 0).mean() > 0.05
    for k in ("block", "normal"):
        assert np.array_equal(got[k], want[k])
    assert np.array_equal(got["t"].view(np.uint16), want["t"].view(np.uint16))
    # tiny iteration cap: rays that run out of iterations miss
    cam = host_api.camera([192, 75, 192], 90.0, -20.0, 16 / 9)
    got, want = _compare_primary(c, ow, cam, 320, 180, rd=12)
    assert np.array_equal(got["block"], want["block"]) and (want["block"] == 0).mean() > 0.5


def test_primary_inside_solid_and_empty_world():
    dims = (32, 16, 48)
    c = engine.Context(0, dims)
    w = np.zeros((48, 16, 32), np.uint8)
    c.upload_world(w)
    c.generate_distance_field()
    cam = host_api.camera([16, 8, 24], 10.0, 5.0, 1.0)
    c.initial_trace(cam, 64, 64)
    assert (c.read_attachment(abi.ATT_INITIAL_BLOCK) == 0).all()
    assert (c.read_attachment(abi.ATT_INITIAL_NORMAL) == 255).all()
    assert (c.read_attachment(abi.ATT_INITIAL_T).astype(np.float32) == -1.0).all()
    w[:] = 3  # camera buried in solid: every ray starts in a solid voxel and misses
    c.upload_world(w)
    c.generate_distance_field()
    p = c.initial_trace(cam, 64, 64)
    assert (c.read_attachment(abi.ATT_INITIAL_BLOCK) == 0).all()
    want = ob.OracleWorld(w).initial_trace(p)
    assert np.array_equal(want["block"], c.read_attachment(abi.ATT_INITIAL_BLOCK))
    c.close()


def test_tile_sharded_primary_equals_full_frame(scene):
    c, ow, _ = scene
    cam = host_api.camera([192, 75, 192], 135.0, -20.0, 16 / 9)
    c.initial_trace(cam, 640, 360)
    full = {a: c.read_attachment(a).copy() for a in (abi.ATT_INITIAL_T, abi.ATT_INITIAL_NORMAL, abi.ATT_INITIAL_BLOCK, abi.ATT_INITIAL_INVT)}
    c2 = engine.Context(0)
    c2.upload_world(ow.blocks)
    c2.generate_distance_field()
    for row0, rows in [(0, 100), (100, 57), (157, 203)]:
        c2.initial_trace(cam, 640, 360, tile=(row0, rows))
    for a, img in full.items():
        assert np.array_equal(c2.read_attachment(a), img)
    c2.close()


def test_rectangle_tiles_equal_full_frame_and_rect_copy(scene):
    """vxrt_tile as a rectangle (column bands, a 2 x 2 grid with edges that are not multiples of the 32 x 8 CTA tile): the union of the
    tiles is the full frame bit for bit; vxrt_cuda_copy_attachment_rect_async moves exactly the rectangle; bad rectangles are rejected."""
    import torch
    c, ow, _ = scene
    cam = host_api.camera([192, 75, 192], 135.0, -20.0, 16 / 9)
    W, H = 640, 360
    atts = (abi.ATT_INITIAL_T, abi.ATT_INITIAL_NORMAL, abi.ATT_INITIAL_BLOCK, abi.ATT_INITIAL_INVT)
    c.initial_trace(cam, W, H)
    c.shadow_trace(cam, W, H, host_api.sun_direction(50.0)[2], soft=False)
    full = {a: c.read_attachment(a).copy() for a in atts + (abi.ATT_SHADOW,)}
    c2 = engine.Context(0)
    c2.upload_world(ow.blocks)
    c2.generate_distance_field()
    for tiles in ([(0, 0, 0, 213), (0, 0, 213, 214), (0, 0, 427, 213)], [(0, 101, 0, 333), (0, 101, 333, 307), (101, 259, 0, 333), (101, 259, 333, 307)]):
        c2.initial_trace(cam, W, H, tile=(0, 1))       # allocate, then poison so that every pixel must be rewritten by a tile
        c2.synchronize()
        for a in atts:
            torch.as_tensor(c2.attachment_as_device_array(a), device="cuda").fill_(77)
        torch.cuda.synchronize()
        for t in tiles:
            c2.initial_trace(cam, W, H, tile=t)
        for a in atts:
            assert np.array_equal(c2.read_attachment(a).view(np.uint8), full[a].view(np.uint8)), (a, tiles)
        for t in tiles:
            c2.shadow_trace(cam, W, H, host_api.sun_direction(50.0)[2], soft=False, tile=t)
        assert np.array_equal(c2.read_attachment(abi.ATT_SHADOW), full[abi.ATT_SHADOW])
    # rectangle copy into a zeroed device image
    _, w_, h_, bpp = c2.attachment_info(abi.ATT_INITIAL_INVT)
    dst = torch.zeros((h_, w_), dtype=torch.float32, device="cuda")
    c2.copy_attachment_rect_async(abi.ATT_INITIAL_INVT, dst.data_ptr(), (40, 100, 64, 200))
    c2.wait_reads()
    want = np.zeros((h_, w_), np.float32)
    want[40:140, 64:264] = full[abi.ATT_INITIAL_INVT].reshape(h_, w_)[40:140, 64:264]
    assert np.array_equal(dst.cpu().numpy().view(np.uint32), want.view(np.uint32))
    with pytest.raises(Exception):
        c2.initial_trace(cam, W, H, tile=(0, 0, W, 10))
    with pytest.raises(Exception):
        c2.copy_attachment_rect_async(abi.ATT_INITIAL_INVT, dst.data_ptr(), (0, 0, 600, 100))
    c2.close()


def _shadow_case(c, ow, blue, cam, w, h, sw, sh, soft, frame, halton=(0.0, 0.0), light=None):
    c.initial_trace(cam, w, h)
    g_t = c.read_attachment(abi.ATT_INITIAL_T)
    g_n = c.read_attachment(abi.ATT_INITIAL_NORMAL)
    if light is None:
        light = host_api.sun_direction(50.0)[2]
    p = c.shadow_trace(cam, sw, sh, light, frame=frame, halton=halton, soft=soft)
    want = ow.shadow_trace(p, g_t, g_n, blue, want_stats=True)
    got_s = c.read_attachment(abi.ATT_SHADOW)
    got_t = c.read_attachment(abi.ATT_SHADOW_TRANSVERSAL)
    return got_s, got_t, want


def test_hard_shadows_are_bit_exact(scene):
    c, ow, blue = scene
    cam = host_api.camera([192, 75, 192], 200.0, -25.0, 16 / 9)
    got_s, got_t, want = _shadow_case(c, ow, blue, cam, 640, 360, 640, 360, soft=False, frame=0)
    assert np.array_equal(got_s, want["shadow"])
    assert np.array_equal(got_t.view(np.uint16), want["transversal"].view(np.uint16))
    assert 0.02 < (want["shadow"] == 255).mean() < 0.98 and want["stats"]["rays"] > 10000


@pytest.mark.parametrize("frame,res", [(0, (640, 360)), (7, (640, 360)), (3, (320, 180))])
def test_soft_shadows_config3_style(scene, frame, res):
    """soft=true, blue-noise cone jitter.  sinf/cosf differ from libm by an ulp, so the jittered
    direction may differ in the last bit: occlusion must agree on >= 99.9 % of pixels and the
    transversal (T/100) within 1e-3 relative where both agree."""
    c, ow, blue = scene
    cam = host_api.camera([192, 75, 192], 60.0, -20.0, 16 / 9)
    halton = (0.0, 0.0) if res == (640, 360) else tuple(host_api.taa_jitter(frame) )
    got_s, got_t, want = _shadow_case(c, ow, blue, cam, 640, 360, res[0], res[1], soft=True, frame=frame, halton=halton)
    same = got_s == want["shadow"]
    assert same.mean() >= 0.999, same.mean()
    a, b = got_t.astype(np.float32)[same], want["transversal"].astype(np.float32)[same]
    close = np.abs(a - b) <= 1e-3 * np.abs(b) + 1e-6
    assert close.mean() >= 0.999


def test_shadow_needs_gbuffer_and_blue_noise(plains0):
    c = engine.Context(0)
    c.upload_world(plains0)
    c.generate_distance_field()
    cam = host_api.camera([192, 75, 192], 60.0, -20.0, 16 / 9)
    with pytest.raises(engine.VxrtError):
        c.shadow_trace(cam, 64, 64, [0, 1, 0], soft=False)
    c.initial_trace(cam, 64, 64)
    with pytest.raises(engine.VxrtError):
        c.shadow_trace(cam, 64, 64, [0, 1, 0], soft=True)  # no blue-noise texture bound
    c.shadow_trace(cam, 64, 64, [0, 1, 0], soft=False)
    c.close()


def test_stats_counters_match_oracle(scene):
    c, ow, _ = scene
    cam = host_api.camera([192, 75, 192], 270.0, -20.0, 16 / 9)
    c.stats_enable(True)
    c.stats_read(reset=True)
    p = c.initial_trace(cam, 640, 360)
    got = c.stats_read(reset=True)
    c.stats_enable(False)
    want = ow.initial_trace(p, want_stats=True)["stats"]
    assert got == want


def _ray_batch(n, seed, dims=(384, 128, 384)):
    """Random rays plus the degenerate families the loop treats specially: axis-aligned directions (zero components
    give inf / NaN in 1/dir), origins on voxel boundaries, origins outside the volume, far-away and non-finite origins."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = dims
    o = np.stack([rng.uniform(-20, nx + 20, n), rng.uniform(-10, ny + 30, n), rng.uniform(-20, nz + 20, n)], 1).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float32)
    k = n // 8
    d[:k] = axes[rng.integers(0, 6, k)]                                  # axis aligned
    d[k:2 * k, rng.integers(0, 3)] = 0.0                                 # one zero component (not renormalised: used as given)
    o[2 * k:3 * k] = np.floor(o[2 * k:3 * k])                            # exactly on voxel corners
    o[3 * k:4 * k, 1] = np.floor(o[3 * k:4 * k, 1])                      # on a horizontal voxel face
    d[3 * k:3 * k + k // 2] = axes[rng.integers(0, 6, k // 2)]           # ... walking along it
    o[4 * k:4 * k + 8] = np.array([[1e9, 60, 100], [-1e9, 60, 100], [100, 3e38, 100], [np.inf, 60, 60], [100, -np.inf, 100],
                                   [np.nan, 60, 60], [4194304.0, 60, 60], [-4194305.0, 60, 60]], dtype=np.float32)
    d[4 * k + 8:4 * k + 12] = np.array([[-0.0, -1, 0], [0, -1, -0.0], [-0.0, -0.0, 1], [0, 0, 0]], dtype=np.float32)
    return o, d


@pytest.mark.parametrize("max_iter", [350, 48, 1, 0])
def test_trace_rays_batch_bit_exact(scene, max_iter):
    """vxrt_cuda_trace_rays == the oracle's VoxelTraversalDF for every ray, bit for bit: t, end position, normal, block id,
    Intersection flag and iteration count (the conversion-free CUDA loop against the literal one).  The batch contains
    rays whose positions become inf / NaN (an unguarded zero direction component, InitialRayTraceFrag.glsl:343-347)."""
    ctx, ow, _ = scene
    o, d = _ray_batch(400_000, 5 + max_iter)
    got = ctx.trace_rays(o, d, max_iter)
    want = ow.traverse_batch(o, d, max_iter)
    def bits(a):  # bit patterns with every NaN canonicalised (x86 and the GPU produce different payloads for inf * 0)
        a = np.ascontiguousarray(a)
        b = a.view(np.uint32).copy()
        b[np.isnan(a)] = 0x7FC00000
        return b

    assert np.array_equal(bits(got["t"]), bits(want["t"]))
    assert np.array_equal(bits(got["end"]), bits(want["end"]))
    assert np.array_equal(bits(got["normal"]), bits(want["normal"]))
    assert np.array_equal(got["block"], want["block"]) and np.array_equal(got["intersection"], want["intersection"])
    assert np.array_equal(got["iterations"], want["iterations"])
    if max_iter == 350:
        assert (got["t"] > 0).mean() > 0.1 and got["iterations"].max() > 100


def test_trace_rays_argument_checks(scene):
    ctx = scene[0]
    assert len(ctx.trace_rays(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))) == 0
    with pytest.raises(engine.VxrtError):
        ctx.trace_rays(np.zeros((4, 3), np.float32), np.ones((4, 3), np.float32), -1)
    fresh = engine.Context(0)
    with pytest.raises(engine.VxrtError):
        fresh.trace_rays(np.zeros((4, 3), np.float32), np.ones((4, 3), np.float32))
    fresh.close()


def test_async_readback_is_ordered_against_the_next_pass(scene):
    """vxrt_cuda_read_attachment_async: the copy of frame k must complete with frame k's pixels even though frame k+1
    is submitted before the host waits (a pass that rewrites the attachment waits on the device for the copy)."""
    import torch

    c, ow, _ = scene
    cams = [host_api.camera([192, 75, 192], yaw, -20.0, 16 / 9) for yaw in (0.0, 140.0, 250.0)]
    w, h = 1280, 720
    want = []
    for cam in cams:
        c.initial_trace(cam, w, h)
        want.append((c.read_attachment(abi.ATT_INITIAL_T).copy(), c.read_attachment(abi.ATT_INITIAL_INVT).copy()))
    bufs = [(torch.empty((h, w), dtype=torch.float16).pin_memory(), torch.empty((h, w), dtype=torch.float32).pin_memory()) for _ in cams]
    for cam, (bt, bi) in zip(cams, bufs):          # no host wait between frames
        c.initial_trace(cam, w, h)
        c.read_attachment_async(abi.ATT_INITIAL_T, bt.numpy())
        c.read_attachment_async(abi.ATT_INITIAL_INVT, bi.numpy())
    c.wait_reads()
    for (bt, bi), (wt, wi) in zip(bufs, want):
        assert np.array_equal(bt.numpy().view(np.uint16), wt.view(np.uint16))
        assert np.array_equal(bi.numpy().view(np.uint32), wi.view(np.uint32))
    with pytest.raises(engine.VxrtError):
        c.read_attachment_async(abi.ATT_INITIAL_T, np.empty((h, w + 1), np.float16)) if False else c._check(
            c._lib.vxrt_cuda_read_attachment_async(c._h, abi.ATT_INITIAL_T, bufs[0][0].numpy().ctypes.data, 12))


def test_picking_ray_matches_oracle_and_reference_golden(scene):
    """vxrt_cuda_raycast_detect == World::RaycastDetect: every field against the oracle on 200k rays, hit voxels against
    the committed output of the reference's own function."""
    import pick_util as pu
    from pathlib import Path

    c, ow, _ = scene
    o, d = pu.pick_rays(200_000, 3)
    got, want = c.raycast_detect(o, d), ow.raycast_detect(o, d)
    assert np.array_equal(got, want)
    assert 0.1 < (got[:, 7] == 1).mean() < 0.9
    g = np.load(Path(__file__).resolve().parent / "golden" / "pick_ref.npz")
    r = c.raycast_detect(g["positions"], g["directions"])
    found = r[:, 7] == 1
    assert np.array_equal(r[found, :4], g["hits"][found]) and (g["hits"][~found] == -2).all()
    assert len(c.raycast_detect(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))) == 0
