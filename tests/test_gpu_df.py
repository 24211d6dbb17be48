"""CUDA distance field vs the oracle, through the C ABI (bit-exact; BASELINE config 2)."""
import numpy as np
import pytest

from conftest import random_small_world
from oracle import binding as ob
from voxeltracing_b200 import engine, host_api

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = engine.Context(0)
    yield c
    c.close()


def test_df_small_grids_match_oracle_and_brute_force():
    for dims in [(32, 16, 48), (16, 16, 16), (64, 32, 16), (48, 20, 40)]:
        c = engine.Context(0, dims)
        for seed, density in [(0, 0.002), (1, 0.05), (2, 0.5)]:
            w = random_small_world(seed, density, dims)
            c.upload_world(w)
            c.generate_distance_field()
            got = c.download_distance_field()
            assert np.array_equal(got, ob.distance_field(w)), (dims, seed)
            assert np.array_equal(got, ob.distance_field(w, "brute")), (dims, seed)
        c.close()


def test_df_adversarial_full_size(ctx):
    nz, ny, nx = 384, 128, 384
    empty = np.zeros((nz, ny, nx), np.uint8)
    ctx.upload_world(empty)
    ctx.generate_distance_field()
    assert (ctx.download_distance_field() == 254).all()
    full = np.full((nz, ny, nx), 255, np.uint8)
    ctx.upload_world(full)
    ctx.generate_distance_field()
    assert (ctx.download_distance_field() == 0).all()
    for corner in [(0, 0, 0), (nz - 1, ny - 1, nx - 1), (nz - 1, 0, 0), (0, ny - 1, nx - 1)]:
        w = empty.copy()
        w[corner] = 128
        ctx.upload_world(w)
        ctx.generate_distance_field()
        assert np.array_equal(ctx.download_distance_field(), ob.distance_field(w))
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    checker = (((x + y + z) & 1) * 9).astype(np.uint8)
    ctx.upload_world(checker)
    ctx.generate_distance_field()
    assert np.array_equal(ctx.download_distance_field(), (checker == 0).astype(np.uint8))


def test_config2_regeneration_after_1024_edits(ctx, plains0):
    """plains(seed=0) + 1024 random toggles (mt19937(1234)) -> regenerate -> memcmp vs oracle."""
    ctx.upload_world(plains0)
    ctx.generate_distance_field()
    assert np.array_equal(ctx.download_distance_field(), ob.distance_field(plains0))
    edited = plains0.copy()
    edits = host_api.random_edits(edited, 1024, 1234)
    ctx.edit_blocks(edits)
    assert np.array_equal(ctx.download_world(), edited)
    ctx.generate_distance_field()
    got = ctx.download_distance_field()
    want = ob.distance_field(edited)
    assert np.array_equal(got, want)
    # regeneration is idempotent and does not disturb the block grid
    ctx.generate_distance_field()
    assert np.array_equal(ctx.download_distance_field(), want)
    assert np.array_equal(ctx.download_world(), edited)


def test_edit_semantics(ctx, plains0):
    ctx.upload_world(plains0)
    # a later edit of the same voxel wins, like sequential glTexSubImage3D calls
    ctx.edit_blocks(np.array([[5, 100, 7, 9], [5, 100, 7, 0], [5, 100, 7, 33], [6, 100, 7, 1]], np.int32))
    w = ctx.download_world()
    assert w[7, 100, 5] == 33 and w[7, 100, 6] == 1
    with pytest.raises(engine.VxrtError):
        ctx.edit_blocks(np.array([[384, 0, 0, 1]], np.int32))
    with pytest.raises(engine.VxrtError):
        ctx.edit_blocks(np.array([[0, 0, 0, 256]], np.int32))
    ctx.edit_blocks(np.zeros((0, 4), np.int32))  # empty list is a no-op


def test_state_errors():
    c = engine.Context(0)
    with pytest.raises(engine.VxrtError):
        c.generate_distance_field()
    with pytest.raises(engine.VxrtError):
        c.download_distance_field()
    with pytest.raises(ValueError):
        c.upload_world(np.zeros((4, 4, 4), np.uint8))
    c.close()


def test_worlds_other_than_plains(ctx):
    for kind, seed in [("rooms", 2), ("town", 3)]:
        w = host_api.gen_world(kind, seed)
        ctx.upload_world(w)
        ctx.generate_distance_field()
        assert np.array_equal(ctx.download_distance_field(), ob.distance_field(w)), kind


@pytest.mark.parametrize("xyver,zver", [(1, 1), (2, 1), (1, 2), (2, 2), (2, 4), (2, 5), (2, 9), (3, 2), (3, 1)])
def test_every_kernel_version_is_bit_exact(xyver, zver, plains0):
    """The kernel generations stay selectable (set_option df_xyver / df_zver) as each other's cross-check: every combination equals the
    oracle on a terrain world, an enclosed world, sparse voxels with block ids >= 128 (the mask multiply must ignore the high bit
    it builds on), dense noise and single voxels on the faces of the grid."""
    c = engine.Context(0)
    c.set_option("df_xyver", xyver); c.set_option("df_zver", zver)
    nz, ny, nx = 384, 128, 384
    rng = np.random.default_rng(17)
    sparse = np.zeros((nz, ny, nx), np.uint8)
    idx = rng.integers(0, sparse.size, 600)
    sparse.reshape(-1)[idx] = rng.integers(1, 256, 600).astype(np.uint8)
    dense = (rng.random((nz, ny, nx)) < 0.3).astype(np.uint8) * rng.integers(1, 256, (nz, ny, nx)).astype(np.uint8)
    faces = np.zeros((nz, ny, nx), np.uint8)
    for p in [(0, 64, 200), (383, 5, 3), (100, 0, 0), (200, 127, 383), (191, 63, 0), (192, 64, 383)]:
        faces[p] = 255
    for name, w in (("plains0", plains0), ("rooms2", host_api.gen_world("rooms", 2)), ("sparse", sparse), ("dense", dense), ("faces", faces)):
        c.upload_world(w)
        c.generate_distance_field()
        assert np.array_equal(c.download_distance_field(), ob.distance_field(w)), (xyver, zver, name)
    c.close()
