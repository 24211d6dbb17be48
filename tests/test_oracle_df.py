"""Oracle self-checks for the distance field; CPU only."""
import numpy as np
import pytest

from conftest import random_small_world
from oracle import binding as ob


@pytest.mark.parametrize("seed,density", [(0, 0.002), (1, 0.02), (2, 0.3), (3, 0.9)])
def test_df_equals_brute_force_and_literal(seed, density):
    w = random_small_world(seed, density)
    fast = ob.distance_field(w)
    assert np.array_equal(fast, ob.distance_field(w, "brute"))
    # float-carrying restatement of the shaders' imageLoad/imageStore round trips
    assert np.array_equal(fast, ob.distance_field(w, "literal"))


def test_df_adversarial_grids():
    nx, ny, nz = 32, 16, 48
    empty = np.zeros((nz, ny, nx), np.uint8)
    assert (ob.distance_field(empty) == min(254, nx + ny + nz)).all()  # X.comp:51
    full = np.full((nz, ny, nx), 7, np.uint8)
    assert (ob.distance_field(full) == 0).all()
    for corner in [(0, 0, 0), (nz - 1, ny - 1, nx - 1), (0, ny - 1, 0), (nz - 1, 0, nx - 1)]:
        w = empty.copy()
        w[corner] = 1
        z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        want = np.minimum(254, abs(z - corner[0]) + abs(y - corner[1]) + abs(x - corner[2]))
        assert np.array_equal(ob.distance_field(w), want.astype(np.uint8))
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    checker = (((x + y + z) & 1) * 5).astype(np.uint8)
    assert np.array_equal(ob.distance_field(checker), (checker == 0).astype(np.uint8))


def test_df_clamps_at_254_on_engine_sized_grid():
    w = np.zeros((384, 128, 384), np.uint8)
    w[0, 0, 0] = 1
    df = ob.distance_field(w)
    assert df[0, 0, 0] == 0 and df[0, 0, 200] == 200 and df[383, 127, 383] == 254 and df.max() == 254
    assert df[100, 100, 53] == 253 and df[100, 100, 54] == 254


def test_unorm8_round_trip_is_identity_and_step_table():
    # val/255.0f -> unorm8 -> floor(r*255.0f) must be the identity
    k = np.arange(256, dtype=np.float32)
    r = (k / np.float32(255.0)).astype(np.float32)
    stored = np.rint(r * np.float32(255.0)).astype(np.int64)
    assert np.array_equal(stored, np.arange(256))
    loaded = np.floor(r * np.float32(255.0))
    assert np.array_equal(loaded, k)
    # ... so the traversal step is a pure function of the byte; CUDA uses this closed form
    table = ob.step_table()
    closed = np.where(k == 1, 1, np.floor(k * np.float32(0.57735026918))).astype(np.int32)
    assert np.array_equal(table, closed)
    assert table[0] == 0 and list(table[1:4]) == [1, 1, 1] and table[4] == 2 and table[254] == 146
