"""Light propagation volume flood fill, CPU side: the oracle restatement (oracle/vxrt_oracle_lpv.cpp) against the
reference's own Core/VolumetricFloodFill.cpp compiled in oracle/_ref (only where /root/reference is mounted) and against the golden
outputs of that build (tests/golden/lpv_ref.npz, everywhere).  Byte work: bit-exact."""
import os
from pathlib import Path

import numpy as np
import pytest

import lpv_util as lu
import world_util as wu
from oracle import world_binding as wb

REF = Path(os.environ.get("VXRT_REFERENCE", "/root/reference"))
needs_ref = pytest.mark.skipif(not (wb.ref_available() and REF.exists() and hasattr(wb.ref(), "vxref_lpv_edit")),
                               reason="oracle/_ref world library / reference tree not present")
TABLE = wu.emissive_table()


@pytest.fixture(scope="module")
def case0():
    blocks = lu.lamp_world(2, 400, "rooms")
    return blocks, wb.collect_lights(blocks, TABLE)


def test_repropagate_matches_golden(case0):
    g = lu.golden()
    blocks, lights = case0
    assert len(lights) == int(g["c0_n_lights"])
    for limit in (4, 8, 2, 3, 0, 11):
        level, color = wb.lpv_repropagate(blocks, lights, limit)
        assert np.array_equal(lu.crc(level, color), g[f"c0_l{limit}_crc"]), limit
    level, color = wb.lpv_repropagate(blocks, lights, 4)
    assert np.array_equal(level, lu.dense(g["c0_l4_level_idx"], g["c0_l4_level_val"], blocks.shape))
    assert np.array_equal(color, lu.dense(g["c0_l4_color_idx"], g["c0_l4_color_val"], blocks.shape))
    level, color = wb.lpv_repropagate(blocks, lights[::-1], 8)
    assert np.array_equal(lu.crc(level, color), g["c0_rev_crc"])
    # the order of the queue decides block types, not levels
    assert g["c0_rev_crc"][0] == g["c0_l8_crc"][0] and g["c0_rev_crc"][1] != g["c0_l8_crc"][1]


def test_repropagate_second_world_matches_golden():
    g = lu.golden()
    blocks = lu.lamp_world(1, 3000, "plains")
    lights = wb.collect_lights(blocks, TABLE)
    assert len(lights) == int(g["c1_n_lights"])
    for limit in (4, 8):
        assert np.array_equal(lu.crc(*wb.lpv_repropagate(blocks, lights, limit)), g[f"c1_l{limit}_crc"]), limit


def test_levels_are_distance_to_the_nearest_lamp():
    """Self-check on a small open grid: with one seed level L the level volume is max(0, L - L1 distance) over air reachable through
    lit air (no obstacles here), cut at 2 (a node of level 2 spreads no further; level-1 voxels never appear)."""
    dims = (40, 24, 32)
    blocks = np.zeros((dims[2], dims[1], dims[0]), dtype=np.uint8)
    lamps = np.array([[10, 12, 16], [13, 12, 16], [30, 5, 8], [1, 1, 1], [0, 3, 3]], dtype=np.int32)
    for x, y, z in lamps:
        blocks[z, y, x] = 12
    level, color = wb.lpv_repropagate(blocks, lamps, 8)
    zz, yy, xx = np.meshgrid(np.arange(dims[2]), np.arange(dims[1]), np.arange(dims[0]), indexing="ij")
    want = np.zeros_like(level, dtype=np.int64)
    for x, y, z in lamps[:4]:   # the fifth lies on the plane x = 0: outside the volume for the flood fill
        want = np.maximum(want, 8 - (np.abs(xx - x) + np.abs(yy - y) + np.abs(zz - z)))
    want[want < 2] = 0
    want[(xx == 0) | (yy == 0) | (zz == 0)] = 0
    # where two lamps' fields overlap the later one only overwrites voxels at least 3 darker: levels may stay up to 2 below the maximum
    assert np.all(level <= want) and np.all(want - level <= 2)
    alone = (np.abs(xx - 30) + np.abs(yy - 5) + np.abs(zz - 8)) <= 6
    assert np.array_equal(level[alone], want[alone].astype(np.uint8))
    assert not np.any(level == 1) and set(np.unique(color)) <= {0, 12}


def test_edit_sequence_matches_golden(case0):
    g = lu.golden()
    blocks, lights = case0
    for limit in (8, 4):
        b = blocks.copy()
        level, color = wb.lpv_repropagate(b, lights, limit)
        for k, e in enumerate(lu.edit_sequence(blocks, TABLE, 48, seed=5)):
            lu.apply_edit(b, e)
            wb.lpv_edit(b, e[0], e[1], e[2], e[3], limit, level, color)
            assert np.array_equal(lu.crc(level, color), g[f"edit_l{limit}_crc"][k]), (limit, k, e)
    assert np.array_equal(level, lu.dense(g["edit_l4_level_idx"], g["edit_l4_level_val"], blocks.shape))
    assert np.array_equal(color, lu.dense(g["edit_l4_color_idx"], g["edit_l4_color_val"], blocks.shape))


@needs_ref
def test_repropagate_matches_reference(case0):
    blocks, lights = case0
    rng = np.random.default_rng(3)
    for limit, order in ((4, lights), (8, lights), (7, lights[rng.permutation(len(lights))]), (5, lights[::-1])):
        lo, co = wb.lpv_repropagate(blocks, order, limit)
        lr, cr = wb.ref_lpv_repropagate(blocks, order, limit, iterations=4)
        assert np.array_equal(lo, lr) and np.array_equal(co, cr), limit


@needs_ref
def test_edit_sequence_matches_reference(case0):
    blocks, lights = case0
    b = blocks.copy()
    lo, co = wb.lpv_repropagate(b, lights, 6)
    lr, cr = lo.copy(), co.copy()
    for e in lu.edit_sequence(blocks, TABLE, 30, seed=11):
        lu.apply_edit(b, e)
        wb.lpv_edit(b, e[0], e[1], e[2], e[3], 6, lo, co)
        wb.ref_lpv_edit(b, e[0], e[1], e[2], e[3], 6, lr, cr)
        assert np.array_equal(lo, lr) and np.array_equal(co, cr), e


def test_average_block_colors_match_golden():
    """BlockAverageColorData (PrecomputeAverageBlockColor.comp): oracle == the compiled shader's output in the fixture, bit for bit"""
    import scene_util as su
    from oracle import binding as ob
    g = lu.golden()
    for size in (64, 512):
        inp = su.SceneInputs(size)
        sc = ob.OracleScene(ob.OracleWorld(np.zeros((16, 16, 16), np.uint8)))
        inp.apply_to_oracle(sc)
        a = sc.lpv_average_colors()
        assert np.array_equal(a.view(np.uint32), g[f"average_colors_{size}"].view(np.uint32)), size
        has = inp.table[0] >= 0
        assert (a[~has] == 0).all() and (a[has, :3] > 0).all() and (a[:, 3] == 0).all() and a.max() <= 1.0


def test_sample_lpv_data_matches_golden(case0):
    """SampleLPVData (ReflectionTraceFrag.glsl:1516-1528): oracle == the compiled shader function's output in the fixture, bit for bit"""
    g = lu.golden()
    blocks, lights = case0
    level, color = wb.lpv_repropagate(blocks, lights, 8)
    avg, pts, dithers = lu.sample_case(level)
    for k, d in enumerate(dithers):
        got = wb.lpv_sample(level, color, avg, pts, d)
        assert np.array_equal(got.view(np.uint32), g["sample_rgb"][k].view(np.uint32)), k
    assert (got.sum(axis=1) > 0).mean() > 0.9 and np.isfinite(got).all()
    # dark voxels give no light; the light scales with the level
    assert not wb.lpv_sample(np.zeros_like(level), color, avg, pts[:100], dithers[0]).any()


@needs_ref
def test_sample_lpv_data_matches_reference(case0):
    blocks, lights = case0
    level, color = wb.lpv_repropagate(blocks, lights[::-1], 5)
    avg, pts, dithers = lu.sample_case(level)
    for d in dithers:
        assert np.array_equal(wb.lpv_sample(level, color, avg, pts, d).view(np.uint32), wb.ref_lpv_sample(level, color, avg, pts, d).view(np.uint32))
