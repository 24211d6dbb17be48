"""Light propagation volume flood fill on the GPU through the C ABI: vxrt_cuda_lpv_repropagate (ordered level-synchronous
flood fill) and vxrt_cuda_lpv_edit (the exact queue of the block edit) against the oracle and against the golden outputs of the
reference's own Core/VolumetricFloodFill.cpp (tests/golden/lpv_ref.npz).  Byte work: bit-exact, both volumes."""
import numpy as np
import pytest

import lpv_util as lu
import world_util as wu
from oracle import world_binding as wb
from voxeltracing_b200 import engine

pytestmark = pytest.mark.gpu
TABLE = wu.emissive_table()


@pytest.fixture(scope="module")
def ctx():
    c = engine.Context(0)
    c.set_block_data(TABLE)
    yield c
    c.close()


@pytest.fixture(scope="module")
def case0():
    blocks = lu.lamp_world(2, 400, "rooms")
    return blocks, wb.collect_lights(blocks, TABLE)


@pytest.fixture(params=[1, 0], ids=["cooperative", "kernel-per-phase"])
def coop(ctx, request):
    """both implementations of the repropagation: one persistent cooperative kernel (default) and one kernel per phase"""
    ctx.set_option("lpv_coop", request.param)
    yield request.param
    ctx.set_option("lpv_coop", 1)


def test_repropagate_matches_golden_and_oracle(ctx, case0, coop):
    g = lu.golden()
    blocks, lights = case0
    ctx.upload_world(blocks)
    for limit in (4, 8, 2, 3, 0, 11):
        ctx.lpv_repropagate(None, limit)       # light list scanned on the device
        level, color = ctx.lpv_download()
        assert np.array_equal(lu.crc(level, color), g[f"c0_l{limit}_crc"]), limit
        lo, co = wb.lpv_repropagate(blocks, lights, limit)
        assert np.array_equal(level, lo) and np.array_equal(color, co), limit
    ctx.lpv_repropagate(lights, 4)             # light list from the host
    level, color = ctx.lpv_download()
    assert np.array_equal(level, lu.dense(g["c0_l4_level_idx"], g["c0_l4_level_val"], blocks.shape))
    assert np.array_equal(color, lu.dense(g["c0_l4_color_idx"], g["c0_l4_color_val"], blocks.shape))


def test_repropagate_follows_the_queue_order(ctx, case0, coop):
    g = lu.golden()
    blocks, lights = case0
    ctx.upload_world(blocks)
    ctx.lpv_repropagate(lights[::-1], 8)
    assert np.array_equal(lu.crc(*ctx.lpv_download()), g["c0_rev_crc"])
    rng = np.random.default_rng(3)
    order = lights[rng.permutation(len(lights))]
    order = np.concatenate([order, order[:50]])   # lights queued twice
    ctx.lpv_repropagate(order, 7)
    level, color = ctx.lpv_download()
    lo, co = wb.lpv_repropagate(blocks, order, 7)
    assert np.array_equal(level, lo) and np.array_equal(color, co)


def test_repropagate_dense_lights(ctx, coop):
    g = lu.golden()
    blocks = lu.lamp_world(1, 3000, "plains")
    ctx.upload_world(blocks)
    for limit in (4, 8):
        ctx.lpv_repropagate(None, limit)
        assert np.array_equal(lu.crc(*ctx.lpv_download()), g[f"c1_l{limit}_crc"]), limit
    # a wall of lamps: frontiers of several hundred thousand voxels, many CTAs in the ordered compaction
    blocks = np.zeros_like(blocks)
    blocks[100:140, 30, :] = 12
    blocks[100:140:3, 30, ::5] = 41
    blocks[:, 90, 200] = 41
    ctx.upload_world(blocks)
    ctx.lpv_repropagate(None, 8)
    level, color = ctx.lpv_download()
    lo, co = wb.lpv_repropagate(blocks, wb.collect_lights(blocks, TABLE), 8)
    assert np.array_equal(level, lo) and np.array_equal(color, co)


def test_repropagate_edge_cases(ctx, coop):
    blocks = np.zeros((384, 128, 384), dtype=np.uint8).reshape(384, 128, 384)
    ctx.upload_world(blocks)
    ctx.lpv_repropagate(None, 8)               # no lights
    level, color = ctx.lpv_download()
    assert not level.any() and not color.any()
    ctx.lpv_repropagate(np.zeros((0, 3), dtype=np.int32), 8)
    assert not ctx.lpv_download()[0].any()
    # lamps on and next to the faces, a lamp sealed in stone, lights given at air voxels and outside the grid
    blocks[:, :50, :] = 3
    for x, y, z in ((0, 60, 60), (1, 60, 90), (383, 60, 60), (60, 127, 60), (60, 0, 60), (60, 60, 0), (60, 60, 383), (200, 20, 200)):
        blocks[z, y, x] = 12
    ctx.upload_world(blocks)
    lights = np.concatenate([wb.collect_lights(blocks, TABLE), np.array([[100, 100, 100], [-1, 5, 5], [400, 5, 5]], dtype=np.int32)])
    ctx.lpv_repropagate(lights, 8)
    level, color = ctx.lpv_download()
    lo, co = wb.lpv_repropagate(blocks, lights, 8)
    assert np.array_equal(level, lo) and np.array_equal(color, co)
    assert level[200, 20, 200] == 8 and level[200, 21, 200] == 0
    with pytest.raises(engine.VxrtError):
        ctx.lpv_repropagate(None, -1)


def test_repropagate_other_dims():
    dims = (64, 48, 32)
    c = engine.Context(0, dims)
    c.set_block_data(TABLE)
    rng = np.random.default_rng(9)
    blocks = (rng.random((dims[2], dims[1], dims[0])) < 0.3).astype(np.uint8) * 3
    for _ in range(40):
        blocks[rng.integers(0, dims[2]), rng.integers(0, dims[1]), rng.integers(0, dims[0])] = 12
    c.upload_world(blocks)
    c.lpv_repropagate(None, 8)
    level, color = c.lpv_download()
    lo, co = wb.lpv_repropagate(blocks, wb.collect_lights(blocks, TABLE), 8)
    assert np.array_equal(level, lo) and np.array_equal(color, co)
    c.close()


def test_edit_sequence_matches_golden_and_oracle(ctx, case0):
    g = lu.golden()
    blocks, lights = case0
    for limit in (8, 4):
        b = blocks.copy()
        ctx.upload_world(b)
        ctx.lpv_repropagate(None, limit)
        lo, co = wb.lpv_repropagate(b, lights, limit)
        for k, e in enumerate(lu.edit_sequence(blocks, TABLE, 48, seed=5)):
            op, (x, y, z), blk, emissive = e
            lu.apply_edit(b, e)
            ctx.edit_blocks(np.array([[x, y, z, blk if op == 1 else 0]], dtype=np.int32))
            ctx.lpv_edit(op, (x, y, z), blk, limit)
            wb.lpv_edit(b, op, (x, y, z), blk, emissive, limit, lo, co)
            if k % 6 == 5 or k == 47:
                level, color = ctx.lpv_download()
                assert np.array_equal(level, lo) and np.array_equal(color, co), (limit, k, e)
                assert np.array_equal(lu.crc(level, color), g[f"edit_l{limit}_crc"][k]), (limit, k)
    level, color = ctx.lpv_download()
    assert np.array_equal(level, lu.dense(g["edit_l4_level_idx"], g["edit_l4_level_val"], blocks.shape))
    assert np.array_equal(color, lu.dense(g["edit_l4_color_idx"], g["edit_l4_color_val"], blocks.shape))


def test_edit_after_upload_and_argument_checks(ctx, case0):
    blocks, lights = case0
    b = blocks.copy()
    ctx.upload_world(b)
    lo, co = wb.lpv_repropagate(b, lights[::-1], 8)
    ctx.lpv_upload(lo, co)                     # Volumetrics::Reupload: volumes handed over by the caller
    e = (1, (int(lights[5][0]) + 2, int(lights[5][1]), int(lights[5][2])), 41, True)
    if b[e[1][2], e[1][1], e[1][0]] == 0:
        lu.apply_edit(b, e)
        ctx.edit_blocks(np.array([[*e[1], 41]], dtype=np.int32))
        ctx.lpv_edit(1, e[1], 41, 8)
        wb.lpv_edit(b, 1, e[1], 41, True, 8, lo, co)
    level, color = ctx.lpv_download()
    assert np.array_equal(level, lo) and np.array_equal(color, co)
    for bad in ((0, 5, 5), (5, 0, 5), (5, 5, 0), (384, 5, 5), (5, 128, 5), (5, 5, 384)):
        with pytest.raises(engine.VxrtError):
            ctx.lpv_edit(1, bad, 12, 8)
    with pytest.raises(engine.VxrtError):
        ctx.lpv_edit(2, (5, 5, 5), 12, 8)


def test_average_block_colors():
    """vxrt_cuda_lpv_average_colors (PrecomputeAverageBlockColor.comp): ten trilinear samples are exact arithmetic shared with the oracle,
    the final pow(x, 1.8) is CUDA's powf vs libm's: 4 float ulps."""
    import scene_util as su
    from oracle import binding as ob
    g = lu.golden()
    for size in (64, 512):
        inp = su.SceneInputs(size)
        c = engine.Context(0, (64, 32, 64))
        try:
            with pytest.raises(engine.VxrtError):
                c.lpv_average_colors()            # no albedo array yet
            inp.apply_to_context(c)
            got = c.lpv_average_colors()
        finally:
            c.close()
        want = g[f"average_colors_{size}"]
        assert np.array_equal(got == 0, want == 0)
        assert np.all(np.abs(got - want) <= 4 * np.spacing(np.abs(want))), np.abs(got - want).max()
        assert (got.view(np.uint32) == want.view(np.uint32)).mean() > 0.9


def test_sample_lpv_data(ctx, case0):
    """vxrt_cuda_lpv_sample (SampleLPVData, ReflectionTraceFrag.glsl:1516-1528): exact float arithmetic only, so bit-identical to the oracle
    and to the compiled shader function's output in the golden fixture"""
    import scene_util as su
    g = lu.golden()
    blocks, lights = case0
    ctx.upload_world(blocks)
    ctx.lpv_repropagate(None, 8)
    level, color = ctx.lpv_download()
    avg, pts, dithers = lu.sample_case(level)
    c = engine.Context(0)
    try:
        with pytest.raises(engine.VxrtError):
            c.lpv_sample(pts[:4])                    # no volume yet
        c.lpv_upload(level, color)
        with pytest.raises(engine.VxrtError):
            c.lpv_sample(pts[:4])                    # no colour table yet
        # the colour table comes from lpv_average_colors: give the context an albedo array whose averages we then replace by the seeded table
        su.SceneInputs(64).apply_to_context(c)
        c.lpv_average_colors()
        table = c.lpv_average_colors()
        for k, d in enumerate(dithers):
            got = c.lpv_sample(pts, d)
            want = wb.lpv_sample(level, color, table, pts, d)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), k
        c.lpv_set_average_colors(avg)                # the seeded table of the fixture
        for k, d in enumerate(dithers):
            got = c.lpv_sample(pts, d)
            assert np.array_equal(got.view(np.uint32), g["sample_rgb"][k].view(np.uint32)), k
        assert (got.sum(axis=1) > 0).mean() > 0.9
        assert c.lpv_sample(np.zeros((0, 3), np.float32)).shape == (0, 3)
    finally:
        c.close()
