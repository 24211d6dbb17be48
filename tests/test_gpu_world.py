"""World producers on the GPU through the C ABI: vxrt_cuda_generate_world, vxrt_cuda_import_sections,
vxrt_cuda_collect_lights against the oracle and against the golden outputs of the reference's own code
(tests/golden/world_ref.npz).  Byte / index work: bit-exact."""
from types import SimpleNamespace

import numpy as np
import pytest

import world_util as wu
from oracle import binding as ob
from oracle import world_binding as wb
from voxeltracing_b200 import engine, host_api

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = engine.Context(0)
    yield c
    c.close()


def test_generate_world_matches_golden_and_oracle(ctx):
    g = wu.golden()
    for k, (s, b) in enumerate(wu.GEN_SEEDS):
        ctx.generate_world(1, s, b)
        assert np.array_equal(ctx.download_world(), g[f"gen_{k}"]), (s, b)
    ctx.generate_world(0, 1, 2)
    assert np.array_equal(ctx.download_world(), g["gen_flat"])
    for s, b, ids in ((17, 31337, (1, 2, 3, 5)), (49999, 0, (11, 22, 33, 44)), (-5, 123456789, (255, 2, 3, 5))):
        ctx.generate_world(1, s, b, *ids)
        assert np.array_equal(ctx.download_world(), wb.generate_world(1, s, b, ids)), (s, b)


def test_generate_world_other_dims():
    for dims in ((64, 64, 48), (128, 32, 16)):
        c = engine.Context(0, dims)
        c.generate_world(1, 4242, 999)
        assert np.array_equal(c.download_world(), wb.generate_world(1, 4242, 999, dims=dims)), dims
        c.close()


def test_generated_world_feeds_the_distance_field(ctx):
    s, b = wu.GEN_SEEDS[0]
    ctx.generate_world(1, s, b)
    with pytest.raises(engine.VxrtError):
        ctx.download_distance_field()          # a new world invalidates the field
    ctx.generate_distance_field()
    assert np.array_equal(ctx.download_distance_field(), ob.distance_field(wu.golden()["gen_0"]))
    with pytest.raises(engine.VxrtError):
        ctx.generate_world(1, 1, 2, grass=300)


def test_import_region_files_matches_golden(ctx):
    hs = host_api.RegionSections(wu.SYNTH_DIR)
    ctx.import_sections(hs, np.trunc(np.array(wu.SYNTH_ORIGIN)).astype(np.int32), wu.mc_lut())
    assert np.array_equal(ctx.download_world(), wu.golden()["import_synth"])


def test_import_random_sections_match_oracle(ctx):
    lut = wu.mc_lut()
    for seed, n, origin in ((3, 200, (0, 0, 0)), (4, 700, (37, -21, 5)), (5, 1, (-200, 100, 200))):
        s = wu.random_sections(seed, n)
        ctx.import_sections(s, origin, lut)
        assert np.array_equal(ctx.download_world(), wb.import_sections(s, origin, lut)), seed
    # additive import over an existing world
    base = wu.golden()["gen_1"]
    ctx.upload_world(base)
    s = wu.random_sections(6, 300)
    ctx.import_sections(s, (0, 0, 0), lut, clear_first=False)
    assert np.array_equal(ctx.download_world(), wb.import_sections(s, (0, 0, 0), lut, into=base.copy()))
    # an empty batch clears the world
    none = SimpleNamespace(block_ids=np.zeros((0, 4096), np.uint8), data_nibbles=np.zeros((0, 2048), np.uint8), has_data=np.zeros(0, np.uint8),
                           origins=np.zeros((0, 3), np.int32))
    ctx.import_sections(none, (0, 0, 0), lut)
    assert (ctx.download_world() == 0).all()


def test_import_needs_a_world_unless_it_clears():
    c = engine.Context(0, (32, 16, 48))
    with pytest.raises(engine.VxrtError):
        c.import_sections(wu.random_sections(1, 4), (0, 0, 0), wu.mc_lut(), clear_first=False)
    c.import_sections(wu.random_sections(1, 40), (0, 0, 0), wu.mc_lut())
    assert np.array_equal(c.download_world(), wb.import_sections(wu.random_sections(1, 40), (0, 0, 0), wu.mc_lut(), dims=(32, 16, 48)))
    c.close()


def test_collect_lights_matches_oracle(ctx):
    w = wu.golden()["gen_0"].copy()
    rng = np.random.default_rng(4)
    idx = rng.choice(w.size, 5000, replace=False)
    w.reshape(-1)[idx] = rng.choice(np.array([12, 41, 200, 3], dtype=np.uint8), 5000)
    w[0, 0, 0] = 12
    w[-1, -1, -1] = 41
    ctx.upload_world(w)
    bare = engine.Context(0, (32, 16, 48))
    with pytest.raises(engine.VxrtError):
        bare.collect_lights()                                           # no world yet
    bare.upload_world(np.ones((48, 16, 32), np.uint8))
    assert len(bare.collect_lights()) == 0                              # the block table starts out as "no emissive texture" everywhere
    bare.close()
    ctx.set_block_data(wu.emissive_table())
    want = wb.collect_lights(w, wu.emissive_table())
    got = ctx.collect_lights()
    assert len(want) > 1000 and np.array_equal(got, want)
    assert np.array_equal(ctx.collect_lights(capacity=10), want)        # grows to the reported count
    n = np.zeros(1, np.int32)
    few = np.zeros((7, 3), np.int32)
    import ctypes as C
    ctx._check(ctx._lib.vxrt_cuda_collect_lights(ctx._h, few.ctypes.data_as(C.c_void_p), 7, n.ctypes.data_as(C.POINTER(C.c_int32))))
    assert n[0] == len(want) and np.array_equal(few, want[:7])
    ctx.upload_world(np.zeros_like(w))
    assert len(ctx.collect_lights()) == 0
