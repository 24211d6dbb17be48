"""World producers, CPU side: the oracle restatement (oracle/vxrt_oracle_world.cpp) and the host region
reader (voxeltracing_b200/host/vxrt_mca.cpp) against the reference's own code compiled in oracle/_ref
(FastNoise, enkiMI, WorldGenerator.cpp, Importer.cpp; only where /root/reference is mounted) and against the golden
outputs of that build (tests/golden/world_ref.npz, everywhere).  Integer / byte work: bit-exact."""
import glob
import os
from pathlib import Path

import numpy as np
import pytest

import world_util as wu
from oracle import world_binding as wb
from voxeltracing_b200 import host_api

REF = Path(os.environ.get("VXRT_REFERENCE", "/root/reference"))
needs_ref = pytest.mark.skipif(not (wb.ref_available() and REF.exists()), reason="oracle/_ref world library / reference tree not present")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_noise_matches_golden():
    g, pts = wu.golden(), wu.noise_points()
    for k, (s, b) in enumerate(wu.GEN_SEEDS):
        assert np.array_equal(bits(wb.fastnoise_2d(s, True, float(np.float32(0.00385)), 6, pts)), bits(g["noise_height"][k]))
        assert np.array_equal(bits(wb.fastnoise_2d(b, False, float(np.float32(0.01)), 3, pts)), bits(g["noise_biome"][k]))
    assert np.array_equal(bits(wb.fastnoise_2d(77, False, float(np.float32(0.06)), 3, pts)), bits(g["noise_stone"]))


@needs_ref
def test_noise_matches_fastnoise():
    rng = np.random.default_rng(1)
    pts = (rng.random((50_000, 2), dtype=np.float32) * 4000.0 - 2000.0).astype(np.float32)
    for seed in (0, 1337, 49999, -5):
        for fractal, freq, octaves in ((True, 0.00385, 6), (False, 0.01, 3), (True, 0.06, 16), (True, 1.0, 1)):
            f = float(np.float32(freq))
            assert np.array_equal(bits(wb.fastnoise_2d(seed, fractal, f, octaves, pts)), bits(wb.ref_fastnoise_2d(seed, fractal, f, octaves, pts)))


def test_noise_range_and_mean():
    # self-check: simplex noise stays in [-1, 1] and is roughly centred
    rng = np.random.default_rng(2)
    v = wb.fastnoise_2d(9, False, 0.01, 3, (rng.random((100_000, 2), dtype=np.float32) * 1e4).astype(np.float32))
    assert v.min() >= -1.0 and v.max() <= 1.0 and abs(float(v.mean())) < 0.05


def test_generate_world_matches_golden():
    g = wu.golden()
    for k, (s, b) in enumerate(wu.GEN_SEEDS):
        assert np.array_equal(wb.generate_world(1, s, b), g[f"gen_{k}"])
    assert np.array_equal(wb.generate_world(0, 1, 2), g["gen_flat"])


@needs_ref
def test_generate_world_matches_reference():
    for s, b in ((17, 31337), (49999, 0)):
        assert np.array_equal(wb.generate_world(1, s, b), wb.ref_generate_world(1, s, b))
    # other block ids than the default database's
    ids8 = (11, 22, 33, 44, 6, 7, 8, 4)
    assert np.array_equal(wb.generate_world(1, 5, 6, ids=ids8[:4]), wb.ref_generate_world(1, 5, 6, ids8=ids8))


def test_generate_world_structure():
    w = wb.generate_world(1, *wu.GEN_SEEDS[0])
    solid = w != 0
    height = solid.sum(axis=1)                      # columns are filled from y = 0 without holes
    assert np.array_equal(solid, np.arange(128)[None, :, None] < height[:, None, :])
    assert height.min() >= 8 and height.max() <= 48
    top = np.take_along_axis(w, (height - 1)[:, None, :], axis=1)[:, 0, :]
    assert set(np.unique(top)) <= {1, 5}            # grass or sand on top
    assert set(np.unique(w)) <= {0, 1, 2, 3, 5}


def test_region_reader_matches_python_restatement():
    files = sorted(glob.glob(str(wu.SYNTH_DIR / "*.mca")))
    hs, ps = host_api.RegionSections(wu.SYNTH_DIR), wb.PySections(files)
    assert len(hs) == len(ps) == 70 and hs.chunks == ps.chunks == 24   # chunks with an empty Sections list count, sections without Blocks do not
    for k in ("block_ids", "data_nibbles", "has_data", "origins"):
        assert np.array_equal(getattr(hs, k), getattr(ps, k)), k
    assert (hs.has_data == 0).any() and (hs.origins[:, 1] < 0).any() and hs.palette_sections == 0 and hs.bad_chunks == 0
    one = host_api.RegionSections(files[0])
    assert 0 < len(one) < len(hs)
    with pytest.raises(OSError):
        host_api.RegionSections(wu.SYNTH_DIR / "missing.mca")


def test_import_matches_golden():
    hs = host_api.RegionSections(wu.SYNTH_DIR)
    got = wb.import_sections(hs, np.trunc(np.array(wu.SYNTH_ORIGIN)).astype(np.int32), wu.mc_lut())
    want = wu.golden()["import_synth"]
    assert (want != 0).sum() > 50_000 and np.array_equal(got, want)


@needs_ref
def test_import_matches_reference_on_the_engine_test_world():
    d = REF / "Test MC Worlds" / "Medival"
    if not d.exists():
        pytest.skip("Test MC Worlds/Medival not present")
    db = host_api.BlockDatabase(str(REF / "blockdb.txt"))
    lut = db.minecraft_lut()
    assert lut[0] == 0 and lut[255] == db.block_id("INVALID_BLOCK") != 0
    hs = host_api.RegionSections(d)
    ps = wb.PySections(sorted(glob.glob(str(d / "*.mca"))))
    assert len(hs) == len(ps) == 2544 and np.array_equal(hs.block_ids, ps.block_ids) and np.array_equal(hs.origins, ps.origins)
    for origin in ((0.0, 0.0, -400.0), (-384.5, 2.0, -200.25), (136.0, 0.0, -94.0)):
        want = wb.ref_import_world(d, origin, lut)
        got = wb.import_sections(hs, np.trunc(np.array(origin)).astype(np.int32), lut)
        assert np.array_equal(got, want)
    assert (want == 0).all()   # the origin of Origin.txt puts the grid where this world has no chunks


def test_import_edge_cases():
    lut = wu.mc_lut()
    from types import SimpleNamespace
    none = SimpleNamespace(block_ids=np.zeros((0, 4096), np.uint8), data_nibbles=np.zeros((0, 2048), np.uint8), has_data=np.zeros(0, np.uint8),
                           origins=np.zeros((0, 3), np.int32))
    assert (wb.import_sections(none, (0, 0, 0), lut) == 0).all()
    s = wu.random_sections(3, 200)
    a = wb.import_sections(s, (0, 0, 0), lut)
    assert (a != 0).any()
    # additive import keeps what is there where the batch writes nothing
    base = np.full((384, 128, 384), 77, dtype=np.uint8)
    b = wb.import_sections(s, (0, 0, 0), lut, into=base.copy())
    assert np.array_equal(b[a != 0], a[a != 0]) and (b[a == 0] == 77).all()
    # a brute-force numpy restatement of the scatter for one section
    one = SimpleNamespace(block_ids=s.block_ids[:1], data_nibbles=s.data_nibbles[:1], has_data=np.ones(1, np.uint8), origins=np.array([[16, 32, -48]], np.int32))
    got = wb.import_sections(one, (5, -3, 7), lut)
    want = np.zeros_like(got)
    ids = one.block_ids[0].reshape(16, 16, 16)
    nib = one.data_nibbles[0]
    dv = np.stack([nib & 15, nib >> 4], axis=1).reshape(16, 16, 16)
    for y in range(16):
        for z in range(16):
            for x in range(16):
                v = lut[ids[y, z, x]]
                X, Y, Z = 16 + x - 5 + 192, 32 + y + 3, -48 + z - 7 + 192
                if dv[y, z, x] == 0 and v != 0 and 0 <= X < 384 and 0 <= Y < 128 and 0 <= Z < 384:
                    want[Z, Y, X] = v
    assert np.array_equal(got, want)


def test_collect_lights():
    w = wb.generate_world(1, *wu.GEN_SEEDS[0])
    rng = np.random.default_rng(4)
    idx = rng.choice(w.size, 5000, replace=False)
    w.reshape(-1)[idx] = rng.choice(np.array([12, 41, 200, 3], dtype=np.uint8), 5000)
    t = wu.emissive_table()
    got = wb.collect_lights(w, t)
    flat = np.flatnonzero(np.isin(w.reshape(-1), [12, 41]))
    want = np.stack([flat % 384, (flat // 384) % 128, flat // (384 * 128)], axis=1)
    assert len(got) == len(want) > 1000 and np.array_equal(got, want)


def test_minecraft_lut_follows_get_id_from_mcid():
    db = host_api.BlockDatabase(str(Path(__file__).parent / "data" / "blockdb_synth.txt"))
    lut = db.minecraft_lut()
    assert lut[0] == 0
    unlisted = db.block_id("INVALID_BLOCK")   # 0 when the database has no such record, like ParsedBlockDataList[...] default
    listed = lut != unlisted
    assert (lut[~listed] == unlisted).all()
