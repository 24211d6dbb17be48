"""Oracle self-checks for the traversal (SURVEY.md §8c (2),(3)) and format helpers; CPU only."""
import numpy as np

from oracle import binding as ob


def _rays(n, seed, dims):
    rng = np.random.default_rng(seed)
    nx, ny, nz = dims
    o = np.stack([rng.uniform(1, nx - 1, n), rng.uniform(ny * 0.55, ny - 1, n), rng.uniform(1, nz - 1, n)], 1)
    d = rng.normal(size=(n, 3))
    d[:, 1] = -np.abs(d[:, 1]) - 0.05
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o.astype(np.float32), d.astype(np.float32)


def test_df_dda_hits_same_voxel_as_plain_dda(plains0, plains0_oracle):
    ow = plains0_oracle
    o, d = _rays(4000, 7, (384, 128, 384))
    same = total = 0
    for i in range(len(o)):
        start = np.floor(o[i]).astype(int)
        if plains0[start[2], start[1], start[0]] != 0:
            h = ow.traverse(o[i], d[i], 350)
            assert h.t == -1.0  # a ray that starts inside a solid voxel misses
            continue
        h = ow.traverse(o[i], d[i], 350)
        hit, vox = ow.plain_dda(o[i], d[i], 4000)
        if h.iterations >= 350:
            continue
        assert (h.t > 0) == hit
        if hit:
            total += 1
            end = np.floor(np.array(h.end[:])).astype(int)
            same += int((end == vox).all())
            # the face normal opposes the ray on the hit axis
            n = np.array(h.normal[:])
            assert abs(n).sum() == 1.0 and np.dot(n, d[i]) < 0
            # t is the distance travelled (up to the accumulated 1e-4 nudges)
            assert abs(h.t - np.linalg.norm(np.array(h.end[:]) - o[i])) < 1e-3
    assert total > 1000
    # fp32 nudges may flip a grazing edge; everything else must agree with the exact walk
    assert same / total > 0.999


def test_step_table():
    """The CUDA loop replaces floor(float(k) * 0.57735026918f) by (k * 9459) >> 14 (csrc/traverse.cuh): both must be the
    oracle's table for every byte, with k == 1 -> 1 (InitialRayTraceFrag.glsl:89-92,331-333)."""
    t = ob.step_table()
    k = np.arange(256)
    want = np.floor(k.astype(np.float32) * np.float32(0.57735026918)).astype(np.int64)
    want[1] = 1
    assert np.array_equal(t, want)
    fast = (k * 9459) >> 14
    fast[1] = 1
    assert np.array_equal(t, fast)
    assert set(np.nonzero(t == 1)[0]) == {1, 2, 3} and t[0] == 0 and t[4] == 2 and t[254] == 146


def test_batch_traverse_equals_single_ray_calls(plains0_oracle):
    o, d = _rays(300, 11, (384, 128, 384))
    hits = plains0_oracle.traverse_batch(o, d, 350)
    for i in range(len(o)):
        h = plains0_oracle.traverse(o[i], d[i], 350)
        assert hits["t"][i].tobytes() == np.float32(h.t).tobytes() and hits["block"][i] == h.block
        assert hits["iterations"][i] == h.iterations and list(hits["end"][i]) == list(h.end[:])


def test_iteration_cap_and_leaving_volume_miss(plains0_oracle):
    ow = plains0_oracle
    h = ow.traverse([192.0, 120.0, 192.0], [0.0, 1.0, 0.0], 350)  # straight up and out
    assert h.t == -1.0 and h.intersection == 0
    h = ow.traverse([192.3, 120.0, 192.6], [0.0, -1.0, 0.0], 1)  # cap reached in empty space
    assert h.t == -1.0 and h.iterations == 1
    h = ow.traverse([192.3, 120.0, 192.6], [0.0, -1.0, 0.0], 350)
    assert h.t > 0 and list(h.normal[:]) == [0.0, 1.0, 0.0] and h.block in (1, 5, 7)


def test_half_and_unorm_conversions_match_numpy():
    rng = np.random.default_rng(3)
    vals = np.concatenate([
        rng.uniform(-70000, 70000, 4000), rng.uniform(-1, 1, 4000) * 1e-4, rng.uniform(-1, 1, 2000) * 6e-8,
        np.array([0.0, -0.0, 1.0, -1.0, 65504.0, 65519.9, 65520.0, 1e9, -1e9, 2.0 ** -24, 2.0 ** -25, 3 * 2.0 ** -26,
                  np.inf, -np.inf, 0.0425, 64.0, 196.0, 1e-5]),
    ]).astype(np.float32)
    L = ob.lib()
    got = np.array([L.vxo_float_to_half(float(v)) for v in vals], dtype=np.uint16)
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    assert np.array_equal(got, want)
    hs = np.arange(0, 65536, 7, dtype=np.uint16)
    back = np.array([L.vxo_half_to_float(int(h)) for h in hs], dtype=np.float32)
    ref = hs.view(np.float16).astype(np.float32)
    ok = np.isnan(ref) | (back == ref)
    assert ok.all()
    # face ids as the R8 target stores them decode with round(n*10).  The pinned
    # float->unorm8 rule is RNE(f*255.0f) in fp32: 0.3f*255 rounds to the tie 76.5 -> 76.
    codes = [L.vxo_float_to_unorm8(np.float32(i) / np.float32(10.0)) for i in range(6)] + [L.vxo_float_to_unorm8(1.0)]
    assert codes == [0, 26, 51, 76, 102, 128, 255]
    assert [int(np.rint(np.float32(c) / np.float32(255.0) * np.float32(10.0))) for c in codes] == [0, 1, 2, 3, 4, 5, 10]
