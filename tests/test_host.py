"""Host-side producers: camera matrices, jitter, sun direction, worlds, raw save format (CPU only)."""
import numpy as np

from voxeltracing_b200 import host_api


def test_camera_matches_glm_formulas():
    cam = host_api.camera([192, 75, 192], 90.0, 0.0, 16 / 9)
    P = cam.projection.reshape(4, 4).T  # column-major -> math layout
    t = np.tan(np.radians(90.0) / 2)
    assert np.isclose(P[0, 0], 1 / (16 / 9 * t)) and np.isclose(P[1, 1], 1 / t)
    assert np.isclose(P[2, 2], -(1000 + 0.1) / (1000 - 0.1)) and P[3, 2] == -1.0
    assert np.isclose(P[2, 3], -(2 * 1000 * 0.1) / (1000 - 0.1))
    V, Vi = cam.view.reshape(4, 4).T, cam.inv_view.reshape(4, 4).T
    assert np.allclose(V @ Vi, np.eye(4), atol=1e-5)
    assert np.allclose(Vi[:3, 3], [192, 75, 192])
    assert np.allclose(cam.projection.reshape(4, 4).T @ cam.inv_projection.reshape(4, 4).T, np.eye(4), atol=1e-4)
    # yaw 90 looks down +Z: view-space -Z axis maps to world +Z
    assert np.allclose(Vi[:3, :3] @ [0, 0, -1], [0, 0, 1], atol=1e-6)
    # inverse agrees with numpy for an arbitrary pose
    cam = host_api.camera([10.5, 99.25, 300.0], 33.0, -20.0, 4 / 3)
    assert np.allclose(cam.inv_view.reshape(4, 4).T, np.linalg.inv(cam.view.reshape(4, 4).T.astype(np.float64)), atol=1e-4)


def test_taa_jitter_is_halton_2_3():
    def halton(i, b):
        f, r = 1.0, 0.0
        while i > 0:
            f /= b
            r += f * (i % b)
            i //= b
        return r

    for frame in (0, 1, 5, 63, 64, 200):
        j = host_api.taa_jitter(frame)
        assert np.allclose(j, [halton(frame % 64 + 1, 2), halton(frame % 64 + 1, 3)], atol=1e-6)


def test_sun_direction_default_tick():
    sun, moon, strong = host_api.sun_direction(50.0)
    assert np.allclose(sun, [-0.6688, 0.4683, 0.5774], atol=1e-3)
    assert np.allclose(moon, [-sun[0], -sun[1], sun[2]]) and np.allclose(strong, sun)
    _, moon, strong = host_api.sun_direction(140.0)  # sun below the horizon -> moon is the stronger light
    assert np.allclose(strong, moon)


def test_worlds_are_deterministic_and_plausible(tmp_path):
    a = host_api.gen_world("plains", 0)
    b = host_api.gen_world("plains", 0)
    c = host_api.gen_world("plains", 1)
    assert a.shape == (384, 128, 384) and np.array_equal(a, b) and not np.array_equal(a, c)
    assert a[:, 0, :].all() and not a[:, 100:, :].any()
    assert set(np.unique(a)) <= {0, 1, 2, 3, 5, 6, 7}
    rooms = host_api.gen_world("rooms", 2)
    assert (rooms == 12).any() or (rooms == 27).any() or (rooms == 26).any()  # emissive lamps
    town = host_api.gen_world("town", 3)
    assert (town == 8).any() or (town == 4).any()
    # raw headerless dump round trip (WorldFileHandler.cpp:26,50)
    path = tmp_path / "w"
    host_api.save_world(str(path), a)
    assert path.stat().st_size == 18874368
    import os

    os.environ["VXRT_WORLDS"] = str(tmp_path)
    try:
        back, desc = host_api.load_named_world("w", "plains", 9)
        assert desc == "file:w" and np.array_equal(back, a)
        _, desc = host_api.load_named_world("missing", "plains", 9)
        assert desc.startswith("stand-in:")
    finally:
        del os.environ["VXRT_WORLDS"]


def test_random_edits_toggle_and_are_reproducible():
    w = host_api.gen_world("flat", 0)
    w2 = w.copy()
    e = host_api.random_edits(w2, 1024, 1234)
    assert e.shape == (1024, 4) and (e[:, 0] >= 1).all() and (e[:, 0] <= 382).all() and (e[:, 1] <= 126).all()
    w3 = w.copy()
    for x, y, z, i in e:
        w3[z, y, x] = i
    assert np.array_equal(w2, w3) and not np.array_equal(w, w2)
    w4 = w.copy()
    assert np.array_equal(host_api.random_edits(w4, 1024, 1234), e)
