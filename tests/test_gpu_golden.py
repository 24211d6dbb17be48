"""CUDA path vs the committed reference golden vectors (made from the reference's own shaders)."""
import numpy as np
import pytest

import golden_util as gu
from voxeltracing_b200 import abi, engine, host_api

pytestmark = pytest.mark.gpu


def test_cuda_df_matches_reference_golden():
    _, hashes = gu.df_golden()
    c = engine.Context(0)
    for name, w in gu.worlds().items():
        c.upload_world(w)
        c.generate_distance_field()
        assert gu.sha(c.download_distance_field()) == hashes[name], name
    c.close()


def test_cuda_primary_and_shadow_match_reference_golden():
    g = gu.trace_golden()
    ws = gu.worlds()
    c = engine.Context(0)
    c.set_blue_noise_texture(gu.BLUE)
    W, H = gu.mg.W, gu.mg.H
    light = host_api.sun_direction(50.0)[2]
    for wname in ("plains0", "rooms2"):
        c.upload_world(ws[wname])
        c.generate_distance_field()
        for pi, (pos, yaw, pitch) in enumerate(gu.mg.POSES):
            cam = host_api.camera(pos, yaw, pitch, W / H)
            jitter = host_api.taa_jitter(5) if pi == 1 else None
            c.initial_trace(cam, W, H, 350, jitter)
            key = f"{wname}_pose{pi}"
            for att, k in ((abi.ATT_INITIAL_T, "t"), (abi.ATT_INITIAL_NORMAL, "normal"), (abi.ATT_INITIAL_BLOCK, "block"), (abi.ATT_INITIAL_INVT, "inv_t")):
                assert np.array_equal(c.read_attachment(att).view(np.uint8), g[f"{key}_{k}"].view(np.uint8)), (key, k)
            # hard shadows: no transcendental on the path -> bit exact
            c.shadow_trace(cam, W, H, light, frame=0, soft=False)
            assert np.array_equal(c.read_attachment(abi.ATT_SHADOW), g[f"{key}_shadow0"])
            assert np.array_equal(c.read_attachment(abi.ATT_SHADOW_TRANSVERSAL).view(np.uint16), g[f"{key}_transversal0"].view(np.uint16))
            # soft shadows: sinf/cosf of the cone sample differ by <= 1-2 ulp between CUDA and libm
            c.shadow_trace(cam, W, H, light, frame=7, soft=True)
            same = c.read_attachment(abi.ATT_SHADOW) == g[f"{key}_shadow1"]
            assert same.mean() >= 0.999, (key, same.mean())
    c.close()
