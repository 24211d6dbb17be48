#!/usr/bin/env python
"""Generate tests/golden/*.npz from oracle/_ref — the reference's own shaders compiled for the CPU
(oracle/build_ref.py).  Run in the container where /root/reference is mounted:

    python tests/golden/make_golden.py

The fixtures are small (a few hundred KB) and let every box — including the GPU box, where the
reference tree does not exist — check the oracle and the CUDA path against the reference's output.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_binding as rb  # noqa: E402
from voxeltracing_b200 import abi, host_api  # noqa: E402

OUT = Path(__file__).resolve().parent
POSES = [([192, 75, 192], 45.0, -20.0), ([192, 75, 192], 200.0, -25.0), ([-60, 150, -40], 45.0, -30.0), ([100.5, 60.25, 300.75], 300.0, 10.0)]
W, H = 160, 90


def fill(dst, src):
    for i, v in enumerate(np.asarray(src, np.float32).ravel()):
        dst[i] = float(v)


def primary_params(cam, jitter=None):
    p = abi.PrimaryParams()
    fill(p.inv_view, cam.inv_view); fill(p.inv_projection, cam.inv_projection)
    p.width, p.height, p.render_distance = W, H, 350
    if jitter is not None:
        p.jitter[0], p.jitter[1], p.jitter_on = float(jitter[0]), float(jitter[1]), 1
    return p


def shadow_params(cam, frame, soft):
    s = abi.ShadowParams()
    fill(s.inv_view, cam.inv_view); fill(s.inv_projection, cam.inv_projection)
    s.width, s.height = W, H
    fill(s.light_direction, host_api.sun_direction(50.0)[2])
    s.current_frame, s.soft_shadows, s.max_iterations = frame, int(soft), 350
    return s


def main():
    assert rb.available("df") and rb.available("initial") and rb.available("shadow"), "build oracle/_ref first"
    worlds = {"plains0": host_api.gen_world("plains", 0), "rooms2": host_api.gen_world("rooms", 2)}
    edited = worlds["plains0"].copy()
    host_api.random_edits(edited, 1024, 1234)
    worlds["plains0_edited1024"] = edited
    df_hashes = {}
    dfs = {}
    for name, w in worlds.items():
        dfs[name] = rb.distance_field(w)
        df_hashes[name] = hashlib.sha256(dfs[name].tobytes()).hexdigest()
    # a few z-slices of the field itself, so a mismatch can be localised
    np.savez_compressed(OUT / "df_ref.npz", hashes=np.array([f"{k}:{v}" for k, v in df_hashes.items()]),
                        plains0_z192=dfs["plains0"][192], plains0_edited1024_z100=dfs["plains0_edited1024"][100],
                        rooms2_z190=dfs["rooms2"][190])
    blue = np.random.default_rng(11).integers(0, 256, (256, 256, 4), dtype=np.uint8)
    out = {}
    for wi, wname in enumerate(("plains0", "rooms2")):
        w, df = worlds[wname], dfs[wname]
        for pi, (pos, yaw, pitch) in enumerate(POSES):
            cam = host_api.camera(pos, yaw, pitch, W / H)
            jitter = host_api.taa_jitter(5) if pi == 1 else None
            g = rb.initial_trace(w, df, primary_params(cam, jitter))
            key = f"{wname}_pose{pi}"
            for k in ("t", "normal", "block", "inv_t"):
                out[f"{key}_{k}"] = g[k]
            for soft, frame in ((0, 0), (1, 7)):
                s = rb.shadow_trace(w, df, shadow_params(cam, frame, soft), g["t"], g["normal"], blue)
                out[f"{key}_shadow{soft}"] = s["shadow"]
                out[f"{key}_transversal{soft}"] = s["transversal"]
    np.savez_compressed(OUT / "trace_ref.npz", **out)
    print("wrote", OUT / "df_ref.npz", OUT / "trace_ref.npz")


if __name__ == "__main__":
    main()
