#!/usr/bin/env python
"""Generate tests/golden/alpha_ref.npz from oracle/_ref: the reference's InitialRayTraceFrag.glsl / ShadowRayTraceFrag.glsl
compiled for the CPU with u_ShouldAlphaTest = true (VoxelTraversalDF_AlphaTest).  Run where /root/reference is mounted:

    python tests/golden/make_golden_alpha.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import alpha_util as au  # noqa: E402
from oracle import binding as ob  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402
from voxeltracing_b200 import host_api  # noqa: E402

OUT = Path(__file__).resolve().parent
FOV = 70.0
BLUE = np.random.default_rng(11).integers(0, 256, (256, 256, 4), dtype=np.uint8)


def main():
    assert rb.available("initial") and rb.available("shadow"), "build oracle/_ref first"
    w = host_api.gen_world("plains", 0)
    df = ob.distance_field(w)
    inp = au.alpha_inputs()
    rb.set_scene(w, df, inp.table, inp.blue, inp.textures, inp.sky)
    out = {}
    for pi, (pos, yaw, pitch) in enumerate(au.POSES):
        cam = host_api.camera(pos, yaw, pitch, au.W / au.H, FOV)
        g = rb.initial_trace(w, df, au.primary_params(cam, alpha=True, fov=FOV))
        for k in ("t", "normal", "block", "inv_t"):
            out[f"pose{pi}_{k}"] = g[k]
        s = rb.shadow_trace(w, df, au.shadow_params(cam, alpha=True, fov=FOV, soft=False, frame=3), g["t"], g["normal"], BLUE)
        out[f"pose{pi}_shadow"] = s["shadow"]
        out[f"pose{pi}_transversal"] = s["transversal"]
    np.savez_compressed(OUT / "alpha_ref.npz", **out)
    print("wrote", OUT / "alpha_ref.npz")


if __name__ == "__main__":
    main()
