#!/usr/bin/env python
"""Generate tests/golden/shade_ref.npz from oracle/_ref: the reference's own GenerateGBuffer.glsl, ColorPassFrag.glsl (direct
term), DiffuseRayTraceFrag.glsl and ReflectionTraceFrag.glsl compiled for the CPU (oracle/build_ref.py) and dispatched like
Pipeline.cpp does.  Run in the container where /root/reference is mounted:

    python tests/golden/make_golden_shade.py

SURVEY 8 rows a8 - a11.  Every pass of a case is fed the reference's own outputs of the passes before it, so the fixture is one
self-consistent frame of the reference per case; tests/shade_golden_util.py rebuilds the same inputs anywhere.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import shade_golden_util as sg  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402

OUT = Path(__file__).resolve().parent


class RefBackend:
    """The four shaders + primary / shadow passes through oracle/_ref."""

    def __init__(self, blocks, df, inputs):
        self.blocks, self.df = blocks, df
        rb.set_scene(blocks, df, inputs.table, inputs.blue, inputs.textures, inputs.sky)

    def initial_trace(self, p):
        return rb.initial_trace(self.blocks, self.df, p)

    def shadow_trace(self, p, g_t, g_n):
        return rb.shadow_trace(self.blocks, self.df, p, g_t, g_n, None)

    generate_gbuffer = staticmethod(rb.generate_gbuffer)
    shade_direct = staticmethod(rb.shade_direct)
    diffuse_trace = staticmethod(rb.diffuse_trace)
    reflection_trace = staticmethod(rb.reflection_trace)
    set_lpv = staticmethod(rb.set_lpv)


def main():
    for what in ("df", "initial", "shadow", "gbuffer", "diffuse", "reflection", "color"):
        assert rb.available(what), f"oracle/_ref lacks {what}: build it first (python oracle/build_ref.py)"
    out = {}
    inputs = sg.inputs()
    for wname in sg.WORLDS:
        blocks = sg.world(wname)
        be = RefBackend(blocks, rb.distance_field(blocks), inputs)
        for case in sg.CASES:
            if case["world"] != wname:
                continue
            res = sg.run_case(be, case, inputs)
            for k, v in res.items():
                out[f"{case['name']}_{k}"] = v
    np.savez_compressed(OUT / "shade_ref.npz", **out)
    print("wrote", OUT / "shade_ref.npz", (OUT / "shade_ref.npz").stat().st_size, "bytes,", len(out), "arrays")
    # ApproximateGILPV: the reflection outputs with u_LPVGI on (the volume itself comes from the flood-fill restatement, which
    # tests/test_oracle_lpv.py pins against the reference's VolumetricFloodFill.cpp)
    case = sg.LPV_CASE
    blocks = sg.world(case["world"])
    be = RefBackend(blocks, rb.distance_field(blocks), inputs)
    res = sg.run_case(be, case, inputs)
    lpv = {f"{case['name']}_{k}": v for k, v in res.items() if k.startswith("refl")}
    np.savez_compressed(OUT / "shade_lpv_ref.npz", **lpv)
    print("wrote", OUT / "shade_lpv_ref.npz", (OUT / "shade_lpv_ref.npz").stat().st_size, "bytes,", len(lpv), "arrays")


if __name__ == "__main__":
    main()
