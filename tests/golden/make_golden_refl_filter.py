#!/usr/bin/env python
"""tests/golden/refl_filter_ref.npz: outputs of Core/Shaders/SpecularTemporalFilter.glsl compiled through the GLSL shim
(oracle/_ref, vxref_specular_temporal) over the 5-frame sequence of tests/refl_filter_util.py on plains(seed=0), default flags.
R16F bit patterns of the three images of every frame's temporal set.  Run where /root/reference is mounted."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import refl_filter_util as rf  # noqa: E402
from oracle import binding as ob  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402
from voxeltracing_b200 import host_api  # noqa: E402

if __name__ == "__main__":
    assert rb.available("specular_temporal"), "build oracle/_ref first"
    L = rb.lib()
    L.vxref_specular_temporal.restype = None
    seq = rf.frames(host_api.gen_world("plains", 0))
    out = {}
    for k, o in enumerate(rf.run_chain(seq, lambda *x: ob.specular_temporal(*x, fn=L.vxref_specular_temporal))):
        for name in ("color", "frames", "hitdist"):
            out[f"{name}{k}"] = np.ascontiguousarray(o[name]).view(np.uint16)
    np.savez_compressed(Path(__file__).resolve().parent / "refl_filter_ref.npz", **out)
    print("wrote refl_filter_ref.npz", {k: v.shape for k, v in out.items() if k.endswith("0")})
