#!/usr/bin/env python
"""tests/golden/refl_filter_ref.npz: outputs of Core/Shaders/SpecularTemporalFilter.glsl compiled through the GLSL shim
(oracle/_ref, vxref_specular_temporal) over the 5-frame sequence of tests/refl_filter_util.py on plains(seed=0), default flags:
R16F bit patterns of the three images of every frame's temporal set; and of ReflectionDenoiserNew.glsl (vxref_reflection_denoise),
x and y pass, on the temporal sets of frames 1, 3 and 4.  Run where /root/reference is mounted."""
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
    L.vxref_reflection_denoise.restype = None
    sets = rf.run_chain(seq, lambda *x: ob.specular_temporal(*x, fn=L.vxref_specular_temporal))
    for k in (1, 3, 4):   # ReflectionDenoiserNew.glsl, x then y pass, on the temporal set of frame k
        x, y = rf.run_denoise(seq[k], sets[k], rf.sets_for(k)[1], lambda *a: ob.reflection_denoise(*a, fn=L.vxref_reflection_denoise))
        out[f"denoise_x{k}"] = np.ascontiguousarray(x).view(np.uint16)
        out[f"denoise_y{k}"] = np.ascontiguousarray(y).view(np.uint16)
    np.savez_compressed(Path(__file__).resolve().parent / "refl_filter_ref.npz", **out)
    print("wrote refl_filter_ref.npz", {k: v.shape for k, v in out.items() if k.endswith("0")})
