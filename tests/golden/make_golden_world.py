#!/usr/bin/env python
"""tests/golden/world_ref.npz + tests/data/synth_region/: outputs of the reference's own world producers, compiled in
oracle/_ref/libvxrt_ref_world.so by oracle/build_ref_world.py (FastNoise.cpp, enkimi.c, miniz.c as they are;
Core/WorldGenerator.cpp and Core/NBT/Importer.cpp lifted), on seeded inputs.  Run where /root/reference is mounted.

  noise_*      FastNoise::GetNoise at world_util.noise_points() for the three generator configurations of GenerateWorld
  gen_<k>      VoxelRT::GenerateWorld(gen_type = 1, structures off) for world_util.GEN_SEEDS[k]; gen_flat: gen_type = 0
  import_synth MCWorldImporter::ImportWorld over tests/data/synth_region (written by world_util.write_synth_regions, our
               own writer) with world_util.SYNTH_ORIGIN and world_util.mc_lut()
The grids are stored whole (np.savez_compressed: terrain compresses to a few hundred KB)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import world_util as wu  # noqa: E402
from oracle import world_binding as wb  # noqa: E402

if __name__ == "__main__":
    assert wb.ref_available(), "build oracle/_ref first (python oracle/build_ref_world.py)"
    wu.write_synth_regions()
    out = {}
    pts = wu.noise_points()
    out["noise_height"] = np.stack([wb.ref_fastnoise_2d(s, True, float(np.float32(0.00385)), 6, pts) for s, _ in wu.GEN_SEEDS])
    out["noise_biome"] = np.stack([wb.ref_fastnoise_2d(b, False, float(np.float32(0.01)), 3, pts) for _, b in wu.GEN_SEEDS])
    out["noise_stone"] = wb.ref_fastnoise_2d(77, False, float(np.float32(0.06)), 3, pts)
    for k, (s, b) in enumerate(wu.GEN_SEEDS):
        out[f"gen_{k}"] = wb.ref_generate_world(1, s, b)
    out["gen_flat"] = wb.ref_generate_world(0, 1, 2)
    out["import_synth"] = wb.ref_import_world(wu.SYNTH_DIR, wu.SYNTH_ORIGIN, wu.mc_lut())
    np.savez_compressed(wu.GOLD, **out)
    print("wrote", wu.GOLD, {k: (v.shape, int((v != 0).sum())) for k, v in out.items()})
