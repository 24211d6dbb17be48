#!/usr/bin/env python
"""tests/golden/pick_ref.npz: World::RaycastDetect of the reference (Core/World.cpp, lifted into oracle/_ref by
oracle/build_ref.py) on 20,000 seeded rays over plains(seed=0).  Run where /root/reference is mounted."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import pick_util as pu  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402
from voxeltracing_b200 import host_api  # noqa: E402

if __name__ == "__main__":
    assert rb.available("raycast"), "build oracle/_ref first"
    o, d = pu.pick_rays(20_000, 7)
    hits = rb.raycast_detect(host_api.gen_world("plains", 0), o, d)
    np.savez_compressed(Path(__file__).resolve().parent / "pick_ref.npz", positions=o, directions=d, hits=hits)
    print("wrote pick_ref.npz", (hits[:, 0] >= 0).mean())
