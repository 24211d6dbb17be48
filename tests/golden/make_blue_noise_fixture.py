#!/usr/bin/env python
"""Extract the Heitz et al. 2019 blue-noise sampler tables (sobol_256spp_256d, scramblingTile, rankingTile)
from the literals of the reference's Core/BlueNoiseDataSSBO.cpp:4,9,14 into a compact uint8 fixture.

The GPU box has no reference tree, so the tables travel as tests/golden/blue_noise_tables.npz.  The layout
the passes consume is sobol ++ scramble ++ ranking as int32 (BlueNoiseDataSSBO.cpp:19-25)."""
import re
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/Core/BlueNoiseDataSSBO.cpp")
OUT = Path(__file__).resolve().parent / "blue_noise_tables.npz"


def main():
    text = REF.read_text()
    out = {}
    for name, n in (("sobol_256spp_256d", 256 * 256), ("scramblingTile", 128 * 128 * 8), ("rankingTile", 128 * 128 * 8)):
        m = re.search(name + r"\s*=\s*\{([^}]*)\}", text)
        vals = np.array([int(v) for v in m.group(1).split(",") if v.strip()], dtype=np.int64)
        assert vals.size == n and vals.min() >= 0 and vals.max() <= 255, (name, vals.size, vals.min(), vals.max())
        out[name] = vals.astype(np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    sys.exit(main())
