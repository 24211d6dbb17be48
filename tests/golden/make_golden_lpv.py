#!/usr/bin/env python3
"""tests/golden/lpv_ref.npz — outputs of the reference's own flood fill (Core/VolumetricFloodFill.cpp compiled in
oracle/_ref/libvxrt_ref_world.so, driven like Pipeline.cpp:1602-1611 and the block edit of World.cpp:273-485) on the seeded worlds
of tests/lpv_util.py, so that the GPU box (no reference tree) checks the oracle and the CUDA path against the reference's output.
Every case leaves the CRC-32 of both volumes and the number of lit voxels; the default-limit cases also the volumes themselves, sparse
(index + value of the non-zero voxels); every edit of the sequences the CRC-32 of both volumes after it.
Run here (needs /root/reference): python tests/golden/make_golden_lpv.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import lpv_util as lu  # noqa: E402
import world_util as wu  # noqa: E402
from oracle import world_binding as wb  # noqa: E402

CASES = [("rooms", 2, 400), ("plains", 1, 3000)]   # (world kind, seed, extra lamps)
LIMITS = (4, 8, 2, 3, 0, 11)
N_EDITS = 48


def main():
    assert wb.ref_available(), "build oracle/_ref first (python -m voxeltracing_b200.build)"
    table = wu.emissive_table()
    out = {}
    for c, (kind, seed, lamps) in enumerate(CASES):
        blocks = lu.lamp_world(seed, lamps, kind)
        lights = wb.collect_lights(blocks, table)
        out[f"c{c}_n_lights"] = np.int64(len(lights))
        for limit in LIMITS:
            level, color = wb.ref_lpv_repropagate(blocks, lights, limit, iterations=3)
            out[f"c{c}_l{limit}_crc"] = lu.crc(level, color)
            if limit == 4 and c == 0:
                out[f"c{c}_l{limit}_level_idx"], out[f"c{c}_l{limit}_level_val"] = lu.sparse(level)
                out[f"c{c}_l{limit}_color_idx"], out[f"c{c}_l{limit}_color_val"] = lu.sparse(color)
        # the light list in reverse order: the block-type volume follows the queue order, the level volume does not
        level, color = wb.ref_lpv_repropagate(blocks, lights[::-1], 8)
        out[f"c{c}_rev_crc"] = lu.crc(level, color)
    # edit sequence on case 0, limit 8 (the slider's maximum: the largest removals), then limit 4 (the default)
    kind, seed, lamps = CASES[0]
    blocks = lu.lamp_world(seed, lamps, kind)
    lights = wb.collect_lights(blocks, table)
    for limit in (8, 4):
        b = blocks.copy()
        level, color = wb.ref_lpv_repropagate(b, lights, limit)
        crcs = []
        for e in lu.edit_sequence(blocks, table, N_EDITS, seed=5):
            lu.apply_edit(b, e)
            wb.ref_lpv_edit(b, e[0], e[1], e[2], e[3], limit, level, color)
            crcs.append(lu.crc(level, color))
        out[f"edit_l{limit}_crc"] = np.stack(crcs)
        if limit == 4:
            out[f"edit_l{limit}_level_idx"], out[f"edit_l{limit}_level_val"] = lu.sparse(level)
            out[f"edit_l{limit}_color_idx"], out[f"edit_l{limit}_color_val"] = lu.sparse(color)
    # BlockAverageColorData of PrecomputeAverageBlockColor.comp (compiled through oracle/_ref) on the seeded test textures
    import scene_util as su
    from oracle import binding as ob
    from oracle import ref_binding as rb
    assert rb.available("lpv_average")
    for size in (64, 512):
        inp = su.SceneInputs(size)
        ow = ob.OracleWorld(np.zeros((16, 16, 16), np.uint8))
        rb.set_scene(np.zeros((384, 128, 384), np.uint8), np.zeros((384, 128, 384), np.uint8), inp.table, inp.blue, inp.textures, inp.sky)
        out[f"average_colors_{size}"] = rb.lpv_average_colors()
    # SampleLPVData of ReflectionTraceFrag.glsl (compiled through oracle/_ref) on case 0 at limit 8
    level, color = wb.ref_lpv_repropagate(blocks, lights, 8)
    avg, pts, dithers = lu.sample_case(level)
    out["sample_rgb"] = np.stack([wb.ref_lpv_sample(level, color, avg, pts, d) for d in dithers])
    np.savez_compressed(lu.GOLD, **out)
    print("wrote", lu.GOLD, lu.GOLD.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
