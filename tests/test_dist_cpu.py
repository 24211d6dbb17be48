"""world_size-2 gloo tests of the multi-GPU host logic (screen-tile band gather, z-slab distance field with
one boundary exchange) on CPU."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_band_rows_partition_the_frame():
    from voxeltracing_b200.pipeline import band_rows

    for h in (1080, 2160, 360, 92, 7):
        for world in (1, 2, 3, 4, 8):
            covered = np.zeros(h, int)
            for r in range(world):
                row0, rows = band_rows(h, r, world)
                assert row0 % 8 == 0 and rows >= 0
                covered[row0:row0 + rows] += 1
            assert (covered == 1).all(), (h, world)


def test_slab_bounds():
    from voxeltracing_b200.sharding import slab_bounds

    assert slab_bounds(384, 8) == [0, 48, 96, 144, 192, 240, 288, 336, 384]
    assert slab_bounds(10, 3) == [0, 4, 7, 10]


def test_two_rank_gloo_bands_and_sharded_distance_field():
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "dist_worker.py")], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    line = [ln for ln in outs[0].splitlines() if ln.startswith("RESULT ")]
    assert line, outs[0]
    res = json.loads(line[0][7:])
    assert res["bands_t_equal"] and res["bands_block_equal"] and res["df_equal"] and res["all_ranks_ok"], res
    assert res["slabs"] == [0, 48, 96]
