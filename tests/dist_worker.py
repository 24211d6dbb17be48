"""Worker of the world_size-2 gloo tests (tests/test_dist_cpu.py): exercises the multi-GPU host logic of
voxeltracing_b200/sharding.py on CPU tensors, with the oracle standing in for the per-rank renderer."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import binding as ob  # noqa: E402
from voxeltracing_b200 import abi, host_api, sharding  # noqa: E402
from voxeltracing_b200.pipeline import band_rows  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ob.set_threads(2)
    out = {}

    # ---- screen-tile sharding: each rank renders its band of rows, bands gathered to rank 0 ----
    blocks = host_api.gen_world("plains", 0)
    ow = ob.OracleWorld(blocks)
    W, H = 160, 92  # 92 rows: bands of 8 do not divide evenly
    cam = host_api.camera([192, 75, 192], 45.0, -20.0, W / H)
    p = abi.PrimaryParams()
    for i in range(16):
        p.inv_view[i] = float(cam.inv_view[i]); p.inv_projection[i] = float(cam.inv_projection[i])
    p.width, p.height, p.render_distance = W, H, 350
    row0, rows = band_rows(H, rank, world)
    abi.set_tile(p.tile, (row0, rows))
    part = ow.initial_trace(p)
    full_t = torch.from_numpy(part["t"].view(np.int16).copy())
    full_b = torch.from_numpy(part["block"].copy())
    sharding.gather_bands(full_t, H, rank, world)
    sharding.gather_bands(full_b, H, rank, world)
    if rank == 0:
        abi.set_tile(p.tile, (0, 0))
        whole = ow.initial_trace(p)
        out["bands_t_equal"] = bool(np.array_equal(full_t.numpy().view(np.float16).view(np.uint16), whole["t"].view(np.uint16)))
        out["bands_block_equal"] = bool(np.array_equal(full_b.numpy(), whole["block"]))
        out["band_rows"] = [band_rows(H, r, world) for r in range(world)]

    # ---- z-slab sharded distance field: one boundary exchange + slab all-gather ----
    small = host_api.gen_world("plains", 3)[:, :, :][:96, :48, :64].copy()
    small[40:44, 20:30, 10:50] = 3
    backend = sharding.NumpySlabBackend(small)
    z0 = sharding.regenerate_distance_field_sharded(backend, rank, world)
    want = ob.distance_field(small)
    out["df_equal"] = bool(np.array_equal(backend.df.numpy(), want))
    out["slabs"] = z0
    flags = torch.tensor([int(all(v for k, v in out.items() if k.endswith("equal")))])
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    out["all_ranks_ok"] = bool(flags.item())
    if rank == 0:
        import json

        print("RESULT " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
