"""Worker of the 2-GPU NCCL test: z-slab sharded distance-field regeneration and screen-tile band gather
through the C ABI, one process per GPU."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import binding as ob  # noqa: E402
from voxeltracing_b200 import abi, engine, host_api, sharding  # noqa: E402
from voxeltracing_b200.pipeline import band_rows  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    blocks = host_api.gen_world("town", 3)
    ctx = engine.Context(local)
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.upload_world(blocks)
    backend = sharding.CudaSlabBackend(ctx, f"cuda:{local}")
    sharding.regenerate_distance_field_sharded(backend, rank, world)
    torch.cuda.synchronize()
    want = ob.distance_field(blocks)
    ok = np.array_equal(ctx.download_distance_field(), want)

    W, H = 640, 360
    cam = host_api.camera([192, 80, 192], 45.0, -20.0, W / H)
    row0, rows = band_rows(H, rank, world)
    ctx.initial_trace(cam, W, H, tile=(row0, rows))
    t = torch.as_tensor(ctx.attachment_as_device_array(abi.ATT_INITIAL_T), device=f"cuda:{local}")
    b = torch.as_tensor(ctx.attachment_as_device_array(abi.ATT_INITIAL_BLOCK), device=f"cuda:{local}")
    sharding.gather_bands(t.view(torch.int16), H, rank, world)
    sharding.gather_bands(b, H, rank, world)
    torch.cuda.synchronize()
    if rank == 0:
        ow = ob.OracleWorld(blocks, want)
        p = abi.PrimaryParams()
        for i in range(16):
            p.inv_view[i] = float(cam.inv_view[i]); p.inv_projection[i] = float(cam.inv_projection[i])
        p.width, p.height, p.render_distance = W, H, 350
        whole = ow.initial_trace(p)
        ok = ok and np.array_equal(ctx.read_attachment(abi.ATT_INITIAL_T).view(np.uint16), whole["t"].view(np.uint16))
        ok = ok and np.array_equal(ctx.read_attachment(abi.ATT_INITIAL_BLOCK), whole["block"])
    flag = torch.tensor([int(ok)], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("RESULT ok" if flag.item() == 1 else "RESULT mismatch")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
