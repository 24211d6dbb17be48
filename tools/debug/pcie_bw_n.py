"""Aggregate device -> pinned-host copy bandwidth with every rank copying at once (the roof of the end-to-end leg at N GPUs).
usage: torchrun --nproc-per-node N tools/debug/pcie_bw_n.py"""
import os, time, json
import torch, torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 96 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8).pin_memory()
res = {}
for label, together in (("alone", False), ("all_ranks_at_once", True)):
    for r in range(world if not together else 1):
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
        if together or r == rank:
            for _ in range(3):
                h.copy_(d, non_blocking=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                h.copy_(d, non_blocking=True)
            torch.cuda.synchronize()
            res[label] = 20 * n / (time.perf_counter() - t0) / 1e9
        if world > 1:
            dist.barrier()
out = [None] * world
if world > 1:
    dist.all_gather_object(out, res)
else:
    out = [res]
if rank == 0:
    print(json.dumps({"n_gpus": world, "d2h_gbs_alone_per_rank": [round(o["alone"], 1) for o in out],
                      "d2h_gbs_all_at_once_per_rank": [round(o["all_ranks_at_once"], 1) for o in out],
                      "aggregate_all_at_once": round(sum(o["all_ranks_at_once"] for o in out), 1)}))
if world > 1:
    dist.destroy_process_group()
