import torch, time
d = torch.empty(128 << 20, dtype=torch.uint8, device="cuda"); h = torch.empty(128 << 20, dtype=torch.uint8).pin_memory()
for direction in ("d2h", "h2d"):
    for _ in range(3):
        (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True)); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(direction, "%.1f GB/s" % (20 * (128 << 20) / dt / 1e9))
