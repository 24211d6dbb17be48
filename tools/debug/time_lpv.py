"""Times vxrt_cuda_lpv_repropagate per distance limit for both implementations (CUDA events, L2 flushed) and reports the launch counts."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from voxeltracing_b200 import engine, host_api
c = engine.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); c.set_stream(stream.cuda_stream)
blocks = host_api.gen_world("rooms", 2)
rng = np.random.default_rng(4)
nz, ny, nx = blocks.shape
blocks[rng.integers(1, nz, 2000), rng.integers(1, ny, 2000), rng.integers(1, nx, 2000)] = 12
t = np.full((6, 128), -1, dtype=np.int32); t[3, 12] = 0
c.set_block_data(t); c.upload_world(blocks)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for coop in (1, 0):
    c.set_option("lpv_coop", coop)
    for limit in (0, 3, 4, 6, 8):
        l0 = c.launch_count; c.lpv_repropagate(None, limit); dl = c.launch_count - l0
        for do_flush in (True, False):
            ts = []
            for _ in range(12):
                if do_flush: flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); c.lpv_repropagate(None, limit); b.record(stream); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            print(f"coop={coop} limit={limit} launches={dl} flush={do_flush}: median {np.median(ts[2:]):.1f} us min {min(ts[2:]):.1f} us")
c.close()
