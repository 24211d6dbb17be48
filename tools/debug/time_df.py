"""Times vxrt_cuda_generate_distance_field (median of 30, L2 flushed / not flushed) and checks it against the oracle."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from oracle import binding as ob
from voxeltracing_b200 import engine, host_api
blocks = host_api.gen_world("plains", 0)
ctx = engine.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
ctx.upload_world(blocks); ctx.generate_distance_field()
ok = np.array_equal(ctx.download_distance_field(), ob.distance_field(blocks))
print("bit-exact vs oracle:", ok)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for do_flush in (True, False):
    ts = []
    for _ in range(40):
        if do_flush: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); ctx.generate_distance_field(); b.record(stream); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = np.array(ts[5:])
    print(f"flush={do_flush}: median {np.median(ts):.1f} us  min {ts.min():.1f} us  -> {2*blocks.size/np.median(ts)/1e3:.0f} GB/s algorithmic")
ctx.close()
