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
for stage, do_flush, sx, sy in ((0, True, 0, 0), (1, True, 0, 0), (2, True, 0, 0), (1, True, 1, 1), (1, True, 3, 1), (1, True, 1, 4), (1, True, 3, 2), (0, True, 1, 1)):
    ctx.set_option("df_stage", stage); ctx.set_option("df_sx", sx); ctx.set_option("df_sy", sy)
    ts = []
    for _ in range(40):
        if do_flush: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); ctx.generate_distance_field(); b.record(stream); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = np.array(ts[5:])
    print(f"stage={stage} sx={sx} sy={sy} flush={do_flush}: median {np.median(ts):.1f} us  min {ts.min():.1f} us  -> {2*blocks.size/np.median(ts)/1e3:.0f} GB/s algorithmic")
ctx.set_option("df_stage", 0); ctx.set_option("df_sx", 0); ctx.set_option("df_sy", 0)
ctx.generate_distance_field()
print("bit-exact again:", np.array_equal(ctx.download_distance_field(), ob.distance_field(blocks)))
ctx.close()
