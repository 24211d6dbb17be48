import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from oracle import binding as ob
from voxeltracing_b200 import engine, host_api
from test_gpu_trace import _ray_batch
blocks = host_api.gen_world("plains", 0)
ctx = engine.Context(0); ctx.upload_world(blocks); ctx.generate_distance_field()
ow = ob.OracleWorld(blocks)
for mi in (350, 48, 1, 0):
    o, d = _ray_batch(400_000, 5 + mi)
    got = ctx.trace_rays(o, d, mi); want = ow.traverse_batch(o, d, mi)
    def B(a):
        a = np.ascontiguousarray(a); b = a.view(np.uint32).copy(); b[np.isnan(a)] = 0x7FC00000; return b
    print("rays with NaN end:", int(np.isnan(want["end"]).any(1).sum()), "rays reaching the tail (NaN coordinate inside):",
          int((np.isnan(want["end"]).any(1) & (want["iterations"] > 1)).sum()))
    bad = np.nonzero((B(got["t"]) != B(want["t"])) | (B(got["end"]) != B(want["end"])).any(1)
                     | (got["iterations"] != want["iterations"]) | (got["intersection"] != want["intersection"]) | (got["block"] != want["block"])
                     | (B(got["normal"]) != B(want["normal"])).any(1))[0]
    print("max_iter", mi, "mismatches", len(bad), "of", len(o))
    for i in bad[:12]:
        print(i, "o", o[i], "d", d[i], "\n   got ", got[i], "\n   want", want[i][["t", "normal", "end", "block", "intersection", "iterations"]])
