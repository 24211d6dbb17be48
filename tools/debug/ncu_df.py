"""Three regenerations for an ncu capture (the kernels of the last one are the ones to read): usage ncu_df.py XYVER ZVER"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import torch
from voxeltracing_b200 import engine, host_api
ctx = engine.Context(0)
ctx.set_option("df_xyver", int(sys.argv[1])); ctx.set_option("df_zver", int(sys.argv[2]))
ctx.upload_world(host_api.gen_world("plains", 0))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_(); torch.cuda.synchronize()
    ctx.generate_distance_field(); ctx.synchronize()
ctx.close()
