"""Renders a few frames of a bench workload with options set from the command line - the process to put under ncu.
usage: one_frame.py [workload] [frames] [option=value ...] [tile=row0,rows,col0,cols]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import bench, scene_util as su
from voxeltracing_b200 import engine
from voxeltracing_b200.pipeline import FrameRenderer

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "config4_1080p_gi"]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
blocks, _ = bench.build_world(wl["world"])
inputs = su.SceneInputs(512, sky="constant" if wl["camera"] == "rooms" else "gradient")
ctx = engine.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
ctx.upload_world(blocks); ctx.generate_distance_field(); ctx.set_blue_noise_texture(bench.BLUE_TEX); inputs.apply_to_context(ctx)
tile = (0, 0, 0, 0)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    if k == "tile":
        tile = tuple(int(x) for x in v.split(","))
    else:
        ctx.set_option(k, int(v, 0))
fr = FrameRenderer(ctx, bench.frame_config(wl), inputs.grass, inputs.cactus)
for s in range(frames):
    fr.submit(fr.prepare(bench.camera_for(wl, s + 3), s + 3, tile))
torch.cuda.synchronize()
ctx.close()
