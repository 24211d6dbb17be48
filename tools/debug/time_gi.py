"""Times the passes of config 4 with set_option toggles: usage time_gi.py name=value[,name=value...] ..."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import bench, scene_util as su
from voxeltracing_b200 import engine
from voxeltracing_b200.pipeline import FrameRenderer

wl = bench.WORKLOADS["config4_1080p_gi"]
blocks, _ = bench.build_world(wl["world"])
inputs = su.SceneInputs(512, sky="constant")
ctx = engine.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
ctx.upload_world(blocks); ctx.generate_distance_field(); ctx.set_blue_noise_texture(bench.BLUE_TEX); inputs.apply_to_context(ctx)
fr = FrameRenderer(ctx, bench.frame_config(wl), inputs.grass, inputs.cactus)
N = 16
prepared = [fr.prepare(bench.camera_for(wl, s), s) for s in range(N)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for setting in (sys.argv[1:] or ["gi_fuse_final=1", "gi_fuse_final=0"]):
    for kv in setting.split(","):
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    ev = [{p: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for p in wl["passes"]} for _ in range(N)]
    for s in range(2):
        fr.submit(prepared[s])
    for s in range(N):
        flush.zero_()
        fr.submit(prepared[s], hook=lambda name, where, s=s: ev[s][name][0 if where == "begin" else 1].record(stream))
    torch.cuda.synchronize()
    ms = {p: float(np.mean([e[p][0].elapsed_time(e[p][1]) for e in ev])) for p in wl["passes"]}
    print(f"{setting:28s}", " ".join(f"{p}={v:.3f}" for p, v in ms.items()), "total=%.3f" % sum(ms.values()), flush=True)
ctx.close()
