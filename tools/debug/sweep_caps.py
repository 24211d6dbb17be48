"""Times the passes of a workload for several iteration-cap schedules of the queue trace kernels (set_option "trace_caps";
caps as bytes, low byte first, 0 = the uncapped round-1 kernels) or, with VXRT_SWEEP_OPTION=trace_spill, for several hand-over thresholds
of the adaptive variant.  usage: sweep_caps.py [workload] [caps,caps,...]"""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import bench, scene_util as su
from voxeltracing_b200 import abi, engine
from voxeltracing_b200.pipeline import FrameRenderer

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "config4_1080p_gi"]
sched = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "12-24", "8-20", "12", "6-12-24", "10-20-32", "16-32"]
blocks, _ = bench.build_world(wl["world"])
inputs = su.SceneInputs(512, sky="constant" if wl["camera"] == "rooms" else "gradient")
ctx = engine.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
ctx.upload_world(blocks); ctx.generate_distance_field(); ctx.set_blue_noise_texture(bench.BLUE_TEX); inputs.apply_to_context(ctx)
fr = FrameRenderer(ctx, bench.frame_config(wl), inputs.grass, inputs.cactus)
N = 12
prepared = [fr.prepare(bench.camera_for(wl, s), s) for s in range(N)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for sc in sched:
    caps = [int(v) for v in sc.split("-")]
    ctx.set_option(os.environ.get("VXRT_SWEEP_OPTION", "trace_caps"), sum(k << (8 * j) for j, k in enumerate(caps)))
    ev = [{p: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for p in wl["passes"]} for _ in range(N)]
    for s in range(2):
        fr.submit(prepared[s])
    for s in range(N):
        flush.zero_()
        fr.submit(prepared[s], hook=lambda name, where, s=s: ev[s][name][0 if where == "begin" else 1].record(stream))
    torch.cuda.synchronize()
    ms = {p: float(np.mean([e[p][0].elapsed_time(e[p][1]) for e in ev])) for p in wl["passes"]}
    out = [ctx.read_attachment(a).copy() for a in (abi.ATT_GI_SH, abi.ATT_GI_AOSKY, abi.ATT_REFL_COLOR, abi.ATT_REFL_HITDIST) if a in fr.outputs]
    same = "-" if ref is None else all(np.array_equal(a.view(np.uint8), b.view(np.uint8)) for a, b in zip(out, ref))
    ref = ref or out
    print(f"caps={sc:10s}", " ".join(f"{p}={v:.3f}" for p, v in ms.items()), "total=%.3f" % sum(ms.values()), "identical to first:", same, flush=True)
ctx.close()
