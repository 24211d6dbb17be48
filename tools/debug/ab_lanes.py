"""A / B of the pass-level concurrency options on config 4: device-resident frame time (event pair per step, L2 flushed between steps, like
bench.py's timed region) and the end-to-end leg (per-pass asynchronous copies of the output attachments to page-locked host memory, two buffer
sets, wall clock over the steps).  usage: ab_lanes.py name=value[,name=value...] ...   (pass_overlap = 1, wf_bands = 2 underneath)"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import bench, scene_util as su
from voxeltracing_b200 import engine
from voxeltracing_b200.pipeline import FrameRenderer, PASS_OUTPUTS

wl = bench.WORKLOADS["config4_1080p_gi"]
blocks, _ = bench.build_world(wl["world"])
inputs = su.SceneInputs(512, sky="constant")
ctx = engine.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
ctx.upload_world(blocks); ctx.generate_distance_field(); ctx.set_blue_noise_texture(bench.BLUE_TEX); inputs.apply_to_context(ctx)
fr = FrameRenderer(ctx, bench.frame_config(wl), inputs.grass, inputs.cactus)
N, W0 = 30, 5
prepared = [fr.prepare(bench.camera_for(wl, s), s) for s in range(N + W0)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
FINAL = ("primary", "shadow", "gi", "reflection", "direct")
out_atts = [a for p in wl["passes"] if p in FINAL for a in PASS_OUTPUTS[p]]
fr.submit(prepared[0]); ctx.synchronize()
host = [{a: torch.empty(int(np.prod(ctx.attachment_info(a)[1:])), dtype=torch.uint8).pin_memory() for a in out_atts} for _ in range(2)]
nbytes = sum(t.numel() for t in host[0].values())
ctx.set_option("pass_overlap", 1); ctx.set_option("wf_bands", 2)
for setting in (sys.argv[1:] or ["lane2_direct=1,refl_defer_gi=1,copy_lanes=1", "lane2_direct=0,refl_defer_gi=0,copy_lanes=0"]):
    copies = True
    for kv in setting.split(","):
        k, v = kv.split("=")
        if k == "copies": copies = int(v) != 0
        else: ctx.set_option(k, int(v))
    res = []
    for rep in range(2):
        for s in range(W0):
            fr.submit(prepared[s])
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(N)]
        torch.cuda.synchronize()
        for s in range(N):
            flush.zero_()
            ev[s][0].record(stream)
            fr.submit(prepared[W0 + s])
            ev[s][1].record(stream)
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in ev]

        def e2e_step(s):
            prep = fr.prepare(bench.camera_for(wl, s), s)

            def hook(name, where):
                if copies and where == "end" and name in FINAL:
                    for a in PASS_OUTPUTS[name]:
                        ctx.copy_attachment_rect_async(a, host[s & 1][a].data_ptr(), (0, 0))
            fr.submit(prep, hook=hook)
        for s in range(3):
            e2e_step(s)
        ctx.wait_reads(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(N):
            e2e_step(W0 + s)
        t_host = time.perf_counter() - t0
        ctx.wait_reads(); torch.cuda.synchronize()
        e2e = (time.perf_counter() - t0) / N * 1e3
        res.append((float(np.mean(ms)), float(np.median(ms)), e2e, t_host / N * 1e3))
    print(f"{setting:52s}", " | ".join("dev mean %.4f med %.4f  e2e %.4f (host issue %.3f) ms" % r for r in res),
          "| d2h %.1f MB -> %.1f GB/s" % (nbytes / 1e6, nbytes / (min(r[2] for r in res) * 1e-3) / 1e9 if copies else 0.0), flush=True)
ctx.close()
