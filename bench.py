#!/usr/bin/env python
"""bench.py — Mrays/s of DF-DDA traversal (BASELINE.json metric) on the frame workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One step = one frame of the workload (every VoxelTraversalDF invocation counts as one ray).  With
N > 1 (torchrun, one rank per GPU) every rank renders its own frame of the camera-path batch per step
(grids replicated, weak scaling) and the frames' output attachments are gathered to rank 0 over NCCL.
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # BASELINE.json configs[2]: 1080p primary + sun-shadow rays (+ direct shading) on generated plains
    "config3_1080p_primary_shadow": dict(width=1920, height=1080, world=("plains_structures", "plains", 1),
                                         passes=("primary", "shadow")),
}
DEFAULT_WORKLOAD = "config3_1080p_primary_shadow"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_world(spec, dims=(384, 128, 384)):
    from voxeltracing_b200 import host_api

    name, kind, seed = spec
    blocks, desc = host_api.load_named_world(name, kind, seed, dims)
    return blocks, desc


# --------------------------------------------------------------------------------------------------
# reference arm: the CPU implementation of the path on the host cores (oracle/_ref if it was built
# from the reference's shaders, else the oracle port)
# --------------------------------------------------------------------------------------------------

def cpu_frame_runner(blocks, wl):
    """Returns (run(frame_index) -> rays, kind, description). Prefers oracle/_ref (the reference's own
    shader sources compiled for the CPU)."""
    from oracle import binding as ob
    from voxeltracing_b200 import abi, host_api
    from voxeltracing_b200.pipeline import orbit_camera

    kind = "port"
    ow = ob.OracleWorld(blocks)
    W, H = wl["width"], wl["height"]
    light = host_api.sun_direction(50.0)[2]
    rng = np.random.default_rng(11)
    blue = rng.integers(0, 256, (256, 256, 4), dtype=np.uint8)

    def fill(dst, src):
        for i, v in enumerate(np.asarray(src, np.float32).ravel()):
            dst[i] = float(v)

    def run(frame: int, rows=None) -> int:
        cam = orbit_camera(frame, W / H)
        rays = 0
        p = abi.PrimaryParams()
        fill(p.inv_view, cam.inv_view); fill(p.inv_projection, cam.inv_projection)
        p.width, p.height, p.render_distance = W, H, 350
        if rows:
            p.tile.row0, p.tile.rows = rows
        g = ow.initial_trace(p, want_stats=True)
        rays += g["stats"]["rays"]
        if "shadow" in wl["passes"]:
            s = abi.ShadowParams()
            fill(s.inv_view, cam.inv_view); fill(s.inv_projection, cam.inv_projection)
            s.width, s.height = W, H
            fill(s.light_direction, light)
            s.current_frame, s.soft_shadows, s.max_iterations = frame, 1, 350
            if rows:
                s.tile.row0, s.tile.rows = rows
            o = ow.shadow_trace(s, g["t"], g["normal"], blue, want_stats=True)
            rays += o["stats"]["rays"]
        return rays

    return run, kind, ob.get_threads()


def run_reference_arm(args, wl, rank, world_size):
    if rank != 0:
        return
    blocks, world_desc = build_world(wl["world"])
    run, kind, cores = cpu_frame_runner(blocks, wl)
    H = wl["height"]
    # bounded sample: a band of rows of each frame, sized so a step is ~0.25 s of CPU work
    t0 = time.perf_counter(); probe_rays = run(0, rows=(H // 2 - 8, 16)); dt = time.perf_counter() - t0
    rows = int(min(H, max(16, 16 * 0.25 / max(dt, 1e-4))))
    band = ((H - rows) // 2, rows)
    for i in range(args.warmup):
        run(i, rows=band)
    rays = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        rays += run(i, rows=band)
    dt = time.perf_counter() - t0
    mrays = rays / dt / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s DF-DDA traversal", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "world": world_desc, "width": wl["width"], "height": wl["height"]},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": kind,
                         "sample": f"rows [{band[0]},{band[0] + band[1]}) of each {wl['width']}x{wl['height']} frame, {args.steps} frames"},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------

def main():
    args = parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, wl, rank, world_size)
        return

    import torch
    import torch.distributed as dist

    from voxeltracing_b200 import abi, engine
    from voxeltracing_b200.pipeline import FrameConfig, FrameRenderer, orbit_camera

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W, H = wl["width"], wl["height"]
    blocks, world_desc = build_world(wl["world"])
    ctx = engine.Context(local_rank)
    # all work of this rank (our kernels, NCCL gathers, timing events) goes on one explicit torch stream
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.upload_world(blocks)
    ctx.generate_distance_field()
    rng = np.random.default_rng(11)
    ctx.set_blue_noise_texture(rng.integers(0, 256, (256, 256, 4), dtype=np.uint8))
    cfg = FrameConfig(width=W, height=H, passes=wl["passes"])
    fr = FrameRenderer(ctx, cfg)

    # every rank renders its own frame of the camera path per step (weak scaling)
    def frame_of(step):
        return step * world_size + rank

    cams = [orbit_camera(frame_of(s), W / H) for s in range(args.warmup + args.steps + 1)]

    # rays per step are deterministic: count them once with the stats variant, outside the timed region
    ctx.stats_enable(True)
    ctx.stats_read(reset=True)
    rays_per_step = []
    iters_total = 0
    for s in range(args.warmup, args.warmup + args.steps):
        fr.render(cams[s], frame=frame_of(s))
        st = ctx.stats_read(reset=True)
        rays_per_step.append(st["rays"])
        iters_total += st["iterations"]
    ctx.stats_enable(False)
    total_rays = int(sum(rays_per_step))
    mean_iters = iters_total / max(total_rays, 1)
    # primary-only statistics for the roofline of the dominant kernel
    ctx.stats_enable(True); ctx.stats_read(reset=True)
    prim = FrameRenderer(ctx, FrameConfig(width=W, height=H, passes=("primary",)))
    for s in range(args.warmup, args.warmup + args.steps):
        prim.render(cams[s], frame=frame_of(s))
    pst = ctx.stats_read(reset=True)
    ctx.stats_enable(False)
    fr.render(cams[0])  # restore full attachments

    # gather buffers (N > 1): output attachments of every rank land on rank 0
    outs = [torch.as_tensor(ctx.attachment_as_device_array(a), device=f"cuda:{local_rank}") for a in cfg.outputs]
    gather_bufs = None
    if world_size > 1:
        gather_bufs = [[torch.empty_like(o) for _ in range(world_size)] if rank == 0 else None for o in outs]

    # the inputs (37.7 MB of grids) are smaller than L2, so L2 is flushed between timed steps by
    # overwriting a 256 MiB buffer; the flush is outside the per-step event pairs.
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")

    def step(s, hook=None):
        fr.render(cams[s], frame=frame_of(s), hook=hook)
        if world_size > 1:
            for o, gb in zip(outs, gather_bufs):
                dist.gather(o, gb, dst=0)

    for s in range(args.warmup):
        step(s)
    torch.cuda.synchronize()

    # ---- timed region: device-resident ----
    dom = "primary"
    dom_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]

    def make_hook(i):
        def hook(name, where):
            if name == dom:
                dom_ev[i][0 if where == "begin" else 1].record(stream)
        return hook

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world_size > 1:
        dist.barrier()
    torch.cuda.synchronize()
    step_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches_before = ctx.launch_count
    for i in range(args.steps):
        flush_buf.zero_()
        step_ev[i][0].record(stream)
        step(args.warmup + i, make_hook(i))
        step_ev[i][1].record(stream)
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
    ms = float(sum(a.elapsed_time(b) for a, b in step_ev))  # exactly K steps, flushes excluded
    launches = ctx.launch_count - launches_before
    clocks = sampler.stop() if rank == 0 else None
    dom_ms = float(np.mean([a.elapsed_time(b) for a, b in dom_ev]))

    # ---- distance-field regeneration (BASELINE config 2), L2 flushed before every regeneration ----
    df_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
    for a, b in df_ev:
        flush_buf.zero_()
        a.record(stream)
        ctx.generate_distance_field()
        b.record(stream)
    torch.cuda.synchronize()
    df_us = float(np.median([a.elapsed_time(b) for a, b in df_ev])) * 1e3

    # ---- end to end through the C ABI with host buffers (params in, attachments out) ----
    host_out = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
    host_np = [h.numpy() for h in host_out]
    d2h = sum(h.nbytes for h in host_np)
    h2d = 2 * (16 * 4 * 2 + 64)  # the per-pass parameter blocks (matrices + scalars) are the only per-step inputs

    def e2e_step(s):
        fr.render(cams[s], frame=frame_of(s))
        for att, buf in zip(cfg.outputs, host_np):
            ctx.read_attachment(att, buf)

    for s in range(min(3, args.warmup)):
        e2e_step(s)
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
    e2e_steps = max(3, min(args.steps, 50))
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(args.warmup + i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_rays = int(sum(rays_per_step[:e2e_steps]))

    # ---- reduce over ranks ----
    if world_size > 1:
        t = torch.tensor([ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
        r = torch.tensor([total_rays, e2e_rays, launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        total_rays, e2e_rays, launches = int(r[0]), int(r[1]), int(r[2])

    if rank == 0:
        peak, peak_src = measured_peaks()
        mrays = total_rays / (ms * 1e-3) / 1e6
        # roofline of the dominant kernel (primary trace): algorithmic bytes per ray = S + 1 + W
        S = pst["iterations"] / max(pst["rays"], 1)
        rays_per_launch = pst["rays"] / args.steps
        alg_bytes = rays_per_launch * (S + 1 + 8)
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        sector_bytes = rays_per_launch * (32 * (S + 1) + 8)
        line = {
            "metric": "Mrays/s DF-DDA traversal", "value": mrays, "unit": "Mrays/s", "n_gpus": world_size,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "world": world_desc, "width": W, "height": H, "passes": list(cfg.passes),
                       "rays_per_step_per_gpu": total_rays / args.steps / world_size, "mean_iterations_per_ray": mean_iters,
                       "sharding": "one frame of the camera path per rank per step; outputs gathered to rank 0 (NCCL)" if world_size > 1 else "single GPU",
                       "l2": "flushed between timed steps (256 MiB memset outside the per-step CUDA-event pairs); camera pose changes every step"},
            "clocks": clocks,
            "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": launches,
            "roofline": {"kernel": "initial_trace_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "mean_iterations_per_ray": S, "rays_per_launch": rays_per_launch, "avg_launch_ms": dom_ms,
                         "algorithmic_bytes_per_ray": S + 1 + 8, "l2_sector_gbs": sector_bytes / (dom_ms * 1e-3) / 1e9,
                         "kernel_mrays": rays_per_launch / (dom_ms * 1e-3) / 1e6},
        }
        nvox = blocks.size
        line["df_regen"] = {"us_per_regeneration": df_us, "algorithmic_bytes": 2 * nvox,
                            "achieved_gbs": 2 * nvox / (df_us * 1e-6) / 1e9, "frac_of_hbm_peak": 2 * nvox / (df_us * 1e-6) / 1e9 / peak,
                            "l2": "flushed before each regeneration", "launches_per_regeneration": 2}
        if not args.no_cpu_baseline and world_size == 1:
            run, kind, cores = cpu_frame_runner(blocks, wl)
            t0 = time.perf_counter(); run(0, rows=(H // 2 - 8, 16)); dt = time.perf_counter() - t0
            rows = int(min(H, max(16, 16 * 1.0 / max(dt, 1e-4))))
            band = ((H - rows) // 2, rows)
            n, rays, t0 = 0, 0, time.perf_counter()
            while time.perf_counter() - t0 < args.cpu_seconds and n < args.steps:
                rays += run(args.warmup + n, rows=band); n += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind,
                                    "sample": f"rows [{band[0]},{band[0] + band[1]}) of {n} frames of the same camera path"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))

    ctx.close()
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
