#!/usr/bin/env python
"""bench.py — Mrays/s of DF-DDA traversal (BASELINE.json metric) on a frame workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One step = one frame of the workload: every VoxelTraversalDF invocation (primary, sun-shadow, GI bounce,
GI shadow, reflection, reflection shadow) counts as one ray.  With N > 1 (torchrun, one rank per GPU)
every rank renders its own frame of the camera-path batch per step (grids and tables replicated, weak
scaling) and the frames' output attachments are gathered to rank 0 over NCCL inside the timed region.
Prints ONE JSON line on rank 0.  DESIGN.md §6 defines every field.
"""
from __future__ import annotations

import argparse
import ctypes
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
sys.path.insert(0, str(ROOT / "tests"))

WORKLOADS = {
    # BASELINE.json configs[2]: 1080p primary + sun-shadow rays + Cook-Torrance direct shading, generated plains
    "config3_1080p_direct": dict(width=1920, height=1080, world=("plains_structures", "plains", 1), camera="orbit",
                                 passes=("primary", "gbuffer", "shadow", "direct")),
    # BASELINE.json configs[3]: 1080p 1-spp 2-bounce path-traced diffuse GI + rough reflections ('Test Worlds/gi')
    "config4_1080p_gi": dict(width=1920, height=1080, world=("gi", "rooms", 2), camera="rooms",
                             passes=("primary", "gbuffer", "gi", "shadow", "reflection", "direct")),
    # BASELINE.json configs[4]: 4K 4-spp GI camera-path batch on the Medival import (stand-in: town)
    "config5_4k_gi4": dict(width=3840, height=2160, world=("Medival", "town", 3), camera="orbit", gi_spp=4,
                           passes=("primary", "gbuffer", "gi", "shadow", "reflection", "direct")),
}
DEFAULT_WORKLOAD = "config4_1080p_gi"
TRACE_PASSES = ("primary", "shadow", "gi", "reflection")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pass-overlap", action="store_true", help="issue every pass on one stream (round 1 / early round 2 behaviour)")
    ap.add_argument("--wf-bands", type=int, default=2, help="row bands the GI / reflection wavefronts are pipelined in (set_option wf_bands)")
    ap.add_argument("--profile", action="store_true", help="only warm-up + steps of the plain step (for ncu)")
    ap.add_argument("--tex-size", type=int, default=512)
    ap.add_argument("--no-svgf", action="store_true", help="skip the SVGF denoiser-chain measurement")
    ap.add_argument("--sharding", default="auto", choices=["auto", "frames", "tiles"],
                    help="N > 1: 'frames' = one frame of the camera path per rank per step (weak scaling); 'tiles' = row strips of ONE frame "
                         "per step spread over the ranks (strong scaling; BASELINE config 5).  auto: tiles for config5, frames otherwise")
    ap.add_argument("--tile-shape", default="cols", choices=["cols", "rows"],
                    help="tiles sharding: 'cols' = column bands (every band holds sky and ground, so ONE tile per rank is balanced and every pass is "
                         "one launch per rank); 'rows' = row strips, interleaved over the frame when --strips > 1 (round 2's first version: 4 strips, "
                         "4 x the launches)")
    ap.add_argument("--strips", type=int, default=0, help="tiles sharding: tiles per rank, interleaved over the frame (default: 1 for cols, 4 for rows)")
    ap.add_argument("--gather", default="push", choices=["push", "nccl"],
                    help="N > 1: 'push' = every rank copies its outputs into rank 0's buffer with the copy engines over NVLink as each pass "
                         "finishes; 'nccl' = round 1's one NCCL gather per frame on a side stream")
    ap.add_argument("--outputs", default="final", choices=["final", "all"],
                    help="what is gathered to rank 0 / read back in the end-to-end leg: 'final' = G-buffer, shadow, GI, reflection and direct "
                         "attachments (44 B/px); 'all' adds the material G-buffer intermediates only later passes on the device consume (61 B/px)")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="vxrt_cuda_set_option(NAME, VALUE) on the bench's context after its own options (A / B runs: lane2_direct, refl_defer_gi, copy_lanes ...)")
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
        self.rows, self.proc, self.gpu = [], None, gpu_index

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


def bind_to_gpu_numa_node(local_rank: int):
    """Run this rank (and first-touch its pinned host buffers) on the CPUs its GPU hangs off: with 8 ranks reading their frames back
    over 8 PCIe links, remote-socket host memory is the bottleneck otherwise.  Best effort: the GPU's CPU affinity from NVML
    (what `nvidia-smi topo -m` prints), else the PCI device's numa_node in sysfs.  Returns a description or None."""
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(local_rank)
        bus_id = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode() if hasattr(bus_id, "encode") else bus_id)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * i + b for i, w_ in enumerate(words) for b in range(64) if (int(w_) >> b) & 1]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return f"nvml cpu affinity: {len(allowed)} cpus ({allowed[0]}-{allowed[-1]})"
        if allowed:
            return f"nvml cpu affinity covers every allowed cpu ({len(allowed)}): single NUMA domain"
    except Exception:
        pass
    try:
        import torch

        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return f"sysfs numa node {node}"
    except Exception:
        return None


def build_world(spec, dims=(384, 128, 384)):
    from voxeltracing_b200 import host_api

    name, kind, seed = spec
    return host_api.load_named_world(name, kind, seed, dims)


def camera_for(wl, frame):
    from voxeltracing_b200.pipeline import orbit_camera, rooms_camera

    aspect = wl["width"] / wl["height"]
    return rooms_camera(frame, aspect) if wl["camera"] == "rooms" else orbit_camera(frame, aspect)


def frame_config(wl):
    from voxeltracing_b200.pipeline import FrameConfig

    return FrameConfig(width=wl["width"], height=wl["height"], passes=wl["passes"], gi_spp=wl.get("gi_spp", 1))


BLUE_TEX = np.random.default_rng(11).integers(0, 256, (256, 256, 4), dtype=np.uint8)


# --------------------------------------------------------------------------------------------------
# CPU implementation of the path on the host cores: oracle/_ref (the reference's own shaders compiled
# for the CPU) when that build is present, else the oracle port
# --------------------------------------------------------------------------------------------------

class CpuFrame:
    def __init__(self, blocks, wl, inputs, prefer_ref=True):
        from oracle import binding as ob
        from oracle import ref_binding as rb
        from voxeltracing_b200.pipeline import FrameRenderer

        self.ob, self.rb, self.wl = ob, rb, wl
        self.ow = ob.OracleWorld(blocks)
        self.scene = ob.OracleScene(self.ow)
        inputs.apply_to_oracle(self.scene)
        need = {"primary": "initial", "shadow": "shadow", "gbuffer": "gbuffer", "gi": "diffuse", "reflection": "reflection", "direct": "color"}
        self.use_ref = (prefer_ref and blocks.shape == (384, 128, 384) and rb.available("df")
                        and all(rb.available(need[p]) for p in wl["passes"]))
        if self.use_ref:
            rb.set_scene(blocks, self.ow.df, inputs.table, inputs.blue, inputs.textures, inputs.sky)
        self.kind = "reference" if self.use_ref else "port"
        # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would leave the CPU
        # arm on one core.  The oracle and oracle/_ref share one OpenMP runtime, so one omp_set_num_threads covers both.
        try:
            n_threads = len(os.sched_getaffinity(0))
        except AttributeError:
            n_threads = os.cpu_count() or 1
        try:
            import ctypes
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n_threads))
        except OSError:
            pass
        ob.set_threads(int(n_threads))
        self.cores = ob.get_threads()
        # parameter marshalling is shared with the GPU arm (params_for needs no Context)
        self.fr = FrameRenderer(None, frame_config(wl), inputs.grass, inputs.cactus)

    def run(self, frame: int, rows=None) -> int:
        """Renders (a band of rows of) one frame on the CPU; returns the number of rays traced.  Every pass is
        restricted to the band: with equal resolutions and screen-space reprojection off (the bench
        configurations) a pass only reads its own pixel of the earlier attachments."""
        rb, ow, sc, wl = self.rb, self.ow, self.scene, self.wl
        cam = camera_for(wl, frame)
        tile = rows if rows else (0, 0)
        rays = 0
        out = {}
        for name in wl["passes"]:
            p = self.fr.params_for(name, cam, frame, tile)
            if name == "primary":
                if self.use_ref:
                    g = rb.initial_trace(ow.blocks, ow.df, p)
                    rays += (rows[1] if rows else wl["height"]) * wl["width"]  # one VoxelTraversalDF call per pixel
                else:
                    g = ow.initial_trace(p, want_stats=True)
                    rays += g["stats"]["rays"]
                out["g"] = g
            elif name == "gbuffer":
                g = out["g"]
                out["gb"] = (rb.generate_gbuffer if self.use_ref else sc.generate_gbuffer)(p, g["inv_t"], g["normal"], g["block"])
            elif name == "shadow":
                g = out["g"]
                if self.use_ref:
                    s = rb.shadow_trace(ow.blocks, ow.df, p, g["t"], g["normal"], BLUE_TEX)
                    # a shadow ray is cast unless the pixel is sky (transversal 64) or faces away (0.01)
                    r0, nr = rows if rows else (0, wl["height"])
                    tr = s["transversal"][r0:r0 + nr].astype(np.float32)
                    rays += int(((tr != 64.0) & (tr != np.float32(np.float16(0.01)))).sum())
                else:
                    s = ow.shadow_trace(p, g["t"], g["normal"], BLUE_TEX, want_stats=True)
                    rays += s["stats"]["rays"]
                out["sh"] = s
            elif name == "gi":
                g = out["g"]
                out["gi"] = sc.diffuse_trace(p, g["t"], g["normal"])
                rays += out["gi"]["stats"]["rays"]
            elif name == "reflection":
                g = out["g"]
                rays += sc.reflection_trace(p, g["t"], g["normal"], out["gb"], out["gi"], out["sh"]["shadow"])["stats"]["rays"]
            elif name == "direct":
                g = out["g"]
                (rb.shade_direct if self.use_ref else sc.shade_direct)(p, g["inv_t"], out["gb"], out["sh"]["shadow"])
        return rays


def cpu_sample(cpu: CpuFrame, wl, first_frame: int, max_steps: int, seconds: float, frame_budget_s: float):
    """Times a bounded sample: a band of rows of each frame, sized so one frame is ~frame_budget_s of CPU work."""
    H = wl["height"]
    probe_rows = 64
    t0 = time.perf_counter(); cpu.run(first_frame, rows=(H // 2 - probe_rows // 2, probe_rows)); dt = time.perf_counter() - t0
    rows = int(min(H, max(8, probe_rows * frame_budget_s / max(dt, 1e-4))))
    band = ((H - rows) // 2, rows)
    n, rays, t0 = 0, 0, time.perf_counter()
    while n < max_steps and (time.perf_counter() - t0 < seconds or n < 2):
        rays += cpu.run(first_frame + n, rows=band); n += 1
    dt = time.perf_counter() - t0
    return rays / dt / 1e6, dt / n * 1e3, n, band


def cpu_arm(blocks, wl, inputs):
    """oracle/_ref runs the reference's own shaders; its GI / reflection drivers do not count rays, so the
    frames that contain those passes are timed on the port (bit-identical to oracle/_ref: tests/test_oracle_golden.py)."""
    return CpuFrame(blocks, wl, inputs, prefer_ref=not any(p in wl["passes"] for p in ("gi", "reflection")))


def config_block(args, wl, world_desc):
    """The workload keys both arms print (identical, so the driver's same_config comparison holds)."""
    cfg = frame_config(wl)
    n = args.gpus
    tiles = n > 1 and (args.sharding == "tiles" or (args.sharding == "auto" and args.workload.startswith("config5")))
    sharding = "single GPU" if n == 1 else (f"tiles: {n} ranks render the tiles of ONE frame per step (strong scaling)" if tiles
                                            else f"frames: each of {n} ranks renders its own frame per step (weak scaling)")
    return {"workload": args.workload, "world": world_desc, "width": wl["width"], "height": wl["height"], "passes": list(wl["passes"]),
            "gi_spp": cfg.gi_spp, "reflection_spp": cfg.refl_spp, "texture_size": args.tex_size, "sharding": sharding,
            "l2": "GPU arm: L2 flushed between timed steps (256 MiB memset outside the per-step CUDA-event pairs), camera pose changes every step; "
                  "CPU arm: not applicable"}


CPU_FRAME_BUDGET_S = 1.0   # one band of rows per frame sized to about this much CPU work, in BOTH the --impl reference arm and the in-line cpu_baseline leg


def run_reference_arm(args, wl, rank):
    if rank != 0:
        return
    import scene_util as su

    blocks, world_desc = build_world(wl["world"])
    inputs = su.SceneInputs(args.tex_size, sky="constant" if wl["camera"] == "rooms" else "gradient")
    cpu = cpu_arm(blocks, wl, inputs)
    for i in range(min(args.warmup, 2)):
        cpu.run(i, rows=(wl["height"] // 2 - 32, 64))
    mrays, ms_step, n, band = cpu_sample(cpu, wl, args.warmup, args.steps, 120.0, CPU_FRAME_BUDGET_S)
    line = {
        "impl": "reference", "metric": "Mrays/s DF-DDA traversal", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": n, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(args, wl, world_desc),
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cpu.cores, "kind": cpu.kind,
                         "sample": f"rows [{band[0]},{band[0] + band[1]}) of each {wl['width']}x{wl['height']} frame (every pass restricted to the band), {n} frames",
                         "note": "the CPU implementation of the path on THIS box's host cores (one box, whatever --gpus says): at N GPUs the driver's ratio is N GPUs over one host"},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------

def world_producers_block(local_rank, stream, flush_buf, ev, peak):
    """World producers on their own context: the terrain generator (writes the whole grid: N algorithmic
    bytes), the scatter of a batch of chunk sections (reads 6,148 B per section, writes up to 4,096) and the light scan
    (reads N); CUDA events on the stream, L2 flushed before every launch.  The section batch is synthetic (seeded)."""
    from types import SimpleNamespace

    import torch

    from voxeltracing_b200 import engine
    c = engine.Context(local_rank)
    c.set_stream(stream.cuda_stream)
    nvox = c.nvox
    rng = np.random.default_rng(9)
    n_sec = 2544   # the section count of the engine's 'Test MC Worlds/Medival'
    ids = np.where(rng.random((n_sec, 4096)) < 0.002, 12, rng.integers(0, 2, (n_sec, 4096)) * 3).astype(np.uint8)   # air, stone, 0.2 % lamps
    org = np.stack([rng.integers(-12, 12, n_sec) * 16, rng.integers(0, 8, n_sec) * 16, rng.integers(-12, 12, n_sec) * 16], axis=1).astype(np.int32)
    sec = SimpleNamespace(block_ids=ids, data_nibbles=np.zeros((n_sec, 2048), np.uint8), has_data=np.ones(n_sec, np.uint8), origins=org)
    lut = np.arange(256, dtype=np.uint8)
    table = np.full((6, 128), -1, dtype=np.int32)
    table[3, 12] = 0
    c.set_block_data(table)

    def timed(fn, reps):
        pairs = []
        for _ in range(reps):
            flush_buf.zero_()
            a, b = ev(), ev()
            a.record(stream)
            fn()
            b.record(stream)
            pairs.append((a, b))
        torch.cuda.synchronize()
        return float(np.median([a.elapsed_time(b) for a, b in pairs])) * 1e3

    gen_us = timed(lambda: c.generate_world(1, 4242, 999), 20)
    t0 = time.perf_counter()
    c.import_sections(sec, (0, 0, 0), lut)     # host arrays -> staging -> scatter, synchronous like ImportWorld
    imp_ms = (time.perf_counter() - t0) * 1e3
    c.generate_world(1, 4242, 999)
    c.import_sections(sec, (0, 0, 0), lut, clear_first=False)
    lights = c.collect_lights()
    flush_buf.zero_()
    a, b = ev(), ev()
    a.record(stream)
    t0 = time.perf_counter()
    lights = c.collect_lights(capacity=len(lights))
    lights_ms = (time.perf_counter() - t0) * 1e3
    b.record(stream)
    torch.cuda.synchronize()
    lights_dev_us = a.elapsed_time(b) * 1e3
    c.close()
    return {"generate_world": {"us_per_world": gen_us, "algorithmic_bytes": nvox, "achieved_gbs": nvox / (gen_us * 1e-6) / 1e9,
                               "frac_of_hbm_peak": nvox / (gen_us * 1e-6) / 1e9 / peak, "launches": 1},
            "import_sections": {"ms_end_to_end": imp_ms, "sections": n_sec, "h2d_bytes": int(ids.nbytes + sec.data_nibbles.nbytes + org.nbytes + n_sec),
                                "note": "wall clock around the ABI call: pageable host arrays -> device staging -> scatter kernel, synchronous"},
            "collect_lights": {"ms_end_to_end": lights_ms, "us_on_device": lights_dev_us, "lights": int(len(lights)), "algorithmic_bytes": nvox,
                               "achieved_gbs": nvox / (lights_dev_us * 1e-6) / 1e9, "launches": 3,
                               "note": "count + scan + write kernels and the read-back of the list; the grid is read twice (second pass from L2)"},
            "l2": "flushed before each generate_world launch"}


def lpv_block(local_rank, stream, flush_buf, ev, peak, with_cpu):
    """Light propagation volume flood fill on its own context: the full repropagation of a world with lamps (start-up
    sequence / World::RepropogateLPV_: clear both volumes = 2 N algorithmic bytes, light scan of the grid = N, then the waves) and one
    block edit (a lamp placed, then broken); CUDA events on the stream, L2 flushed before every repropagation.  The CPU figure is the
    oracle restatement of Core/VolumetricFloodFill.cpp (one thread, like the reference) on the same world."""
    import torch

    from voxeltracing_b200 import engine, host_api
    c = engine.Context(local_rank)
    c.set_stream(stream.cuda_stream)
    nvox = c.nvox
    blocks = host_api.gen_world("rooms", 2)
    rng = np.random.default_rng(4)
    nz, ny, nx = blocks.shape
    n_lamps = 2000
    blocks[rng.integers(1, nz, n_lamps), rng.integers(1, ny, n_lamps), rng.integers(1, nx, n_lamps)] = 12
    table = np.full((6, 128), -1, dtype=np.int32)
    table[3, 12] = 0
    c.set_block_data(table)
    c.upload_world(blocks)
    out = {"world": f"stand-in:rooms(seed=2) + {n_lamps} lamps", "l2": "flushed before each repropagation"}
    for limit in (4, 8):
        c.lpv_repropagate(None, limit)
        pairs = []
        for _ in range(10):
            flush_buf.zero_()
            a, b = ev(), ev()
            a.record(stream)
            c.lpv_repropagate(None, limit)
            b.record(stream)
            pairs.append((a, b))
        torch.cuda.synchronize()
        us = float(np.median([a.elapsed_time(b) for a, b in pairs])) * 1e3
        lit = int(np.count_nonzero(c.lpv_download()[0]))
        alg = 3 * nvox + 4 * lit
        out[f"repropagate_limit{limit}"] = {"us": us, "lit_voxels": lit, "launches": 1, "algorithmic_bytes": alg,
                                            "achieved_gbs": alg / (us * 1e-6) / 1e9, "frac_of_hbm_peak": alg / (us * 1e-6) / 1e9 / peak}
        c.set_option("lpv_coop", 0)   # the same work as one kernel per phase (memset, 3 scan kernels, seed, 4 per level)
        c.lpv_repropagate(None, limit)
        a, b = ev(), ev()
        flush_buf.zero_()
        a.record(stream); c.lpv_repropagate(None, limit); b.record(stream)
        torch.cuda.synchronize()
        out[f"repropagate_limit{limit}"]["us_kernel_per_phase"] = a.elapsed_time(b) * 1e3
        c.set_option("lpv_coop", 1)
    lamp = np.argwhere(blocks == 12)[len(np.argwhere(blocks == 12)) // 2]
    z, y, x = (int(v) for v in lamp)
    times, dev = [], []
    for _ in range(5):   # break the lamp, place it again (limit 8: the largest removal)
        for op, blk in ((0, 0), (1, 12)):
            c.edit_blocks(np.array([[x, y, z, blk]], dtype=np.int32))
            torch.cuda.synchronize()
            a, b = ev(), ev()
            t0 = time.perf_counter()
            a.record(stream)
            c.lpv_edit(op, (x, y, z), 12, 8)
            b.record(stream)
            times.append((time.perf_counter() - t0) * 1e3)
            torch.cuda.synchronize()
            dev.append(a.elapsed_time(b) * 1e3)
    out["edit_limit8"] = {"ms_break_lamp": float(np.median(times[0::2])), "ms_place_lamp": float(np.median(times[1::2])),
                          "us_on_device_break_lamp": float(np.median(dev[0::2])), "us_on_device_place_lamp": float(np.median(dev[1::2])),
                          "note": "ms: wall clock around the synchronous ABI call (launch + stream synchronisation from an idle GPU); us_on_device: CUDA "
                                  "events around the same call; one warp replays the reference's queues in order"}
    # SampleLPVData on 2 M points around the lit voxels (host buffers in, host buffers out: the call copies both ways)
    c.lpv_repropagate(None, 8)
    level = c.lpv_download()[0]
    lit = np.argwhere(level > 0)
    n_pts = 1 << 21
    pts = (lit[rng.integers(0, len(lit), n_pts)][:, ::-1] + rng.random((n_pts, 3)) * 1.5 - 0.25).astype(np.float32)
    avg = np.zeros((128, 4), np.float32)
    avg[:, :3] = rng.random((128, 3)).astype(np.float32)
    c.lpv_set_average_colors(avg)
    c.lpv_sample(pts[:1024])
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        c.lpv_sample(pts, (0.5 / 384, 0.5 / 128, 0.5 / 384))
        ts.append(time.perf_counter() - t0)
    out["sample"] = {"points": n_pts, "ms_end_to_end": float(np.median(ts)) * 1e3, "mpoints_per_s_end_to_end": n_pts / float(np.median(ts)) / 1e6,
                     "h2d_bytes": int(pts.nbytes), "d2h_bytes": int(pts.nbytes),
                     "note": "wall clock around vxrt_cuda_lpv_sample with pageable host arrays (copy in, kernel, copy out, synchronous)"}
    if with_cpu:
        from oracle import world_binding as wb
        lights = wb.collect_lights(blocks, table)
        for limit in (4, 8):
            t0 = time.perf_counter()
            wb.lpv_repropagate(blocks, lights, limit)
            out[f"repropagate_limit{limit}"]["cpu_port_ms"] = (time.perf_counter() - t0) * 1e3
    c.close()
    return out


def parity_check(ctx, fr, wl, blocks, inputs, frame, read_full, rows=32):
    """Post-timing self-check (outside every timed region): a band of `rows` rows of one frame of the benched workload, rendered by
    the CUDA path at the benched resolution / texture size / spp, against the CPU oracle on the same inputs.  Tolerances are the ones
    of tests/test_gpu_shade*.py: integer / texture-fetch attachments bit exact; soft shadows, GI hit masks >= 99.9 % identical; R16F
    radiance within 1e-2 relative (+1e-3) on >= 99.5 % of pixels; direct term within 2^-9 relative on >= 99.9 %.
    read_full(att) returns the full-frame attachment (from this GPU, or from rank 0's gathered frame in tiles mode).
    The oracle is test infrastructure: it is only the checker here."""
    from oracle import binding as ob
    from voxeltracing_b200 import abi

    W, H = wl["width"], wl["height"]
    band = ((H - rows) // 2 // 8 * 8, rows)
    sl = slice(band[0], band[0] + band[1])
    cam = camera_for(wl, frame)
    ow = ob.OracleWorld(blocks)
    sc = ob.OracleScene(ow)
    inputs.apply_to_oracle(sc)
    res, ok = {}, True

    def close(got, want, rtol, atol):
        a, b = got.astype(np.float32), want.astype(np.float32)
        c = np.abs(a - b) <= rtol * np.abs(b) + atol
        return c.all(axis=-1) if c.ndim == 3 else c

    def exact(name, got, want):
        nonlocal ok
        if got is None:
            res[name] = "not among the gathered outputs"
            return
        same = bool(np.array_equal(np.ascontiguousarray(got[sl]).view(np.uint8), np.ascontiguousarray(want[sl]).view(np.uint8)))
        res[name] = "bit-exact" if same else "DIFFERS"
        ok = ok and same

    def frac(name, mask, need):
        nonlocal ok
        f = float(mask.mean())
        res[name] = round(f, 5)
        ok = ok and f >= need

    passes = wl["passes"]
    g = ow.initial_trace(fr.params_for("primary", cam, frame, band))
    for att, k in ((abi.ATT_INITIAL_T, "t"), (abi.ATT_INITIAL_NORMAL, "normal"), (abi.ATT_INITIAL_BLOCK, "block"), (abi.ATT_INITIAL_INVT, "inv_t")):
        exact("primary." + k, read_full(att), g[k])
    sh = None
    if "shadow" in passes:
        sh = ow.shadow_trace(fr.params_for("shadow", cam, frame, band), g["t"], g["normal"], BLUE_TEX)
        frac("shadow.identical", read_full(abi.ATT_SHADOW)[sl] == sh["shadow"][sl], 0.999)
    gb = None
    if "gbuffer" in passes:
        gb = sc.generate_gbuffer(fr.params_for("gbuffer", cam, frame, band), g["inv_t"], g["normal"], g["block"])
        for att, k in ((abi.ATT_GBUF_ALBEDO, "albedo"), (abi.ATT_GBUF_NORMAL, "normal"), (abi.ATT_GBUF_PBR, "pbr"), (abi.ATT_GBUF_TEXAO, "texao")):
            exact("gbuffer." + k, read_full(att), gb[k])
    gi_cuda = None
    if "gi" in passes:
        want = sc.diffuse_trace(fr.params_for("gi", cam, frame, band), g["t"], g["normal"])
        gi_cuda = {"sh": read_full(abi.ATT_GI_SH), "cocg": read_full(abi.ATT_GI_COCG), "utility": read_full(abi.ATT_GI_UTILITY),
                   "aosky": read_full(abi.ATT_GI_AOSKY)}
        frac("gi.paths_identical", (gi_cuda["aosky"][sl] == want["aosky"][sl]).all(axis=-1), 0.999)
        for k in ("sh", "cocg", "utility"):
            frac("gi." + k + "_within_1e-2", close(gi_cuda[k][sl], want[k][sl], 1e-2, 1e-3), 0.995)
    if "reflection" in passes and gb is not None and gi_cuda is not None and sh is not None:
        # the oracle's reflection pass reads the CUDA path's GI and shadow attachments, so only this pass is under test
        want = sc.reflection_trace(fr.params_for("reflection", cam, frame, band), g["t"], g["normal"], gb,
                                   {k: np.ascontiguousarray(v) for k, v in gi_cuda.items()}, np.ascontiguousarray(read_full(abi.ATT_SHADOW)))
        got_c, got_h, got_e = read_full(abi.ATT_REFL_COLOR), read_full(abi.ATT_REFL_HITDIST), read_full(abi.ATT_REFL_EMISSIVE)
        frac("reflection.mask_identical", got_e[sl] == want["emissive"][sl], 0.999)
        frac("reflection.hit_identical", (got_h[sl].astype(np.float32) > 0) == (want["hitdist"][sl].astype(np.float32) > 0), 0.999)
        frac("reflection.color_within_1e-2", close(got_c[sl], want["color"][sl], 1e-2, 1e-3), 0.995)
    if "direct" in passes and gb is not None and sh is not None:
        want = sc.shade_direct(fr.params_for("direct", cam, frame, band), g["inv_t"], gb, np.ascontiguousarray(read_full(abi.ATT_SHADOW)))
        frac("direct.within_2^-9", close(read_full(abi.ATT_DIRECT)[sl], want[sl], 2.0 ** -9, 1e-6), 0.999)
    return ("ok" if ok else "FAILED"), {"frame": frame, "rows": [band[0], band[0] + band[1]], "checks": res}


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else this process (or NCCL, which prints its version
    banner to stdout) writes to fd 1 has been redirected to stderr by quiet_stdout()."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = None


def quiet_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    args = parse_args()
    quiet_stdout()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, wl, rank)
        return

    import torch
    import torch.distributed as dist

    import scene_util as su
    from voxeltracing_b200 import engine
    from voxeltracing_b200.pipeline import PASS_KERNEL, PASS_OUTPUT_BYTES, PASS_OUTPUTS, FrameRenderer, band_rows

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W, H = wl["width"], wl["height"]
    blocks, world_desc = build_world(wl["world"])
    inputs = su.SceneInputs(args.tex_size, sky="constant" if wl["camera"] == "rooms" else "gradient")
    ctx = engine.Context(local_rank)
    # all work of this rank (our kernels, NCCL gathers, timing events) goes on one explicit torch stream
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.upload_world(blocks)
    ctx.generate_distance_field()
    ctx.set_blue_noise_texture(BLUE_TEX)
    inputs.apply_to_context(ctx)
    cfg = frame_config(wl)
    fr = FrameRenderer(ctx, cfg, inputs.grass, inputs.cactus)

    # ---- work decomposition over the ranks ----
    # frames: every rank renders its own frame of the camera path per step (weak scaling)
    # tiles:  every rank renders its tile(s) of THE SAME frame per step (strong scaling, BASELINE config 5), passed as the vxrt_tile of every
    #         pass: by default ONE column band per rank (sky and ground in every band; one launch per pass), or --tile-shape rows --strips k
    #         = k row strips per rank interleaved over the frame (k x the launches: 4 strips measured 3.7 ms per 4K frame at N = 8)
    tiles = world_size > 1 and (args.sharding == "tiles" or (args.sharding == "auto" and args.workload.startswith("config5")))
    n_strips = (args.strips if args.strips > 0 else (1 if args.tile_shape == "cols" else 4)) if tiles else 1

    def frame_of(step):
        return step if tiles else step * world_size + rank

    def strips_of(r):
        """tiles (row0, rows, col0, cols) of rank r; zeros = every row / column"""
        out = []
        for j in range(n_strips):
            if args.tile_shape == "cols":
                col0, cols = band_rows(W, r + j * world_size, world_size * n_strips, band=32)   # edges on the 32-pixel CTA grid
                if cols > 0:
                    out.append((0, 0, col0, cols))
            else:
                row0, rows = band_rows(H, r + j * world_size, world_size * n_strips)
                if rows > 0:
                    out.append((row0, rows, 0, 0))
        return out

    def rect_of(tile, ah, aw):
        row0, rows, col0, cols = tile
        return (row0, rows if rows else ah, col0, cols if cols else aw)

    my_tiles = strips_of(rank) if tiles else [(0, 0, 0, 0)]
    n_total = args.warmup + args.steps
    prepared = [[fr.prepare(camera_for(wl, frame_of(s)), frame_of(s), t) for t in my_tiles] for s in range(n_total)]

    def submit(s, hook_for=None):
        for k, prep in enumerate(prepared[s]):
            fr.submit(prep, hook=hook_for(k) if hook_for else None)

    if args.profile:
        for s in range(n_total):
            submit(s)
        torch.cuda.synchronize()
        ctx.close()
        return

    # rays per step are deterministic: count them once per pass with the stats variant of the kernels,
    # outside the timed region
    trace_passes = [p for p in cfg.passes if p in TRACE_PASSES]
    pass_stats = {p: {"rays": 0, "iterations": 0} for p in trace_passes}
    rays_per_step = []
    ctx.set_option("probe", 1)          # event pairs around every launch of the GI path-ray trace kernel
    ctx.stats_enable(True)
    ctx.stats_read(reset=True)
    ctx.probe_read(reset=True)
    for s in range(args.warmup, n_total):
        acc = [0]

        def stat_hook(name, where, acc=acc):
            if where == "end" and name in pass_stats:
                st = ctx.stats_read(reset=True)
                pass_stats[name]["rays"] += st["rays"]
                pass_stats[name]["iterations"] += st["iterations"]
                acc[0] += st["rays"]

        submit(s, lambda k: stat_hook)
        rays_per_step.append(acc[0])
    probe_stats = ctx.probe_read(reset=True)   # rays / iterations of the probed kernel's launches alone
    ctx.stats_enable(False)
    ctx.set_option("probe", 0)
    # pass-level concurrency (vxrt_cuda_set_option "pass_overlap"): shadow / reflection / direct passes on the context's second stream beside the GI
    # wavefront; FrameRenderer.submit joins the lanes at the end of the frame, so the step events below bracket all of it
    ctx.set_option("pass_overlap", 0 if args.no_pass_overlap else 1)
    ctx.set_option("wf_bands", args.wf_bands)
    for kv in args.opt:
        k_, v_ = kv.split("=", 1)
        ctx.set_option(k_, int(v_, 0))
    total_rays = int(sum(rays_per_step))
    total_iters = int(sum(v["iterations"] for v in pass_stats.values()))

    # ---- what leaves the GPU ----
    # 'final': the attachments later stages outside this path consume (primary G-buffer, shadow, GI, reflection, direct term: 44 B/px);
    # the material G-buffer of GenerateGBuffer.glsl (17 B/px) only feeds the reflection / direct passes on the same device.
    FINAL_PASSES = ("primary", "shadow", "gi", "reflection", "direct")
    out_atts = [a for p_ in cfg.passes if (args.outputs == "all" or p_ in FINAL_PASSES) for a in PASS_OUTPUTS[p_]]
    layout, slot_bytes = {}, 0        # attachment -> (offset in a frame slot, bytes per row, rows)
    for att in out_atts:
        _, aw, ah, bpp = ctx.attachment_info(att)
        layout[att] = (slot_bytes, aw * bpp, ah)
        slot_bytes += (aw * ah * bpp + 255) // 256 * 256
    out_bytes_px = sum(rb for _, rb, _ in layout.values()) / W

    # ---- gather to rank 0 ----
    # push (default): rank 0 owns a buffer of 2 frame slots (x world_size frames in 'frames' mode) that every rank has mapped
    #   (vxrt_cuda_shared_alloc / _open); as soon as a pass is queued its attachments - whole, or the strip's rows - are copied into the
    #   slot by the copy engines over NVLink (vxrt_cuda_copy_attachment_rows_async): no SM on either side, no kernel on rank 0, the copy
    #   overlaps the rest of the frame, and only what the consumer needs travels.
    # nccl: round 1's path (frames mode only) - the passes render into a packed buffer, one dist.gather per frame on a side stream.
    push = world_size > 1 and args.gather == "push"
    nccl_gather = world_size > 1 and not push
    if nccl_gather and tiles:
        raise SystemExit("--gather nccl is the frames-mode path of round 1; use --gather push with --sharding tiles")
    per_slot = slot_bytes * (1 if tiles else world_size)
    shared_base = None
    if push:
        handle = None
        if rank == 0:
            shared_base, handle = ctx.shared_alloc(2 * per_slot)
        box = [handle]
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            shared_base = ctx.shared_open(box[0])

    def push_hook(s, tile):
        slot_off = (s & 1) * per_slot + (0 if tiles else rank * slot_bytes)

        def hook(name, where):
            if where == "end":
                for att in PASS_OUTPUTS[name]:
                    if att in layout:
                        off, row_bytes, _ = layout[att]
                        ctx.copy_attachment_rect_async(att, shared_base + slot_off + off, tile)   # one (strided) DMA copy per attachment
        return hook

    packed, gather_dst, comm = None, None, None
    if nccl_gather:
        nl, total_bytes = [], 0
        for att in out_atts:
            off, rb, ah = layout[att]
            nl.append((att, total_bytes, rb * ah))
            total_bytes += (rb * ah + 255) // 256 * 256
        packed = [torch.empty(total_bytes, dtype=torch.uint8, device=f"cuda:{local_rank}") for _ in range(2)]
        gather_dst = list(torch.empty(world_size * total_bytes, dtype=torch.uint8, device=f"cuda:{local_rank}").chunk(world_size)) if rank == 0 else None
        comm = torch.cuda.Stream(device=local_rank)
        render_done = [torch.cuda.Event(), torch.cuda.Event()]
        gather_done = [torch.cuda.Event(), torch.cuda.Event()]
        for e in gather_done:
            e.record(stream)

    def bind_set(k):
        for att, off, n in nl:
            ctx.bind_attachment(att, packed[k].data_ptr() + off, n)

    # the inputs (37.7 MB of grids + the touched texture mips) are smaller than L2, so L2 is flushed between
    # timed steps by overwriting a 256 MiB buffer; the flush is outside the per-step event pairs
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")

    def step(s, time_hook_for=None):
        def hook_for(k):
            hooks = []
            if time_hook_for:
                hooks.append(time_hook_for(k))
            if push:
                hooks.append(push_hook(s, my_tiles[k]))
            if not hooks:
                return None
            return lambda name, where: [h(name, where) for h in hooks]

        if nccl_gather:
            k = s & 1
            stream.wait_event(gather_done[k])      # the gather that last read this buffer must be finished
            bind_set(k)
            submit(s, hook_for)
            render_done[k].record(stream)
            with torch.cuda.stream(comm):
                comm.wait_event(render_done[k])
                dist.gather(packed[k], gather_dst, dst=0)
                gather_done[k].record(comm)
        else:
            submit(s, hook_for)

    for s in range(args.warmup):
        step(s)
    if push:
        ctx.wait_reads()
    torch.cuda.synchronize()

    # ---- timed region: device-resident ----
    def ev():
        return torch.cuda.Event(enable_timing=True)

    pass_ev = [[{p_: (ev(), ev()) for p_ in cfg.passes} for _ in my_tiles] for _ in range(args.steps)]
    step_ev = [(ev(), ev()) for _ in range(args.steps)]

    def make_hook(i):
        def for_strip(k):
            def hook(name, where):
                pass_ev[i][k][name][0 if where == "begin" else 1].record(stream)
            return hook
        return for_strip

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world_size > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches_before = ctx.launch_count
    for i in range(args.steps):
        flush_buf.zero_()
        step_ev[i][0].record(stream)
        step(args.warmup + i)
        step_ev[i][1].record(stream)
    tail = ev()
    if push:
        ctx.join_reads()                   # the stream waits (on the device) for the exports still in flight
        tail.record(stream)
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
    step_ms = [a.elapsed_time(b) for a, b in step_ev]
    ms = float(sum(step_ms))  # exactly K steps, flushes excluded
    if push:
        ms += max(0.0, step_ev[-1][1].elapsed_time(tail))
    if nccl_gather:  # + whatever of the last gather is still in flight after the last step ended
        with torch.cuda.stream(comm):
            tail.record(comm)
        torch.cuda.synchronize()
        ms += max(0.0, step_ev[-1][1].elapsed_time(tail))
    launches = ctx.launch_count - launches_before
    clocks = sampler.stop() if rank == 0 else None
    # ---- the same K steps once more, serialised, for what cannot be read off overlapping kernels: per-pass times (event pairs around each
    # pass) and the launch durations of the dominant kernel (the library's probe: an event pair around every launch of the GI path-ray trace
    # kernel on its own stream).  The probe switches the pass-level and the in-GI overlap off, so each kernel runs alone, as under ncu. ----
    ctx.set_option("probe", 1)
    ctx.probe_read(reset=True)
    for i in range(args.steps):
        flush_buf.zero_()
        step(args.warmup + i, make_hook(i))
    if push:
        ctx.join_reads()
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
    probe_time = ctx.probe_read(reset=True)    # summed CUDA-event time of the probed kernel over these K steps
    ctx.set_option("probe", 0)
    pass_ms = {p_: float(np.mean([sum(strip[p_][0].elapsed_time(strip[p_][1]) for strip in pe) for pe in pass_ev])) for p_ in cfg.passes}

    # ---- did rank 0 receive what was sent?  (outside the timed region) position-weighted checksums of every exported attachment
    # (or strip) of the last two steps on the sender vs the same bytes in rank 0's buffer ----
    gather_check = None
    if push:
        ctx.wait_reads()
        torch.cuda.synchronize()
        dist.barrier()

        def csum(t):
            v = t.reshape(-1).view(torch.uint8).to(torch.int64)
            w_ = torch.arange(v.numel(), device=v.device, dtype=torch.int64) % 8191 + 1
            return int((v * w_).sum().item())

        s_last = args.warmup + args.steps - 1
        mine = {}
        # the last step's attachments are still in the context; re-render the one before it is not needed: slot (s_last & 1) only
        for att in out_atts:
            off, rb, ah = layout[att]
            full = torch.as_tensor(ctx.attachment_as_device_array(att), device=f"cuda:{local_rank}").reshape(ah, -1).view(torch.uint8)
            bpp = rb // W
            for tile in my_tiles:
                r0, nr, c0, nc = rect_of(tile, ah, W)
                mine[(att, r0, nr, c0, nc)] = csum(full[r0:r0 + nr, c0 * bpp:(c0 + nc) * bpp].contiguous())
        gathered = [None] * world_size
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            buf = torch.as_tensor(engine._DeviceArray(shared_base, (2 * per_slot,), "|u1"), device=f"cuda:{local_rank}")
            bad = 0
            for r in range(world_size):
                slot_off = (s_last & 1) * per_slot + (0 if tiles else r * slot_bytes)
                for (att, r0, nr, c0, nc), want in gathered[r].items():
                    off, rb, ah = layout[att]
                    bpp = rb // W
                    img = buf[slot_off + off: slot_off + off + ah * rb].view(ah, rb)
                    got = csum(img[r0:r0 + nr, c0 * bpp:(c0 + nc) * bpp].contiguous())
                    bad += got != want
            n_chk = sum(len(g) for g in gathered)
            gather_check = "ok (%d attachment%s of %d ranks, checksums equal)" % (n_chk, " tiles" if tiles else "s", world_size) if bad == 0 else "FAILED: %d of %d differ" % (bad, n_chk)

    # ---- distance-field regeneration (BASELINE config 2) ----
    # (a) the figure of round 1: one regeneration between a CUDA-event pair after a 256 MiB memset.  Events tick in ~2 us steps on this
    #     driver, the memset leaves the L2 full of DIRTY lines the regeneration then pays the write-back of, and the two launches
    #     come from the host one after the other; kept as "us_single_launch_after_write_flush" for comparison with BENCH_r01.
    df_ev = [(ev(), ev()) for _ in range(30)]
    for a, b in df_ev:
        flush_buf.zero_()
        a.record(stream)
        ctx.generate_distance_field()
        b.record(stream)
    torch.cuda.synchronize()
    df_us_flush = float(np.median([a.elapsed_time(b) for a, b in df_ev])) * 1e3
    # (b) the reported figure: inputs larger than the L2 instead of a flush.  8 contexts (8 x 37.7 MB of grids against 126 MB of L2)
    #     regenerate in turn, so every regeneration finds its block grid evicted and its output evicting older dirty fields (HBM
    #     traffic = the algorithmic 2 N per regeneration); 4 rounds are captured in a CUDA graph (no host launch gaps) and one event
    #     pair brackets the 32 regenerations of a replay, which also takes the event granularity out.
    df_ctxs = [ctx] + [engine.Context(local_rank) for _ in range(7)]
    for c in df_ctxs[1:]:
        c.set_stream(stream.cuda_stream)
        c.upload_world(blocks)
    for c in df_ctxs:
        c.generate_distance_field()
    torch.cuda.synchronize()
    df_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(df_graph, stream=stream):
        for _ in range(4):
            for c in df_ctxs:
                c.generate_distance_field()
    df_ev = [(ev(), ev()) for _ in range(12)]
    for a, b in df_ev:
        a.record(stream)
        df_graph.replay()
        b.record(stream)
    torch.cuda.synchronize()
    df_us = float(np.median([a.elapsed_time(b) for a, b in df_ev[2:]])) * 1e3 / 32
    del df_graph
    for c in df_ctxs[1:]:
        c.close()

    # ---- SVGF denoiser chain of the GI output: 3 x 3 pre-pass, temporal, variance, 5 a-trous iterations, per-stage CUDA
    # events, L2 flushed before each chain; runs on consecutive frames of the camera path so the history is live ----
    svgf = None
    if world_size == 1 and "gi" in cfg.passes and not args.no_svgf:
        from voxeltracing_b200.pipeline import SvgfChain
        chain = SvgfChain(ctx, W, H, pre_spatial=True)   # PreTemporalSpatialPass is on by default (Pipeline.cpp:246)
        n_chain = 20
        stage_ev = []
        for k in range(n_chain):
            cam = camera_for(wl, 1000 + k) if wl["camera"] != "rooms" else camera_for(wl, 0)   # rooms: hold one pose so frames accumulate
            fr.render(cam, 1000 + k)
            prep = chain.prepare(cam, k)
            evs = {}

            def hook(name, where, evs=evs):
                e = ev(); e.record(stream)
                evs.setdefault(name, []).append(e)
            flush_buf.zero_()
            launches1 = ctx.launch_count
            chain.submit(prep, hook=hook)
            stage_ev.append(evs)
        torch.cuda.synchronize()
        n_px = W * H

        def stage_table(frames):
            per = {}
            for evs in frames:
                for name, (a, b) in evs.items():
                    key = "spatial" if name.startswith("spatial") else name
                    per.setdefault(key, []).append(a.elapsed_time(b))
            stages = {}
            for key, v in per.items():
                calls = 5 if key == "spatial" else 1
                ms_stage = float(np.mean(v))
                by = SvgfChain.STAGE_BYTES[key] * n_px
                stages[key] = {"ms_per_launch": ms_stage, "launches_per_chain": calls, "algorithmic_bytes_per_launch": by,
                               "achieved_gbs": by / (ms_stage * 1e-3) / 1e9, "frac_of_hbm_peak": by / (ms_stage * 1e-3) / 1e9 / measured_peaks()[0]}
            return stages

        # cold: every pixel has fewer than 12 accumulated frames, so the variance pass runs its 9 x 9 pre-filter everywhere
        # (VarianceEstimate.glsl:104-107); warm: frames 14.. of a held pose (rooms) or of the orbit (history partly invalid)
        stages = stage_table(stage_ev[2:8])
        warm = stage_table(stage_ev[14:])
        chain_ms = sum(st["ms_per_launch"] * st["launches_per_chain"] for st in stages.values())
        warm_ms = sum(st["ms_per_launch"] * st["launches_per_chain"] for st in warm.values())
        svgf = {"ms_per_chain": chain_ms, "ms_per_chain_warm": warm_ms, "launches_per_chain": ctx.launch_count - launches1, "resolution": [W, H],
                "stages": stages, "stages_warm": warm, "bytes_per_pixel": SvgfChain.STAGE_BYTES, "l2": "flushed before each chain",
                "note": "ms_per_chain: frames 2..7 after a cold start (9 x 9 variance pre-filter on every pixel); warm: frames 14..19"}

    # ---- sun-shadow denoiser: temporal + spatial filter of the soft-shadow trace, per-stage CUDA events ----
    shadow_dn = None
    if world_size == 1 and "shadow" in cfg.passes and not args.no_svgf:
        from voxeltracing_b200.pipeline import ShadowDenoiser
        den = ShadowDenoiser(ctx, W, H, select=False)
        den_ev = []
        for k in range(12):
            cam = camera_for(wl, 2000 + k) if wl["camera"] != "rooms" else camera_for(wl, 0)
            fr.render(cam, 2000 + k)
            prep = den.prepare(cam, k)
            evs = {}

            def hook(name, where, evs=evs):
                e = ev(); e.record(stream)
                evs.setdefault(name, []).append(e)
            flush_buf.zero_()
            den.submit(prep, hook=hook)
            ctx.end_frame()
            den_ev.append(evs)
        torch.cuda.synchronize()
        stages = {}
        for key in ("temporal", "filter"):
            ms_stage = float(np.mean([evs[key][0].elapsed_time(evs[key][1]) for evs in den_ev[4:]]))
            by = ShadowDenoiser.STAGE_BYTES[key] * W * H
            stages[key] = {"ms_per_launch": ms_stage, "algorithmic_bytes_per_launch": by, "achieved_gbs": by / (ms_stage * 1e-3) / 1e9,
                           "frac_of_hbm_peak": by / (ms_stage * 1e-3) / 1e9 / measured_peaks()[0]}
        shadow_dn = {"ms_per_frame": sum(st["ms_per_launch"] for st in stages.values()), "resolution": [W, H], "stages": stages,
                     "bytes_per_pixel": ShadowDenoiser.STAGE_BYTES, "l2": "flushed before each frame's two launches"}

    # ---- reflection temporal filter: one launch per frame on the live reflection trace, CUDA events ----
    refl_dn = None
    if world_size == 1 and "reflection" in cfg.passes and not args.no_svgf:
        from voxeltracing_b200.pipeline import ReflectionTemporal
        # the temporal images of the engine are ReflectionSuperSampleResolution-sized; here they take the size of the trace
        rt = ReflectionTemporal(ctx, W, H, denoise=True)
        rt_ev = []
        for k in range(12):
            cam = camera_for(wl, 3000 + k) if wl["camera"] != "rooms" else camera_for(wl, 0)
            fr.render(cam, 3000 + k)
            prep = rt.prepare(cam, k)
            evs = {}

            def hook(name, where, evs=evs):
                e = ev(); e.record(stream)
                evs.setdefault(name, []).append(e)
            flush_buf.zero_()
            rt.submit(prep, hook=hook)
            ctx.end_frame()
            rt_ev.append(evs)
        torch.cuda.synchronize()
        stages = {}
        for key in ("temporal", "denoise_x", "denoise_y"):
            ms_stage = float(np.mean([evs[key][0].elapsed_time(evs[key][1]) for evs in rt_ev[4:]]))
            by = ReflectionTemporal.STAGE_BYTES[key] * W * H
            stages[key] = {"ms_per_launch": ms_stage, "algorithmic_bytes_per_launch": by, "achieved_gbs": by / (ms_stage * 1e-3) / 1e9,
                           "frac_of_hbm_peak": by / (ms_stage * 1e-3) / 1e9 / measured_peaks()[0]}
        refl_dn = {"ms_per_frame": sum(st["ms_per_launch"] for st in stages.values()), "resolution": [W, H], "stages": stages,
                   "bytes_per_pixel": ReflectionTemporal.STAGE_BYTES, "l2": "flushed before each frame's three launches"}

    # ---- end to end through the C ABI with host buffers: parameter blocks marshalled from the camera per step, the output
    # attachments (--outputs) read back to pinned host memory, inside the timed region.  Every rank reads back what it rendered
    # (its frame, or its strips of the frame) over its own PCIe link. ----
    if nccl_gather:
        torch.cuda.synchronize()
        for att, off, n in nl:
            ctx.bind_attachment(att, None)      # back to context-owned attachments for the end-to-end leg
        submit(0)
    # two sets of page-locked host buffers: frame k is copied out (vxrt_cuda_copy_attachment_rows_async, the PBO + fence
    # pattern) while frame k+1 renders; a pass that overwrites an attachment waits on the device for its copy
    host_out = [{att: torch.empty(layout[att][1] * layout[att][2], dtype=torch.uint8).pin_memory() for att in out_atts} for _ in range(2)]
    host_ptr = [{att: t.data_ptr() for att, t in hs.items()} for hs in host_out]
    my_pixels = sum(r[1] * r[3] for r in (rect_of(t, H, W) for t in my_tiles))
    d2h = int(sum(layout[att][1] // W * my_pixels for att in out_atts))
    h2d = sum(ctypes.sizeof(p_) for prep in prepared[0] for _, _, p_ in prep)

    def e2e_step(s):
        for tile in my_tiles:
            prep = fr.prepare(camera_for(wl, frame_of(s)), frame_of(s), tile)

            def read_pass_outputs(name, where, tile=tile):   # a pass's attachments start their way to the host as soon as it is queued
                if where == "end":
                    for att in PASS_OUTPUTS[name]:
                        if att in layout:
                            ctx.copy_attachment_rect_async(att, host_ptr[s & 1][att], tile)

            fr.submit(prep, hook=read_pass_outputs)

    for s in range(min(3, args.warmup)):
        e2e_step(s)
    ctx.wait_reads()
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
    e2e_steps = max(3, min(args.steps, 30))
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(args.warmup + i)
    ctx.wait_reads()                      # every frame's attachments are in host memory
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_rays = int(sum(rays_per_step[:e2e_steps]))

    # ---- parity of what was just timed (outside every timed region): the first timed frame against the CPU oracle ----
    parity = None
    if not args.no_parity_check:
        s_chk = args.warmup
        f_chk = frame_of(s_chk)
        if tiles:
            # every rank renders and exports its strips of the frame once more; rank 0 checks the ASSEMBLED frame in its buffer
            for k, prep in enumerate(prepared[s_chk]):
                fr.submit(prep, hook=push_hook(s_chk, my_tiles[k]))
            ctx.wait_reads()
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 0:
                buf = torch.as_tensor(engine._DeviceArray(shared_base, (2 * per_slot,), "|u1"), device=f"cuda:{local_rank}")
                dt = engine._ATT_DTYPES

                def read_full(att):
                    if att not in layout:
                        return None
                    off, rb, ah = layout[att]
                    raw = buf[(s_chk & 1) * per_slot + off:(s_chk & 1) * per_slot + off + rb * ah].cpu().numpy()
                    d, ch = dt[att]
                    a = raw.view(d)
                    return a.reshape(ah, -1, ch) if ch > 1 else a.reshape(ah, -1)
        else:
            if rank == 0:
                fr.render(camera_for(wl, f_chk), f_chk)

                def read_full(att):
                    try:
                        return ctx.read_attachment(att)
                    except engine.VxrtError:
                        return None
        if rank == 0:
            status, detail = parity_check(ctx, fr, wl, blocks, inputs, f_chk, read_full)
            parity = {"status": status, **detail}
        if world_size > 1:
            dist.barrier()

    # ---- N > 1: distance-field regeneration sharded by z-slabs with one boundary-plane exchange, phase by phase, beside
    # the replicated regeneration above (every rank applying the same edits and regenerating the whole field: no communication) ----
    df_sharded = None
    if world_size > 1 and 384 % world_size == 0:
        from voxeltracing_b200 import sharding
        backend = sharding.CudaSlabBackend(ctx, f"cuda:{local_rank}")
        z0 = sharding.slab_bounds(backend.nz, world_size)
        want_df = ctx.download_distance_field()
        names = ("phase_a_slab_local_sweeps", "boundary_planes_all_gather", "phase_b_apply_carries", "slabs_all_gather")
        acc = {k: [] for k in names}
        tot = []
        for it in range(12):
            flush_buf.zero_()
            dist.barrier()
            e = [ev() for _ in range(5)]
            e[0].record(stream)
            backend.phase_a(rank, z0)
            e[1].record(stream)
            first = backend.plane(z0[rank]).contiguous()
            last = backend.plane(z0[rank + 1] - 1).contiguous()
            firsts = torch.empty((world_size,) + tuple(first.shape), dtype=first.dtype, device=first.device)
            lasts = torch.empty_like(firsts)
            dist.all_gather_into_tensor(firsts.view(-1), first.view(-1))
            dist.all_gather_into_tensor(lasts.view(-1), last.view(-1))
            e[2].record(stream)
            backend.phase_b(rank, z0, firsts, lasts)
            e[3].record(stream)
            dist.all_gather_into_tensor(backend.df.view(-1), backend.df[z0[rank]:z0[rank + 1]].clone().view(-1))
            backend.commit()
            e[4].record(stream)
            torch.cuda.synchronize()
            if it >= 2:
                for j, k in enumerate(names):
                    acc[k].append(e[j].elapsed_time(e[j + 1]) * 1e3)
                tot.append(e[0].elapsed_time(e[4]) * 1e3)
        same = bool(np.array_equal(ctx.download_distance_field(), want_df))
        t = torch.tensor([float(np.median(acc[k])) for k in names] + [float(np.median(tot)), 0.0 if same else 1.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        df_sharded = {"us_per_regeneration": float(t[4]), "phases_us": {k: float(t[j]) for j, k in enumerate(names)},
                      "slabs": world_size, "planes_per_slab": 384 // world_size, "boundary_bytes_per_rank": 2 * 384 * 128,
                      "slab_bytes_per_rank": blocks.size // world_size, "identical_to_replicated": float(t[5]) == 0.0,
                      "timing": "CUDA events on the stream the kernels and the NCCL collectives share, max over ranks of the per-phase medians of 10 regenerations, L2 flushed"}

    # ---- reduce over ranks ----
    if world_size > 1:
        t = torch.tensor([ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
        r = torch.tensor([total_rays, e2e_rays, launches, total_iters, d2h], device="cuda", dtype=torch.int64)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        total_rays, e2e_rays, launches, total_iters, d2h_all = int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4])
        all_pass_ms = [None] * world_size
        dist.all_gather_object(all_pass_ms, pass_ms)
    else:
        d2h_all, all_pass_ms = d2h, [pass_ms]

    if world_size == 1:
        sharding_desc = "single GPU"
    elif tiles:
        shape = ("column band" if args.tile_shape == "cols" else "row strip") + ("s" if n_strips > 1 else "")
        sharding_desc = (f"tiles: every rank renders {n_strips} {shape} of the SAME frame per step ({'interleaved over the frame, ' if n_strips > 1 else ''}"
                         "passed as the vxrt_tile rectangle of every pass, one launch per pass per tile); the tile's rectangle of each output attachment "
                         "is copied into rank 0's frame by the copy engines over NVLink (one strided copy) as each pass finishes")
    elif push:
        sharding_desc = ("frames: one frame of the camera path per rank per step; each pass's output attachments are copied into rank 0's buffer by the copy "
                         "engines over NVLink (cudaIpc-mapped peer memory, no SM) as soon as the pass is queued, overlapping the rest of the frame")
    else:
        sharding_desc = ("frames: one frame of the camera path per rank per step; each frame's packed output attachments gathered to rank 0 "
                         "with one NCCL gather on a side stream, overlapped with the next frame (double-buffered outputs)")
    if rank == 0:
        peak, peak_src = measured_peaks()
        mrays = total_rays / (ms * 1e-3) / 1e6
        # roofline of the dominant kernel.  Algorithmic bytes per launch = rays*(S+1) (one distance-field byte per
        # iteration + the block byte) + the bytes the kernel must read / write per ray.
        dom = max(trace_passes, key=lambda p: pass_ms[p])
        if dom == "gi" and probe_time["launches"] > 0:
            # GI is a wavefront pipeline: its dominant kernel is the path-ray trace kernel, timed by the probe
            kname = "wf_trace_paths_kernel"
            n_launch = probe_time["launches"]
            k_ms = probe_time["ms"] / n_launch
            k_rays = probe_stats["rays"] / max(probe_stats["launches"], 1)
            S = probe_stats["iterations"] / max(probe_stats["rays"], 1)
            io_bytes = 40.0 * k_rays   # per ray: origin + direction (2 x float4) in, t + packed hit (8 B) out
            io_bytes_w16 = 16.0 * k_rays   # SURVEY 8(d)'s W for a GI ray: the 16 output bytes of the GI pixel, no queue I/O
        else:
            kname = PASS_KERNEL[dom]
            st = pass_stats[dom]
            S = st["iterations"] / max(st["rays"], 1)
            k_rays = st["rays"] / args.steps
            k_ms = pass_ms[dom]
            io_bytes = PASS_OUTPUT_BYTES[dom] * W * H
            io_bytes_w16 = io_bytes
        alg_bytes = k_rays * (S + 1) + io_bytes
        alg_bytes_w16 = k_rays * (S + 1) + io_bytes_w16
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        sector_bytes = k_rays * 32 * (S + 1) + io_bytes
        gather_peak_gbs = ctx.gather_peak(256) * 32 / 1e9
        try:  # DRAM bytes per launch of this kernel from the committed ncu --set full capture (profiles/traffic.json)
            traffic = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get(args.workload, {}).get(kname)
        except Exception:
            traffic = None
        line = {
            "metric": "Mrays/s DF-DDA traversal", "value": mrays, "unit": "Mrays/s", "n_gpus": world_size,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if tiles else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_block(args, wl, world_desc),     # identical in both arms
            "config_detail": {"rays_per_step_per_gpu": total_rays / args.steps / world_size,
                              "mean_iterations_per_ray": total_iters / max(total_rays, 1),
                              "sharding": sharding_desc,
                              "outputs": {"set": args.outputs, "bytes_per_pixel": out_bytes_px, "bytes_per_frame": int(out_bytes_px * W * H)}},
            "clocks": clocks,
            "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d * (world_size if not tiles else 1), "d2h_bytes_per_step": d2h_all,
                    "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3, "rank0_numa_node": numa_node,
                    "d2h_gbs_per_rank": d2h / (e2e_s / e2e_steps) / 1e9,
                    "note": "every rank reads back what it rendered (its frame / its strips) over its own PCIe link; bytes are summed over the ranks"},
            "gpu_launches": launches,
            "step_ms": {"min": float(np.min(step_ms)), "median": float(np.median(step_ms)), "max": float(np.max(step_ms)), "of": "rank 0's K timed steps"},
            "parity_check": parity["status"] if parity else None, "parity": parity,
            "gather_check": gather_check,
            "pass_overlap": not args.no_pass_overlap,
            "wf_bands": args.wf_bands,
            "pass_ms": pass_ms,
            "pass_ms_note": "passes timed one after the other in K extra steps after the timed region (CUDA-event pairs around each pass, pass-level overlap "
                            "off); with pass_overlap the timed step is shorter than their sum because shadow / reflection / direct run beside the GI wavefront",
            "pass_ms_sum": float(sum(pass_ms.values())),
            "pass_ms_last_rank": all_pass_ms[-1] if world_size > 1 else None,
            "pass_mrays": {p: pass_stats[p]["rays"] / args.steps / (pass_ms[p] * 1e-3) / 1e6 for p in trace_passes},
            "roofline": {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_convention": "rays*(S+1) + 40 B/ray of queue I/O (origin + direction in, t + hit out); *_w16: SURVEY 8(d)'s W = 16 B/ray instead",
                         "achieved_w16": alg_bytes_w16 / (k_ms * 1e-3) / 1e9, "frac_w16": alg_bytes_w16 / (k_ms * 1e-3) / 1e9 / peak,
                         "l2_gather_achieved": sector_bytes / (k_ms * 1e-3) / 1e9, "l2_gather_peak": gather_peak_gbs,
                         "l2_gather_frac": sector_bytes / (k_ms * 1e-3) / 1e9 / gather_peak_gbs,
                         "mean_iterations_per_ray": S, "rays_per_launch": k_rays, "avg_launch_ms": k_ms,
                         "timing": "CUDA-event pair around every launch of the kernel on its own stream (the library's probe), live in this run, over the K steps "
                                   "repeated serialised right after the timed region (overlapping kernels have no launch duration of their own)",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_mrays": k_rays / (k_ms * 1e-3) / 1e6,
                         "l2_gather": {"achieved": sector_bytes / (k_ms * 1e-3) / 1e9, "peak": gather_peak_gbs, "unit": "GB/s",
                                       "frac": sector_bytes / (k_ms * 1e-3) / 1e9 / gather_peak_gbs,
                                       "peak_source": "vxrt_cuda_gather_peak: independent random 1-byte loads of the L2-resident distance field, 32 B per sector, measured in this run",
                                       "note": "the model charges one 32-byte sector per iteration; coherent rays (primary, sun shadow) get 85-95 % of them from L1, so frac can exceed 1 there"},
                         "note": "grids are L2 resident (dram throughput < 1 % in profiles/): HBM is the contract's roof, the operative one is issue slots, then the L2/L1 gather rate (l2_gather); see DESIGN.md §3"},
        }
        nvox = blocks.size
        line["df_regen"] = {"us_per_regeneration": df_us, "algorithmic_bytes": 2 * nvox,
                            "achieved_gbs": 2 * nvox / (df_us * 1e-6) / 1e9, "frac_of_hbm_peak": 2 * nvox / (df_us * 1e-6) / 1e9 / peak,
                            "method": "8 contexts regenerate in turn (302 MB of grids > L2), 32 regenerations per CUDA-graph replay between one event pair, median of 10 replays",
                            "us_single_launch_after_write_flush": df_us_flush, "launches_per_regeneration": 2}
        if df_sharded:
            line["df_regen_sharded"] = {**df_sharded, "replicated_us_per_regeneration": df_us,
                                        "note": "replicated = every rank applies the same edit list and regenerates the whole field (no communication): the fast path"}
        if svgf:
            line["svgf"] = svgf
        if shadow_dn:
            line["shadow_denoiser"] = shadow_dn
        if refl_dn:
            line["reflection_denoiser"] = refl_dn
        if world_size == 1 and not args.no_svgf:
            line["world_producers"] = world_producers_block(local_rank, stream, flush_buf, ev, peak)
            line["lpv"] = lpv_block(local_rank, stream, flush_buf, ev, peak, not args.no_cpu_baseline)
        if not args.no_cpu_baseline and world_size == 1:
            cpu = cpu_arm(blocks, wl, inputs)
            v, ms_step, n, band = cpu_sample(cpu, wl, args.warmup, args.steps, args.cpu_seconds, CPU_FRAME_BUDGET_S)
            line["cpu_baseline"] = {"value": v, "unit": "Mrays/s", "cores": cpu.cores, "kind": cpu.kind,
                                    "sample": f"rows [{band[0]},{band[0] + band[1]}) of {n} frames of the same camera path"}
        else:
            line["cpu_baseline"] = None
        emit(line)

    if push:
        ctx.wait_reads()
        torch.cuda.synchronize()
        if rank != 0:
            ctx.shared_close(shared_base)
        dist.barrier()
        if rank == 0:
            ctx.shared_free(shared_base)
    ctx.close()
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
