"""The screen-space filter chains (SVGF of the GI, sun-shadow denoiser, reflection temporal + denoiser) on config-4 frames at 1080p:
per-chain device time and - against the bit-faithful mode (filter_snap = 0) - the error of the tolerance mode.
usage: time_filters.py [snap,snap,...]   (units of 1/65536; 256 = 1/256)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import bench, scene_util as su
from voxeltracing_b200 import abi, engine
from voxeltracing_b200.pipeline import FrameRenderer, ReflectionTemporal, ShadowDenoiser, SvgfChain

snaps = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,256,1024").split(",")]
wl = bench.WORKLOADS["config4_1080p_gi"]
W, H = wl["width"], wl["height"]
blocks, _ = bench.build_world(wl["world"])
inputs = su.SceneInputs(512, sky="constant")
ctx = engine.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
ctx.upload_world(blocks); ctx.generate_distance_field(); ctx.set_blue_noise_texture(bench.BLUE_TEX); inputs.apply_to_context(ctx)
fr = FrameRenderer(ctx, bench.frame_config(wl), inputs.grass, inputs.cactus)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
N = 10
OUT = {"svgf": (abi.ATT_SVGF_DENOISE_A, abi.ATT_SVGF_DENOISE_A + 1, abi.ATT_SVGF_DENOISE_B, abi.ATT_SVGF_DENOISE_B + 1),
       "shadow": (abi.ATT_SHADOW_FILTERED,), "reflection": (abi.ATT_REFL_DENOISED_A, abi.ATT_REFL_DENOISED_B)}


def run(snap):
    ctx.set_option("filter_snap", snap)
    chains = {"svgf": SvgfChain(ctx, W, H, pre_spatial=True), "shadow": ShadowDenoiser(ctx, W, H), "reflection": ReflectionTemporal(ctx, W, H, denoise=True)}
    ms = {k: [] for k in chains}
    for k in range(N):
        cam = bench.camera_for(wl, k // 3)        # a pose held for three frames, then a cut: history both valid and invalid
        fr.render(cam, k)
        for name, ch in chains.items():
            prep = ch.prepare(cam, k)
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); ch.submit(prep); b.record(stream)
            torch.cuda.synchronize()
            if k >= 2:
                ms[name].append(a.elapsed_time(b))
    outs = {}
    for name, atts in OUT.items():
        for att in atts:
            try:
                outs[(name, att)] = ctx.read_attachment(att).astype(np.float32)
            except engine.VxrtError:
                pass
    return {k: float(np.mean(v)) for k, v in ms.items()}, outs


ref_ms, ref = run(0)
print("snap=0 (bit-faithful):", " ".join(f"{k}={v:.3f} ms" for k, v in ref_ms.items()), "total=%.3f ms" % sum(ref_ms.values()), flush=True)
for snap in snaps:
    if snap == 0:
        continue
    ms, outs = run(snap)
    errs = []
    for key, want in ref.items():
        got = outs[key]
        ok = np.isfinite(want) & np.isfinite(got)
        d = np.abs(got - want)[ok]
        scale = np.abs(want)[ok]
        rel = d / (scale + 1e-2)
        errs.append(f"{key[0]}:{key[1]} max_rel={rel.max():.2e} p99.9={np.percentile(rel, 99.9):.2e} within_1e-2={(rel <= 1e-2).mean():.5f}")
    print(f"snap={snap} ({snap / 65536:.5f}):", " ".join(f"{k}={v:.3f} ms" for k, v in ms.items()), "total=%.3f ms" % sum(ms.values()))
    for e in errs:
        print("    ", e, flush=True)
ctx.close()
