"""Parity at the configurations bench.py times (not only at the small frames of the other tests): BASELINE config 4 (1920 x 1080, 512^2
block textures, GI 1 spp, reflections) and config 5 (3840 x 2160, GI 4 spp - the sample index walks past the ranking tile,
DiffuseRayTraceFrag.glsl:139-148), a 32-row band of one frame of the benched camera path against the oracle, through the same
bench.parity_check the bench line's "parity_check" comes from; plus the band rendered as a vxrt_tile equals the band of the full frame."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import scene_util as su  # noqa: E402
from voxeltracing_b200 import abi, engine  # noqa: E402
from voxeltracing_b200.pipeline import FrameRenderer  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def inputs512():
    return {"constant": su.SceneInputs(512, sky="constant"), "gradient": None}


def _setup(workload, inputs512):
    wl = bench.WORKLOADS[workload]
    sky = "constant" if wl["camera"] == "rooms" else "gradient"
    if inputs512[sky] is None:
        inputs512[sky] = su.SceneInputs(512, sky=sky)
    inputs = inputs512[sky]
    blocks, _ = bench.build_world(wl["world"])
    ctx = engine.Context(0)
    ctx.upload_world(blocks)
    ctx.generate_distance_field()
    ctx.set_blue_noise_texture(bench.BLUE_TEX)
    inputs.apply_to_context(ctx)
    fr = FrameRenderer(ctx, bench.frame_config(wl), inputs.grass, inputs.cactus)
    return wl, blocks, inputs, ctx, fr


@pytest.mark.parametrize("workload,frame", [("config4_1080p_gi", 3), ("config4_1080p_gi", 10), ("config5_4k_gi4", 5), ("config3_1080p_direct", 2)])
def test_benched_configuration_matches_the_oracle_on_a_band(workload, frame, inputs512):
    wl, blocks, inputs, ctx, fr = _setup(workload, inputs512)
    try:
        fr.render(bench.camera_for(wl, frame), frame)
        status, detail = bench.parity_check(ctx, fr, wl, blocks, inputs, frame, ctx.read_attachment)
        assert status == "ok", detail
        assert len(detail["checks"]) >= 9
        # the same band rendered as a screen tile (what --sharding tiles does per strip) equals the band of the full frame bit for bit
        r0, r1 = detail["rows"]
        full = {a: ctx.read_attachment(a).copy() for a in fr.outputs}
        fr.render(bench.camera_for(wl, frame), frame, tile=(r0, r1 - r0))
        for a in fr.outputs:
            got = ctx.read_attachment(a)
            assert np.array_equal(got[r0:r1].view(np.uint8), full[a][r0:r1].view(np.uint8)), (workload, a)
        # ... and so does a column band crossing it, with edges off the CTA grid (what --sharding tiles does per rank)
        Wf = wl["width"]
        c0, cw = Wf // 3 + 5, Wf // 4 + 3
        import torch
        ctx.synchronize()
        for a in fr.outputs:   # poison the rectangle (only it: passes read their inputs' neighbours across the tile edge): the tile must rewrite it
            torch.as_tensor(ctx.attachment_as_device_array(a), device="cuda")[r0:r1, c0:c0 + cw].fill_(77)
        torch.cuda.synchronize()
        fr.render(bench.camera_for(wl, frame), frame, tile=(r0, r1 - r0, c0, cw))
        for a in fr.outputs:
            got = ctx.read_attachment(a)
            assert np.array_equal(got[r0:r1, c0:c0 + cw].view(np.uint8), full[a][r0:r1, c0:c0 + cw].view(np.uint8)), (workload, a, "rect")
    finally:
        ctx.close()


@pytest.mark.parametrize("workload,reproject", [("config4_1080p_gi", False), ("config4_1080p_gi", True), ("config5_4k_gi4", False)])
def test_pass_overlap_does_not_change_a_bit(workload, reproject, inputs512):
    """set_option("pass_overlap", 1): the sun-shadow trace and the direct term on the context's second stream beside the GI / reflection wavefronts.
    Every output attachment of the frame equals the single-stream frame bit for bit - with and without reprojection in the reflection pass, with a
    read-back queued in the middle of the frame (the push hooks of the multi-GPU bench), and across frames (the next frame's primary pass
    overwrites what lane 1 was reading)."""
    import dataclasses
    wl, blocks, inputs, ctx, fr = _setup(workload, inputs512)
    try:
        if reproject:
            fr.cfg = dataclasses.replace(fr.cfg, refl_reproject=True)
        frames = (3, 4, 5)

        def run(overlap, mid_frame_read):
            ctx.set_option("pass_overlap", overlap)
            out = []
            for f in frames:
                sink = {}

                def hook(name, where):
                    if mid_frame_read and where == "end" and name == "gi":
                        sink["gi"] = ctx.read_attachment(abi.ATT_GI_SH).copy()
                fr.render(bench.camera_for(wl, f), f, hook=hook)
                out.append({a: ctx.read_attachment(a).copy() for a in fr.outputs})
            return out
        want = run(0, False)
        # ... and with the wavefront passes cut into row bands on their own streams (set_option "wf_bands"), alone and together with the lanes
        for bands, overlap, mid in ((1, 1, False), (1, 1, True), (2, 0, False), (3, 1, True), (4, 1, False)):
            ctx.set_option("wf_bands", bands)
            got = run(overlap, mid)
            for fw, fg in zip(want, got):
                for a in fr.outputs:
                    assert np.array_equal(fw[a].view(np.uint8), fg[a].view(np.uint8)), (workload, reproject, bands, overlap, mid, a)
    finally:
        ctx.set_option("pass_overlap", 0); ctx.set_option("wf_bands", 1)
        ctx.close()


@pytest.mark.parametrize("refl_spp", [1, 2])
def test_lane2_direct_deferred_reflection_gi_and_lane_copies_do_not_change_a_bit(refl_spp, inputs512):
    """The last additions to the pass-level concurrency, each behind its option (all on by default): the direct term on lane 2 beside a pending
    reflection pass ("lane2_direct"), the material G-buffer on lane 1 beside the GI's first kernels ("lane1_gbuffer"), the reflection pass meeting the GI only where a sample is accumulated ("refl_defer_gi": Albedo and the AO
    factor travel instead of the ambient product, formed later from the same operands in the same order), and read-backs of a lane-1 / lane-2
    attachment that wait for that lane alone on a second copy stream ("copy_lanes").  Every output attachment equals the single-stream frame with
    all three off, bit for bit: read synchronously after the frame, and through asynchronous copies queued right behind each pass (the bench's
    end-to-end leg) over three consecutive frames into two alternating sets of page-locked buffers."""
    import dataclasses
    import torch
    from voxeltracing_b200.pipeline import PASS_OUTPUTS
    wl, blocks, inputs, ctx, fr = _setup("config4_1080p_gi", inputs512)
    try:
        fr.cfg = dataclasses.replace(fr.cfg, refl_spp=refl_spp)
        frames = (3, 4, 5)

        def run(overlap, lane2, defer, copy_lanes, async_copies):
            ctx.set_option("pass_overlap", overlap); ctx.set_option("lane2_direct", lane2 & 1); ctx.set_option("lane1_gbuffer", lane2 >> 1)
            ctx.set_option("refl_defer_gi", defer); ctx.set_option("copy_lanes", copy_lanes)
            out = []
            host = [{}, {}]
            for k, f in enumerate(frames):
                def hook(name, where, k=k):
                    if async_copies and where == "end":
                        for a in PASS_OUTPUTS[name]:
                            _, w, h, bpp = ctx.attachment_info(a)
                            t = host[k & 1].get(a)
                            if t is None:
                                t = host[k & 1][a] = torch.empty(w * h * bpp, dtype=torch.uint8).pin_memory()
                            ctx.copy_attachment_rect_async(a, t.data_ptr(), (0, 0))
                fr.render(bench.camera_for(wl, f), f, hook=hook)
                if async_copies:
                    if k == 0:   # frame 0 is read out before its buffers are reused; frames 1 and 2 stay in flight: frame 2's passes rewrite the
                        ctx.wait_reads()   # attachments frame 1's copies are still reading, ordered only by the device-side waits
                        out.append({a: host[0][a].numpy().copy() for a in fr.outputs})
                else:
                    out.append({a: ctx.read_attachment(a).copy().view(np.uint8).ravel() for a in fr.outputs})
            if async_copies:
                ctx.wait_reads()
                out.append({a: host[1][a].numpy().copy() for a in fr.outputs})
                out.append({a: host[0][a].numpy().copy() for a in fr.outputs})
            return out
        want = run(0, 0, 0, 0, False)
        # lane2: bit 0 = lane2_direct, bit 1 = lane1_gbuffer
        for overlap, lane2, defer, copy_lanes, async_copies in ((0, 0, 1, 0, False), (1, 0, 0, 0, False), (1, 1, 0, 0, False), (1, 2, 0, 0, False),
                                                                (1, 3, 1, 1, False), (1, 3, 1, 1, True), (1, 3, 1, 0, True), (1, 2, 1, 1, True),
                                                                (1, 1, 1, 1, True)):
            ctx.set_option("wf_bands", 2 if lane2 else 1)
            got = run(overlap, lane2, defer, copy_lanes, async_copies)
            for fw, fg in zip(want, got):
                for a in fr.outputs:
                    assert np.array_equal(fw[a], fg[a]), (refl_spp, overlap, lane2, defer, copy_lanes, async_copies, a)
    finally:
        for k_, v_ in (("pass_overlap", 0), ("wf_bands", 1), ("lane2_direct", 1), ("lane1_gbuffer", 0), ("refl_defer_gi", 1), ("copy_lanes", 1)):
            ctx.set_option(k_, v_)
        ctx.close()
