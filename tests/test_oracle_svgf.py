"""SVGF chain of the diffuse GI: oracle self-checks and oracle == the reference's own shaders compiled
for the CPU (oracle/_ref) on a two-frame sequence.  CPU only."""
import ctypes as C

import numpy as np
import pytest

import svgf_util as sv
from oracle import binding as ob
from oracle import ref_binding as rb
from voxeltracing_b200 import abi, host_api


@pytest.fixture(scope="module")
def seq():
    return sv.frames(host_api.gen_world("plains", 0))


def _same(a, b):
    return all(np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)) for k in ("sh", "cocg", "x", "aosky"))


def _temporal_frames(seq, fn=None):
    """frame 0 against an all-zero history, frame k against frame k-1's output"""
    hist = ob.svgf_alloc(sv.H, sv.W, 3)
    prev_g, prev_cam = sv.zero_gbuf(), seq[0]["cam"]
    outs = []
    for f in seq:
        p = sv.temporal_params(f["cam"], prev_cam, abi.ATT_GI_SH, abi.ATT_SVGF_TEMPORAL_B, abi.ATT_SVGF_TEMPORAL_A)
        out = ob.svgf_temporal(p, f["raw"], hist, f["g"], prev_g, fn)
        outs.append(out)
        hist, prev_g, prev_cam = out, f["g"], f["cam"]
    return outs


def test_temporal_accumulates_where_the_history_is_valid(seq):
    o0, o1, o2 = _temporal_frames(seq)
    # frame 0: an all-zero history reconstructs every previous position at the camera, so only surfaces within the
    # position tolerance of the eye can match it: the rest passes through with zero accumulated frames
    fresh = o0["x"][..., 0] == 0
    assert fresh.mean() > 0.9
    # (up to the bilinear weights of texture() at a pixel centre, which are not exactly 0 / 1 in float)
    assert np.allclose(o0["sh"].astype(np.float32)[fresh], seq[0]["raw"]["sh"].astype(np.float32)[fresh], atol=2e-3)
    assert (np.abs(seq[0]["raw"]["sh"].astype(np.float32)) > 0.01).mean() > 0.3
    # frame 1: pixels that reproject onto frame 0 count one accumulated frame (blend factor still 1); frame 2: two, and
    # the output is a blend of history and current sample
    acc1, acc2 = o1["x"][..., 0].astype(np.float32), o2["x"][..., 0].astype(np.float32)
    assert (acc1 >= 0.99).mean() > 0.25 and (acc1 == 0).mean() > 0.005
    assert (acc2 >= 1.9).mean() > 0.2
    moved = np.abs(o2["sh"].astype(np.float32) - seq[2]["raw"]["sh"].astype(np.float32)).max(-1) > 1e-3
    assert moved[acc2 >= 1.9].mean() > 0.5 and moved[acc2 == 0].mean() < 0.05
    # second moment and luminance history are carried
    assert np.isfinite(o2["x"].astype(np.float32)).all() and (o2["x"][..., 1].astype(np.float32) >= 0).all()


@pytest.mark.skipif(not rb.available("svgf_temporal"), reason="oracle/_ref not built on this box")
@pytest.mark.parametrize("be_useful", [True, False])
def test_oracle_temporal_equals_compiled_reference_shader(seq, be_useful):
    L = rb.lib()
    a = _temporal_frames(seq)
    b = _temporal_frames(seq, L.vxref_svgf_temporal)
    assert all(_same(x, y) for x, y in zip(a, b)) and len(a) == 3
    if not be_useful:
        hist, prev_g = a[0], seq[0]["g"]
        p = sv.temporal_params(seq[1]["cam"], seq[0]["cam"], abi.ATT_GI_SH, abi.ATT_SVGF_TEMPORAL_B, abi.ATT_SVGF_TEMPORAL_A, be_useful=False)
        assert _same(ob.svgf_temporal(p, seq[1]["raw"], hist, seq[1]["g"], prev_g), ob.svgf_temporal(p, seq[1]["raw"], hist, seq[1]["g"], prev_g, L.vxref_svgf_temporal))
