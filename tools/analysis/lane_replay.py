#!/usr/bin/env python
"""Replay of per-ray iteration counts (CPU, oracle) to predict what scheduling of the GI bounce rays buys on the GPU.

A warp of the queue trace kernels runs until its slowest ray ends, so its cost is max(iterations) over its 32 lanes and the lane
efficiency is mean / max.  This tool builds the bounce-0 and bounce-1 rays of a config-4 frame (primary hits of the oracle, cosine
weighted directions from a seeded generator - same distribution as the blue-noise sampler), gets every ray's iteration count from
the oracle's VoxelTraversalDF and evaluates:
  baseline        32 consecutive rays of the 8 x 4 pixel tiling (what wf_trace_paths_kernel does)
  capped relaunch trace <= K iterations, compact the survivors, relaunch (caps K1 < K2 < ...)
  binned          rays sorted by a key (direction octant, coarse origin cell) before forming warps
usage: lane_replay.py [width height]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import binding as ob  # noqa: E402
from voxeltracing_b200 import abi, host_api, pipeline  # noqa: E402

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (960, 540)
FACE_N = np.array([[0, 0, 1], [0, 0, -1], [0, 1, 0], [0, -1, 0], [-1, 0, 0], [1, 0, 0]], np.float32)


def cos_hemisphere(n, rng):
    u, v = rng.random(len(n), dtype=np.float32), rng.random(len(n), dtype=np.float32)
    r, phi = np.sqrt(u), 2 * np.pi * v
    x, y, z = r * np.cos(phi), r * np.sin(phi), np.sqrt(np.maximum(0, 1 - u))
    # tangent frame
    a = np.where(np.abs(n[:, :1]) > 0.9, np.array([[0, 1, 0]], np.float32), np.array([[1, 0, 0]], np.float32))
    t = np.cross(n, a); t /= np.linalg.norm(t, axis=1, keepdims=True)
    b = np.cross(n, t)
    d = t * x[:, None] + b * y[:, None] + n * z[:, None]
    return (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)


def tile_order(w, h):
    """ray index -> position in the 8 x 4 tiling the kernels use (a warp = an 8 x 4 pixel tile, CTA = 32 x 8 pixels)"""
    ys, xs = np.mgrid[0:h, 0:w]
    key = ((ys // 8) * ((w + 31) // 32) + xs // 32) * 256 + ((ys % 8) // 4 * 4 + (xs % 32) // 8) * 32 + (ys % 4) * 8 + xs % 8
    return np.argsort(key.ravel(), kind="stable")


def warp_cost(iters):
    n = len(iters) // 32 * 32
    m = iters[:n].reshape(-1, 32)
    return int(m.max(axis=1).sum()), float(m.sum() / max(m.max(axis=1).sum() * 32, 1))


def capped(iters, caps):
    """warp-iterations when rays run in passes of at most caps[j] more iterations and the survivors are compacted (order kept)"""
    total, done, rem = 0, 0, iters.copy()
    for k in caps:
        if len(rem) == 0:
            break
        step = np.minimum(rem, k - done)
        pad = (-len(step)) % 32
        m = np.concatenate([step, np.zeros(pad, step.dtype)]).reshape(-1, 32)
        total += int(m.max(axis=1).sum())
        rem = rem[rem > k - done] - (k - done)
        done = k
    return total


def spill(iters, thresholds):
    """warp-iterations when a warp hands its survivors to a continuation queue as soon as <= T lanes are still running
    (thresholds[j] for pass j; the last pass runs to the end) and the next pass starts from the dense continuation queue"""
    total, rem = 0, iters.copy()
    moved = []
    for j in range(len(thresholds) + 1):
        if len(rem) == 0:
            break
        pad = (-len(rem)) % 32
        m = np.concatenate([rem, np.zeros(pad, rem.dtype)]).reshape(-1, 32)
        if j == len(thresholds):
            total += int(m.max(axis=1).sum())
            break
        T = thresholds[j]
        srt = np.sort(m, axis=1)[:, ::-1]          # descending: srt[:, T] = iteration count after which <= T lanes are running
        stop = srt[:, T]                           # the warp runs `stop` iterations (lanes with more go on to the next pass)
        total += int(stop.sum())
        left = m - stop[:, None]
        moved.append(int((left > 0).sum()))
        rem = left[left > 0]
    return total, moved


def main():
    blocks = host_api.gen_world("rooms", 2)
    ow = ob.OracleWorld(blocks)
    rng = np.random.default_rng(3)
    rows = []
    for frame in (0, 3, 11):
        cam = pipeline.rooms_camera(frame, W / H)
        p = abi.PrimaryParams()
        for i in range(16):
            p.inv_view[i] = float(cam.inv_view[i]); p.inv_projection[i] = float(cam.inv_projection[i])
        p.width, p.height, p.render_distance = W, H, 350
        g = ow.initial_trace(p)
        t = g["t32"].ravel()
        hit = t > 0
        # reconstruct P like the GI pass does (camera + dir * t); directions from the oracle's own unprojection are not exported,
        # so trace the primary rays again through the batch API for the end points
        ys, xs = np.mgrid[0:H, 0:W]
        uv = np.stack([(xs.ravel() + 0.5) / W, (ys.ravel() + 0.5) / H], 1).astype(np.float32)
        clip = np.concatenate([uv * 2 - 1, -np.ones((len(uv), 1), np.float32), np.ones((len(uv), 1), np.float32)], 1)
        ip = np.asarray(cam.inv_projection, np.float32).reshape(4, 4).T
        iv = np.asarray(cam.inv_view, np.float32).reshape(4, 4).T
        eye = (clip @ ip.T); eye[:, 2] = -1; eye[:, 3] = 0
        d = (eye @ iv.T)[:, :3]; d /= np.linalg.norm(d, axis=1, keepdims=True)
        P = np.asarray(cam.position, np.float32)[None] + d * t[:, None]
        face = np.clip(np.rint(g["normal"].ravel().astype(np.float32) / 255 * 10).astype(int), 0, 5)
        N = FACE_N[face]
        order = tile_order(W, H)
        order = order[hit[order]]
        o0 = (P + N * 0.06)[order].astype(np.float32)
        d0 = cos_hemisphere(N[order], rng)
        h0 = ow.traverse_batch(o0, d0, 48)
        it0 = h0["iterations"].astype(np.int64)
        # bounce 1: from the bounce-0 hits (compacted in warp order, like the queue)
        ok = h0["t"] > 0
        o1 = (h0["end"][ok] + h0["normal"][ok] * 0.06).astype(np.float32)
        d1 = cos_hemisphere(h0["normal"][ok].astype(np.float32), rng)
        h1 = ow.traverse_batch(o1, d1, 48)
        it1 = h1["iterations"].astype(np.int64)
        for name, it, o, dd in (("bounce0", it0, o0, d0), ("bounce1", it1, o1, d1)):
            base, eff = warp_cost(it)
            r = {"frame": frame, "rays": name, "n": len(it), "mean": float(it.mean()), "p50": float(np.median(it)), "p90": float(np.percentile(it, 90)),
                 "max": int(it.max()), "baseline_warp_iters": base, "lane_eff": eff}
            for caps in ((8, 48), (12, 48), (16, 48), (8, 20, 48), (12, 24, 48), (6, 12, 24, 48), (4, 8, 16, 32, 48)):
                r["cap" + "_".join(map(str, caps))] = capped(it, caps) / base
            for th in ((4,), (8,), (12,), (16,), (8, 8), (12, 8), (16, 8), (16, 12, 8), (12, 12, 12)):
                tot, moved = spill(it, th)
                r["spill" + "_".join(map(str, th))] = (round(tot / base, 3), [round(x / len(it), 3) for x in moved])
            octant = (dd[:, 0] > 0).astype(int) | ((dd[:, 1] > 0).astype(int) << 1) | ((dd[:, 2] > 0).astype(int) << 2)
            cell = (o[:, 0].astype(int) >> 3) + ((o[:, 1].astype(int) >> 3) << 6) + ((o[:, 2].astype(int) >> 3) << 12)
            for kname, key in (("sort_octant", octant), ("sort_cell_octant", cell * 8 + octant), ("sort_octant_cell", octant * (1 << 20) + cell)):
                s = np.argsort(key, kind="stable")
                r[kname] = warp_cost(it[s])[0] / base
            # oracle for binning: perfect sort by the iteration count itself
            r["sort_by_iters(ideal)"] = warp_cost(np.sort(it))[0] / base
            rows.append(r)
    for r in rows:
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()})


if __name__ == "__main__":
    main()
