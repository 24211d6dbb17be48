#!/usr/bin/env python
"""Can the iteration count of a GI bounce ray be PREDICTED cheaply enough to bin the bounce queue by it?

lane_replay.py showed that sorting by direction octant / origin cell leaves 88 - 98 % of the warp-iterations.  This tool asks the same
question of predictors a shading kernel could compute when it emits the ray (one to eight extra 1-byte reads of the L2-resident distance
field): cos(theta) against the surface normal, the distance-field value at look-ahead points along the ray, sums of them.  Rays are
binned into B quantile bins of the predictor (stable inside a bin, i.e. the queue keeps its screen order) and warps are formed from 32
consecutive rays; the cost of a warp is the iteration count of its slowest ray.  "ideal" bins by the true iteration count.
usage: lane_predict.py [width height]      (CPU only: oracle traversal of a config-4 frame of the `rooms` world)
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import lane_replay as lr  # noqa: E402
from oracle import binding as ob  # noqa: E402
from voxeltracing_b200 import abi, host_api, pipeline  # noqa: E402

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (960, 540)


def main():
    blocks = host_api.gen_world("rooms", 2)
    ow = ob.OracleWorld(blocks)
    nz, ny, nx = blocks.shape
    df3 = ow.df.reshape(nz, ny, nx)
    rng = np.random.default_rng(3)

    def df_at(p):
        i = np.floor(p).astype(int)
        ok = (i[:, 0] >= 0) & (i[:, 0] < nx) & (i[:, 1] >= 0) & (i[:, 1] < ny) & (i[:, 2] >= 0) & (i[:, 2] < nz)
        v = np.zeros(len(p), np.int64)
        v[ok] = df3[i[ok, 2], i[ok, 1], i[ok, 0]]
        return v

    for frame in (0, 3, 11):
        cam = pipeline.rooms_camera(frame, W / H)
        p = abi.PrimaryParams()
        for i in range(16):
            p.inv_view[i] = float(cam.inv_view[i]); p.inv_projection[i] = float(cam.inv_projection[i])
        p.width, p.height, p.render_distance = W, H, 350
        g = ow.initial_trace(p)
        t = g["t32"].ravel()
        hit = t > 0
        ys, xs = np.mgrid[0:H, 0:W]
        uv = np.stack([(xs.ravel() + 0.5) / W, (ys.ravel() + 0.5) / H], 1).astype(np.float32)
        clip = np.concatenate([uv * 2 - 1, -np.ones((len(uv), 1), np.float32), np.ones((len(uv), 1), np.float32)], 1)
        ip = np.asarray(cam.inv_projection, np.float32).reshape(4, 4).T
        iv = np.asarray(cam.inv_view, np.float32).reshape(4, 4).T
        eye = clip @ ip.T; eye[:, 2] = -1; eye[:, 3] = 0
        d = (eye @ iv.T)[:, :3]; d /= np.linalg.norm(d, axis=1, keepdims=True)
        P = np.asarray(cam.position, np.float32)[None] + d * t[:, None]
        face = np.clip(np.rint(g["normal"].ravel().astype(np.float32) / 255 * 10).astype(int), 0, 5)
        N = lr.FACE_N[face]
        order = lr.tile_order(W, H)
        order = order[hit[order]]
        n0 = N[order]
        o0 = (P + N * 0.06)[order].astype(np.float32)
        d0 = lr.cos_hemisphere(n0, rng)
        it = ow.traverse_batch(o0, d0, 48)["iterations"].astype(np.int64)
        base, eff = lr.warp_cost(it)
        print(f"frame {frame}: {len(it)} bounce rays, mean {it.mean():.1f} iterations, {eff * 32:.1f} of 32 lanes in screen order")
        preds = {"cos(theta)": (d0 * n0).sum(1)}
        for L in (2, 4, 8, 16):
            preds[f"df@{L}"] = -df_at(o0 + d0 * L)
        preds["df@2+4+8"] = -(df_at(o0 + d0 * 2) + df_at(o0 + d0 * 4) + df_at(o0 + d0 * 8))
        preds["df@1+2+3+4+6+8+12+16"] = -sum(df_at(o0 + d0 * L) for L in (1, 2, 3, 4, 6, 8, 12, 16))
        for B in (2, 4, 8):
            for name, key in preds.items():
                bins = np.searchsorted(np.quantile(key, np.linspace(0, 1, B + 1)[1:-1]), key)
                c, e = lr.warp_cost(it[np.argsort(bins, kind="stable")])
                print(f"  {B} bins by {name:22s} warp-iterations x {c / base:.3f}  lanes {e * 32:4.1f}   (correlation with the count {np.corrcoef(key, it)[0, 1]:+.2f})")
            bins = np.searchsorted(np.quantile(it, np.linspace(0, 1, B + 1)[1:-1]), it)
            c, e = lr.warp_cost(it[np.argsort(bins, kind="stable")])
            print(f"  {B} bins by the count itself (ideal) warp-iterations x {c / base:.3f}  lanes {e * 32:4.1f}")


if __name__ == "__main__":
    main()
