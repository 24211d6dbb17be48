"""Thin numpy wrappers over libvxrt_host.so (camera, jitter, sun, worlds).  No compute of the hot path."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from . import abi

WORLD_DIMS = (384, 128, 384)  # Core/Macros.h:3-5


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class Camera:
    """FPSCamera state reduced to what the passes consume (Core/FpsCamera.cpp, Core/Player.cpp:10)."""

    view: np.ndarray
    projection: np.ndarray
    inv_view: np.ndarray
    inv_projection: np.ndarray
    position: np.ndarray


def camera(pos, yaw_deg: float, pitch_deg: float, aspect: float, fov_deg: float = 90.0) -> Camera:
    lib = abi.load_host()
    pos = np.asarray(pos, dtype=np.float32)
    out = [np.zeros(16, dtype=np.float32) for _ in range(4)]
    lib.vxh_camera(_p(pos), yaw_deg, pitch_deg, fov_deg, aspect, *[_p(o) for o in out])
    return Camera(out[0], out[1], out[2], out[3], pos)


def taa_jitter(frame: int) -> np.ndarray:
    out = np.zeros(2, dtype=np.float32)
    abi.load_host().vxh_taa_jitter(frame, _p(out))
    return out


def sun_direction(sun_tick: float = 50.0):
    s, m, st = (np.zeros(3, dtype=np.float32) for _ in range(3))
    abi.load_host().vxh_sun_direction(sun_tick, _p(s), _p(m), _p(st))
    return s, m, st


def gen_world(kind: str, seed: int, dims=WORLD_DIMS, structures: bool = True) -> np.ndarray:
    """Deterministic stand-in worlds.  Returned array is indexed [z, y, x]."""
    nx, ny, nz = dims
    blocks = np.zeros((nz, ny, nx), dtype=np.uint8)
    lib = abi.load_host()
    if kind == "plains":
        lib.vxh_gen_plains(seed, int(structures), nx, ny, nz, _p(blocks))
    elif kind == "rooms":
        lib.vxh_gen_rooms(seed, nx, ny, nz, _p(blocks))
    elif kind == "town":
        lib.vxh_gen_town(seed, nx, ny, nz, _p(blocks))
    elif kind == "flat":
        blocks[:, :50, :] = 3
        blocks[:, 49, :] = 1
    elif kind == "empty":
        pass
    else:
        raise ValueError(f"unknown world kind {kind!r}")
    return blocks


def random_edits(blocks: np.ndarray, n: int, seed: int = 1234) -> np.ndarray:
    """Config-2 edit list; mutates `blocks` in place and returns the n x 4 int32 {x,y,z,id} list."""
    nz, ny, nx = blocks.shape
    e = np.zeros((n, 4), dtype=np.int32)
    abi.load_host().vxh_random_edits(seed, n, nx, ny, nz, _p(blocks), _p(e))
    return e


def load_named_world(name: str, fallback_kind: str, seed: int, dims=WORLD_DIMS):
    """$VXRT_WORLDS/<name> (headerless nx*ny*nz dump, Core/WorldFileHandler.cpp:26,50) if present,
    else the seeded stand-in.  Returns (blocks[z,y,x], description)."""
    nx, ny, nz = dims
    root = os.environ.get("VXRT_WORLDS")
    if root:
        path = Path(root) / name
        if path.exists() and path.stat().st_size == nx * ny * nz:
            blocks = np.zeros((nz, ny, nx), dtype=np.uint8)
            rc = abi.load_host().vxh_world_load(str(path).encode(), _p(blocks), nx * ny * nz)
            if rc == 0:
                return blocks, f"file:{name}"
    return gen_world(fallback_kind, seed, dims), f"stand-in:{fallback_kind}(seed={seed})"


def save_world(path: str, blocks: np.ndarray) -> None:
    rc = abi.load_host().vxh_world_save(str(path).encode(), _p(np.ascontiguousarray(blocks)), blocks.size)
    if rc != 0:
        raise OSError(f"could not write {path}")
