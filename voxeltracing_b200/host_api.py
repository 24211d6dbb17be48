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


class BlockDatabase:
    """blockdb.txt -> block ids, texture-array layers and the BlockData table
    (Core/BlockDatabaseParser.cpp, Core/BlockDatabase.cpp, Core/BlockDataSSBO.cpp)."""

    KINDS = ("albedo", "normal", "pbr", "emissive")

    def __init__(self, path):
        self._lib = abi.load_host()
        self._h = self._lib.vxh_blockdb_parse(str(path).encode())
        if not self._h:
            raise OSError(f"could not open block database {path}")

    def close(self):
        if self._h:
            self._lib.vxh_blockdb_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def block_count(self) -> int:
        return self._lib.vxh_blockdb_block_count(self._h)

    def block_id(self, name: str) -> int:
        return self._lib.vxh_blockdb_block_id(self._h, name.encode())

    def block_name(self, block_id: int) -> str:
        return self._lib.vxh_blockdb_block_name(self._h, block_id).decode()

    def layer_paths(self, kind: int) -> list[str]:
        n = self._lib.vxh_blockdb_layer_count(self._h, kind)
        return [self._lib.vxh_blockdb_layer_path(self._h, kind, i).decode() for i in range(n)]

    def texture(self, kind: int, block_id: int, face: int = 0) -> int:
        return self._lib.vxh_blockdb_texture(self._h, kind, block_id, face)

    def table(self) -> np.ndarray:
        t = np.zeros((6, 128), dtype=np.int32)
        self._lib.vxh_blockdb_table(self._h, _p(t))
        return t

    def face_props(self, name: str) -> np.ndarray:
        o = np.zeros(10, dtype=np.int32)
        self._lib.vxh_blockdb_face_props(self._h, name.encode(), _p(o))
        return o

    def minecraft_lut(self) -> np.ndarray:
        o = np.zeros(256, dtype=np.uint8)
        self._lib.vxh_blockdb_minecraft_lut(self._h, _p(o))
        return o


class RegionSections:
    """Chunk sections of Minecraft Anvil region files, inflated on the host (vxrt_mca.cpp): the input of
    Context.import_sections.  block_ids (n,4096) YZX, data_nibbles (n,2048), has_data (n,), origins (n,3)."""

    def __init__(self, path):
        lib = abi.load_host()
        h = lib.vxh_mca_open()
        try:
            path = Path(path)
            rc = lib.vxh_mca_add_region_dir(h, str(path).encode()) if path.is_dir() else lib.vxh_mca_add_region_file(h, str(path).encode())
            if rc < 0:
                raise OSError(f"could not read region data at {path} (rc {rc})")
            n = lib.vxh_mca_section_count(h)
            self.chunks = lib.vxh_mca_chunk_count(h)
            self.palette_sections = lib.vxh_mca_palette_section_count(h)
            self.bad_chunks = lib.vxh_mca_bad_chunk_count(h)
            if n == 0 and self.palette_sections > 0:
                # Palette / BlockStates sections (Minecraft 1.13+) have no legacy ids for the engine's 8-bit MC-id table; an import
                # that would silently produce an empty grid is an error
                raise ValueError(f"{path}: {self.palette_sections} palette-format (1.13+) sections and no pre-flattening sections: "
                                 "only the legacy Blocks / Data chunk format is supported")

            def view(ptr, shape, dtype):
                if n == 0:
                    return np.zeros(shape, dtype=dtype)
                nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
                return np.frombuffer(C.string_at(ptr, nbytes), dtype=dtype).reshape(shape).copy()
            self.block_ids = view(lib.vxh_mca_block_ids(h), (n, 4096), np.uint8)
            self.data_nibbles = view(lib.vxh_mca_data_nibbles(h), (n, 2048), np.uint8)
            self.has_data = view(lib.vxh_mca_has_data(h), (n,), np.uint8)
            self.origins = view(lib.vxh_mca_section_origins(h), (n, 3), np.int32)
        finally:
            lib.vxh_mca_free(h)

    def __len__(self):
        return len(self.has_data)


def gen_texture_array(kind: int, layers: int, size: int = 512, seed: int = 7) -> np.ndarray:
    """Deterministic synthetic block textures [layers, size, size, 4] uint8 (the reference's PNGs do not
    travel to the GPU box; with the reference mounted, load the PNGs named by BlockDatabase.layer_paths)."""
    out = np.zeros((layers, size, size, 4), dtype=np.uint8)
    abi.load_host().vxh_gen_texture_array(seed, kind, layers, size, _p(out))
    return out


def constant_skymap(res: int = 16, rgb=(0.5, 0.7, 1.0)) -> np.ndarray:
    """Config-4 sky: constant-colour cube faces so the atmosphere model stays out of the loop."""
    sky = np.empty((6, res, res, 3), dtype=np.float32)
    sky[...] = np.asarray(rgb, dtype=np.float32)
    return sky


def gradient_skymap(res: int = 16) -> np.ndarray:
    """A simple analytic sky (horizon-to-zenith gradient + warm +X side) to exercise the cube-map lookup."""
    sky = np.empty((6, res, res, 3), dtype=np.float32)
    for f in range(6):
        for j in range(res):
            for i in range(res):
                s, t = (i + 0.5) / res * 2 - 1, (j + 0.5) / res * 2 - 1
                d = [(1, -t, -s), (-1, -t, s), (s, 1, t), (s, -1, -t), (s, -t, 1), (-s, -t, -1)][f]
                d = np.asarray(d, dtype=np.float64)
                d /= np.linalg.norm(d)
                up = max(d[1], 0.0)
                sky[f, j, i] = (0.55 - 0.3 * up + 0.25 * max(d[0], 0) ** 4, 0.7 - 0.15 * up, 0.95 + 0.3 * up)
    return sky
