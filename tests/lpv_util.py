"""Shared inputs of the light-propagation-volume tests: seeded worlds with lamps, edit sequences, the golden fixture."""
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import world_util as wu  # noqa: E402
from voxeltracing_b200 import host_api  # noqa: E402

GOLD = ROOT / "tests" / "golden" / "lpv_ref.npz"
LAMPS = (12, 41)   # the two glowing blocks of world_util.emissive_table()


def lamp_world(seed: int, n_lamps: int, kind: str = "rooms") -> np.ndarray:
    """A stand-in world with `n_lamps` extra lamps: single ones, touching pairs (two seeds one voxel apart: the case where the order of
    the queue decides which lamp's block type a voxel records), lamps on the planes x / y / z = 0 (outside the volume for the flood
    fill) and on the far faces."""
    blocks = host_api.gen_world(kind, seed)
    nz, ny, nx = blocks.shape
    rng = np.random.default_rng(seed + 100)
    for i in range(n_lamps):
        x, y, z = int(rng.integers(0, nx)), int(rng.integers(0, ny)), int(rng.integers(0, nz))
        if i % 7 == 0:
            x = 0 if i % 14 == 0 else nx - 1
        if i % 11 == 0:
            y = 0 if i % 22 == 0 else ny - 1
        blocks[z, y, x] = LAMPS[i % 2]
        if i % 3 == 0 and x + 2 < nx:   # a second lamp of the other type one or two voxels along x
            blocks[z, y, x + 1 + (i % 2)] = LAMPS[(i + 1) % 2]
    return blocks


def edit_sequence(blocks: np.ndarray, table: np.ndarray, n: int, seed: int):
    """n block edits the way World::Raycast applies them: op 1 places `block` into an air voxel, op 0 breaks a solid voxel; positions
    strictly inside the grid (World.cpp:267-271).  Lamps are broken and placed next to lit voxels so that every branch of
    DepropogateVolume runs.  Returns [(op, (x, y, z), block, emissive)], applying nothing."""
    nz, ny, nx = blocks.shape
    rng = np.random.default_rng(seed)
    b = blocks.copy()
    lamps = np.argwhere((b == LAMPS[0]) | (b == LAMPS[1]))
    lamps = lamps[(lamps[:, 0] > 0) & (lamps[:, 1] > 0) & (lamps[:, 2] > 0)]
    out = []
    for i in range(n):
        z, y, x = (int(v) for v in lamps[rng.integers(0, len(lamps))])
        kind = i % 4
        if kind == 0 and b[z, y, x] != 0:           # break a lamp
            blk = int(b[z, y, x])
            out.append((0, (x, y, z), blk, bool(table[3, blk] >= 0)))
            b[z, y, x] = 0
            continue
        # a voxel within 5 of the lamp
        dx, dy, dz = (int(v) for v in rng.integers(-5, 6, 3))
        x, y, z = min(max(x + dx, 1), nx - 1), min(max(y + dy, 1), ny - 1), min(max(z + dz, 1), nz - 1)
        if b[z, y, x] == 0:
            blk = (3, LAMPS[0], 3, LAMPS[1])[kind]  # stone or a lamp
            out.append((1, (x, y, z), blk, bool(table[3, blk] >= 0)))
            b[z, y, x] = blk
        else:
            blk = int(b[z, y, x])
            out.append((0, (x, y, z), blk, bool(table[3, blk] >= 0)))
            b[z, y, x] = 0
    return out


def apply_edit(blocks: np.ndarray, edit) -> None:
    op, (x, y, z), blk, _ = edit
    blocks[z, y, x] = blk if op == 1 else 0


def sparse(vol: np.ndarray):
    idx = np.flatnonzero(vol).astype(np.int32)
    return idx, vol.reshape(-1)[idx]


def dense(idx: np.ndarray, val: np.ndarray, shape) -> np.ndarray:
    out = np.zeros(int(np.prod(shape)), dtype=np.uint8)
    out[idx] = val
    return out.reshape(shape)


def crc(level: np.ndarray, color: np.ndarray) -> np.ndarray:
    """(CRC-32 of the level volume, CRC-32 of the block-type volume, lit voxels)"""
    return np.array([zlib.crc32(np.ascontiguousarray(level).tobytes()), zlib.crc32(np.ascontiguousarray(color).tobytes()),
                     int(np.count_nonzero(level))], dtype=np.int64)


def sample_case(level: np.ndarray):
    """Seeded inputs of the SampleLPVData tests: a colour table, points in and around the lit voxels plus the corners / faces / outside of
    the volume, three dither vectors."""
    rng = np.random.default_rng(0)
    avg = np.zeros((128, 4), np.float32)
    avg[:, :3] = rng.random((128, 3)).astype(np.float32)
    lit = np.argwhere(level > 0)
    pts = (lit[rng.integers(0, len(lit), 20000)][:, ::-1] + rng.random((20000, 3)) * 1.5 - 0.25).astype(np.float32)
    pts[:50] = rng.random((50, 3)).astype(np.float32) * np.array([384, 128, 384], np.float32)
    pts[50:60] = [[0, 0, 0], [384, 128, 384], [-3, 5, 5], [400, 5, 5], [0.5, 0.5, 0.5], [383.5, 127.5, 383.5], [1, 1, 1], [383, 127, 383],
                  [192, 64, 192], [10.25, 20.5, 30.75]]
    dithers = np.array([[0, 0, 0], [0.5 / 384, 0.5 / 128, 0.5 / 384], [0.9 / 384, 0.1 / 128, 0.3 / 384]], np.float32)
    return avg, pts, dithers


def golden():
    return np.load(GOLD)
