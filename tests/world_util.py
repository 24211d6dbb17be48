"""Shared inputs of the world-producer tests: a writer of synthetic Minecraft region files (our own, so the
fixture under tests/data/ is not reference data), random chunk-section batches, and the golden fixture."""
import struct
import sys
import zlib
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden" / "world_ref.npz"
SYNTH_DIR = ROOT / "tests" / "data" / "synth_region"
SYNTH_ORIGIN = (40.7, 3.2, -25.9)      # float on purpose: the importer truncates it (Importer.cpp:69)
GEN_SEEDS = [(4242, 999), (0, 49999)]  # (height-noise seed, biome seed)
MC_IDS = np.array([0, 0, 0, 1, 1, 2, 3, 4, 7, 12, 17, 18, 49, 89, 200, 255], dtype=np.uint8)


def mc_lut() -> np.ndarray:
    """A GetIDFromMCID table in the reference's shape: 0 -> 0, listed ids -> block, the rest -> INVALID_BLOCK (99)."""
    lut = np.full(256, 99, dtype=np.uint8)
    lut[0] = 0
    for mc, b in ((1, 3), (2, 1), (3, 2), (4, 4), (7, 60), (12, 5), (17, 6), (18, 7), (49, 41), (89, 12)):
        lut[mc] = b
    return lut


def emissive_table() -> np.ndarray:
    """BlockDataSSBO layout, emissive row: blocks 12 and 41 glow."""
    t = np.full((6, 128), -1, dtype=np.int32)
    t[0:3] = 0
    t[4:6] = 0
    t[3, 12] = 0
    t[3, 41] = 1
    return t


# ---- NBT / region writer ----
def _name(s: str) -> bytes:
    b = s.encode()
    return struct.pack(">H", len(b)) + b


def _tag(t: int, name: str, payload: bytes) -> bytes:
    return bytes([t]) + _name(name) + payload


def _byte_array(a: bytes) -> bytes:
    return struct.pack(">i", len(a)) + a


def _section(rng, y, with_y=True, with_data=True, with_blocks=True, with_add=False) -> bytes:
    out = b""
    if with_add:   # "Add" is ignored by the reader the engine uses (enkimi.c:2175-2179)
        out += _tag(7, "Add", _byte_array(bytes(rng.integers(0, 256, 2048, dtype=np.uint8))))
    if with_data:
        nib = rng.integers(0, 256, 2048, dtype=np.uint8)
        nib[rng.random(2048) < 0.7] = 0   # most voxels carry data value 0 and are imported
        out += _tag(7, "Data", _byte_array(bytes(nib)))
    if with_y:
        out += _tag(1, "Y", struct.pack("b", y))
    out += _tag(7, "SkyLight", _byte_array(bytes(2048)))
    if with_blocks:
        out += _tag(7, "Blocks", _byte_array(bytes(MC_IDS[rng.integers(0, len(MC_IDS), 4096)])))
    return out + b"\x00"


def _chunk_nbt(rng, cx, cz, variant) -> bytes:
    sections = []
    if variant == 0:      # plain: Y = 0..3
        sections = [_section(rng, y) for y in range(4)]
    elif variant == 1:    # no Y tags at all: indices count up from 0
        sections = [_section(rng, 0, with_y=False) for _ in range(3)]
    elif variant == 2:    # Y given once, then carried; one section without Data, one without Blocks
        sections = [_section(rng, 2), _section(rng, 0, with_y=False, with_data=False), _section(rng, 0, with_y=False, with_blocks=False),
                    _section(rng, 0, with_y=False, with_add=True)]
    elif variant == 3:    # duplicate Y: the later section replaces the earlier one; a negative and a high section
        sections = [_section(rng, 1), _section(rng, 1), _section(rng, -1), _section(rng, 7), _section(rng, 9)]
    elif variant == 4:    # empty list (element type End)
        sections = []
    body = b"".join(sections)
    sec_list = bytes([10 if sections else 0]) + struct.pack(">i", len(sections)) + body
    entity = _tag(8, "id", _name("Pig")) + _tag(9, "Pos", bytes([6]) + struct.pack(">i", 3) + struct.pack(">ddd", 1.0, 2.0, 3.0)) + b"\x00"
    level = (_tag(1, "LightPopulated", b"\x01") + _tag(3, "zPos", struct.pack(">i", cz)) + _tag(11, "HeightMap", struct.pack(">i", 4) + bytes(16)) +
             _tag(9, "Sections", sec_list) + _tag(9, "Entities", bytes([10]) + struct.pack(">i", 1) + entity) + _tag(4, "LastUpdate", struct.pack(">q", 77)) +
             _tag(3, "xPos", struct.pack(">i", cx)) + b"\x00")
    return b"\x0a" + _name("") + _tag(10, "Level", level) + b"\x00"


def write_region(path: Path, rng, rx: int, rz: int, n_chunks: int, cx_range, cz_range):
    header = bytearray(8192)
    body = bytearray()
    cand = np.array([cz * 32 + cx for cz in range(*cz_range) for cx in range(*cx_range)])
    slots = rng.choice(cand, n_chunks, replace=False)
    for k, slot in enumerate(sorted(int(s) for s in slots)):
        cx, cz = rx * 32 + slot % 32, rz * 32 + slot // 32
        comp = zlib.compress(_chunk_nbt(rng, cx, cz, k % 5), 6)
        rec = struct.pack(">IB", len(comp) + 1, 2) + comp
        rec += bytes(-len(rec) % 4096)
        sector = 2 + len(body) // 4096
        header[4 * slot:4 * slot + 4] = struct.pack(">I", (sector << 8) | (len(rec) // 4096))
        body += rec
    path.write_bytes(bytes(header) + bytes(body))


def write_synth_regions(directory: Path = SYNTH_DIR):
    """Two small region files around the world origin (seeded; 24 chunks, ~60 sections) plus a file the importer must
    ignore (wrong extension)."""
    directory.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(2024)
    write_region(directory / "r.0.0.mca", rng, 0, 0, 14, (0, 16), (0, 12))        # chunks partly beyond +x / +z
    write_region(directory / "r.-1.-1.mca", rng, -1, -1, 10, (20, 32), (16, 32))  # chunks partly beyond -x / -z
    (directory / "notes.txt").write_text("not a region file\n")


def random_sections(seed: int, n: int):
    """A batch in the layout of host_api.RegionSections with origins inside, straddling and outside the grid."""
    rng = np.random.default_rng(seed)
    ids = MC_IDS[rng.integers(0, len(MC_IDS), (n, 4096))]
    nib = rng.integers(0, 256, (n, 2048), dtype=np.uint8)
    nib[rng.random((n, 2048)) < 0.6] = 0
    has = (rng.random(n) < 0.8).astype(np.uint8)
    org = np.stack([rng.integers(-14, 26, n) * 16, rng.integers(-2, 10, n) * 16, rng.integers(-14, 26, n) * 16], axis=1).astype(np.int32)
    # keep one section per location so that the batch is order independent (region files guarantee this)
    _, first = np.unique(org, axis=0, return_index=True)
    keep = np.sort(first)
    return SimpleNamespace(block_ids=ids[keep], data_nibbles=nib[keep], has_data=has[keep], origins=org[keep])


def noise_points(n: int = 4096) -> np.ndarray:
    rng = np.random.default_rng(5)
    pts = (rng.random((n, 2), dtype=np.float32) * 900.0 - 250.0).astype(np.float32)
    pts[:8] = [[0, 0], [-1, -1], [-2, 0], [383, 383], [0.5, -0.5], [-3, -3], [1e-3, 2e-3], [100, -100]]
    return pts


def golden():
    return np.load(GOLD)
