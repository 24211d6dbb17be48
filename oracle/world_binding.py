"""World producers: ctypes view of the oracle restatement (vxrt_oracle_world.cpp), of the reference's own
code compiled in oracle/_ref/libvxrt_ref_world.so, and a pure-Python restatement of the region-file reader the reference
uses (enkiMI).  TEST INFRASTRUCTURE ONLY: imported by tests/ and the golden generator, never by the product."""
from __future__ import annotations

import ctypes as C
import struct
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from voxeltracing_b200 import abi  # noqa: E402  (struct layouts only)
from oracle import binding  # noqa: E402

REF_LIB = ROOT / "oracle" / "_ref" / "libvxrt_ref_world.so"
DIMS = (384, 128, 384)
_ref = None


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _oracle():
    L = binding.lib()
    vp, i32, f = C.c_void_p, C.c_int32, C.c_float
    L.vxo_fastnoise_2d.argtypes = [i32, i32, f, i32, vp, i32, vp]
    L.vxo_fastnoise_2d.restype = None
    L.vxo_generate_world.argtypes = [vp, i32, i32, i32, C.POINTER(abi.WorldGenParams)]
    L.vxo_generate_world.restype = None
    L.vxo_import_sections.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp, i32]
    L.vxo_import_sections.restype = None
    L.vxo_collect_lights.argtypes = [vp, i32, i32, i32, vp, vp, i32]
    L.vxo_collect_lights.restype = i32
    L.vxo_lpv_repropagate.argtypes = [vp, i32, i32, i32, vp, i32, i32, vp, vp]
    L.vxo_lpv_repropagate.restype = None
    L.vxo_lpv_edit.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    L.vxo_lpv_edit.restype = None
    L.vxo_lpv_sample.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32, vp, vp]
    L.vxo_lpv_sample.restype = None
    return L


# ---- oracle restatement ----
def fastnoise_2d(seed: int, fractal: bool, frequency: float, octaves: int, xy: np.ndarray) -> np.ndarray:
    xy = np.ascontiguousarray(xy, dtype=np.float32).reshape(-1, 2)
    out = np.zeros(len(xy), dtype=np.float32)
    _oracle().vxo_fastnoise_2d(seed, int(fractal), frequency, octaves, _p(xy), len(xy), _p(out))
    return out


def generate_world(gen_type: int, noise_seed: int, biome_seed: int, ids=(1, 2, 3, 5), dims=DIMS) -> np.ndarray:
    nx, ny, nz = dims
    out = np.zeros((nz, ny, nx), dtype=np.uint8)
    p = abi.WorldGenParams(int(gen_type), int(noise_seed), int(biome_seed), *[int(i) for i in ids])
    _oracle().vxo_generate_world(_p(out), nx, ny, nz, p)
    return out


def import_sections(sections, import_origin, lut, dims=DIMS, into: np.ndarray | None = None) -> np.ndarray:
    nx, ny, nz = dims
    out = np.zeros((nz, ny, nx), dtype=np.uint8) if into is None else into
    ids = np.ascontiguousarray(sections.block_ids, dtype=np.uint8)
    nib = np.ascontiguousarray(sections.data_nibbles, dtype=np.uint8)
    has = np.ascontiguousarray(sections.has_data, dtype=np.uint8)
    org = np.ascontiguousarray(sections.origins, dtype=np.int32)
    o = np.ascontiguousarray(import_origin, dtype=np.int32)
    l = np.ascontiguousarray(lut, dtype=np.uint8)
    _oracle().vxo_import_sections(_p(out), nx, ny, nz, _p(ids), _p(nib), _p(has), _p(org), len(has), _p(o), _p(l), int(into is None))
    return out


def collect_lights(blocks: np.ndarray, table: np.ndarray) -> np.ndarray:
    nz, ny, nx = blocks.shape
    b = np.ascontiguousarray(blocks)
    t = np.ascontiguousarray(table, dtype=np.int32)
    n = _oracle().vxo_collect_lights(_p(b), nx, ny, nz, _p(t), None, 0)
    out = np.zeros((max(n, 1), 3), dtype=np.int32)
    _oracle().vxo_collect_lights(_p(b), nx, ny, nz, _p(t), _p(out), n)
    return out[:n]


def lpv_repropagate(blocks: np.ndarray, lights: np.ndarray, limit: int = 4):
    """Light level and block-type volumes after clear + seed + propagate (vxrt_oracle_lpv.cpp)."""
    nz, ny, nx = blocks.shape
    b = np.ascontiguousarray(blocks)
    l = np.ascontiguousarray(lights, dtype=np.int32).reshape(-1, 3)
    level, color = np.zeros_like(b), np.zeros_like(b)
    _oracle().vxo_lpv_repropagate(_p(b), nx, ny, nz, _p(l), len(l), int(limit), _p(level), _p(color))
    return level, color


def lpv_edit(blocks_after: np.ndarray, op: int, xyz, block: int, emissive: bool, limit: int, level: np.ndarray, color: np.ndarray):
    """The LPV half of one block edit, in place on level / color."""
    nz, ny, nx = blocks_after.shape
    b = np.ascontiguousarray(blocks_after)
    assert level.flags.c_contiguous and color.flags.c_contiguous
    _oracle().vxo_lpv_edit(_p(b), nx, ny, nz, int(op), int(xyz[0]), int(xyz[1]), int(xyz[2]), int(block), int(emissive), int(limit), _p(level), _p(color))


def lpv_sample(level: np.ndarray, block_type: np.ndarray, avg: np.ndarray, points: np.ndarray, dither) -> np.ndarray:
    """SampleLPVData (ReflectionTraceFrag.glsl:1516-1528) at points given in voxel units: (n, 3) float32"""
    nz, ny, nx = level.shape
    l, b = np.ascontiguousarray(level, np.uint8), np.ascontiguousarray(block_type, np.uint8)
    a, p = np.ascontiguousarray(avg, np.float32).reshape(128, 4), np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(dither, np.float32).reshape(3)
    out = np.zeros_like(p)
    _oracle().vxo_lpv_sample(_p(l), _p(b), nx, ny, nz, _p(a), _p(p), len(p), _p(d), _p(out))
    return out


# ---- the reference's own code (oracle/_ref) ----
def ref_available() -> bool:
    return REF_LIB.exists()


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(str(REF_LIB))
        vp, i32, f = C.c_void_p, C.c_int32, C.c_float
        L.vxref_generate_world.argtypes = [vp, i32, i32, vp, vp]
        L.vxref_generate_world.restype = None
        L.vxref_fastnoise_2d.argtypes = [i32, i32, f, i32, vp, i32, vp]
        L.vxref_fastnoise_2d.restype = None
        L.vxref_import_world.argtypes = [C.c_char_p, vp, vp, vp]
        L.vxref_import_world.restype = i32
        if hasattr(L, "vxref_lpv_repropagate"):
            L.vxref_lpv_repropagate.argtypes = [vp, vp, i32, i32, i32, vp, vp]
            L.vxref_lpv_repropagate.restype = None
            L.vxref_lpv_edit.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp]
            L.vxref_lpv_edit.restype = None
        _ref = L
    return _ref


def ref_fastnoise_2d(seed: int, fractal: bool, frequency: float, octaves: int, xy: np.ndarray) -> np.ndarray:
    xy = np.ascontiguousarray(xy, dtype=np.float32).reshape(-1, 2)
    out = np.zeros(len(xy), dtype=np.float32)
    ref().vxref_fastnoise_2d(seed, int(fractal), frequency, octaves, _p(xy), len(xy), _p(out))
    return out


def ref_generate_world(gen_type: int, noise_seed: int, biome_seed: int, stone_seed: int = 7, structures: bool = False,
                       ids8=(1, 2, 3, 5, 6, 7, 8, 4)) -> np.ndarray:
    """VoxelRT::GenerateWorld as compiled from Core/WorldGenerator.cpp; the function draws its seeds in the order biome,
    height noise, stone (:217-219)."""
    out = np.zeros((DIMS[2], DIMS[1], DIMS[0]), dtype=np.uint8)
    seeds = np.array([biome_seed, noise_seed, stone_seed], dtype=np.int32)
    ids = np.array(ids8, dtype=np.int32)
    ref().vxref_generate_world(_p(out), int(gen_type), int(structures), _p(seeds), _p(ids))
    return out


def ref_import_world(directory, origin, lut) -> np.ndarray:
    out = np.zeros((DIMS[2], DIMS[1], DIMS[0]), dtype=np.uint8)
    o = np.ascontiguousarray(origin, dtype=np.float32)
    l = np.ascontiguousarray(lut, dtype=np.uint8)
    rc = ref().vxref_import_world(str(directory).encode(), _p(o), _p(l), _p(out))
    if rc != 0:
        raise OSError("the reference importer threw")
    return out


# ---- region files: pure-Python restatement of what the reference's reader (enkiMI) extracts ----
class PySections:
    """Chunk sections of region files read the way Dependencies/enkiMI/enkimi.c does for the pre-flattening chunk format:
    header entries :1503-1506, chunk framing :1946-1966, Level / xPos / zPos / Sections :2151-2243 (section index = signed
    byte, set by "Y", advanced after every section; the last section of an index wins; sections need "Blocks"), section
    origin :2282-2289.  Same fields as voxeltracing_b200.host_api.RegionSections."""

    def __init__(self, paths):
        ids, nib, has, org = [], [], [], []
        self.chunks = 0
        for path in paths:
            d = Path(path).read_bytes()
            if len(d) < 8192:
                continue
            for i in range(1024):
                e = d[4 * i:4 * i + 4]
                loc = ((e[0] << 16) + (e[1] << 8) + e[2]) * 4096
                if loc < 8192 or loc + 6 > len(d):
                    continue
                length = struct.unpack(">I", d[loc:loc + 4])[0] - 1
                if length < 0 or length + loc + 5 > len(d):
                    continue
                try:
                    raw = zlib.decompressobj(15 + 32).decompress(d[loc + 5:loc + 5 + length])
                except zlib.error:
                    continue
                chunk = self._chunk(raw)
                if chunk is None:
                    continue
                self.chunks += 1
                x, z, sections = chunk
                for index in sorted(sections):
                    blocks, data = sections[index]
                    ids.append(np.frombuffer(blocks, dtype=np.uint8))
                    nib.append(np.frombuffer(data, dtype=np.uint8) if data is not None else np.zeros(2048, dtype=np.uint8))
                    has.append(0 if data is None else 1)
                    org.append((x * 16, (index - 128) * 16, z * 16))
        n = len(has)
        self.block_ids = np.stack(ids) if n else np.zeros((0, 4096), dtype=np.uint8)
        self.data_nibbles = np.stack(nib) if n else np.zeros((0, 2048), dtype=np.uint8)
        self.has_data = np.array(has, dtype=np.uint8)
        self.origins = np.array(org, dtype=np.int32).reshape(n, 3)

    def __len__(self):
        return len(self.has_data)

    @staticmethod
    def _chunk(raw: bytes):
        pos = 0

        def u8():
            nonlocal pos
            pos += 1
            return raw[pos - 1]

        def be(fmt, size):
            nonlocal pos
            pos += size
            return struct.unpack(fmt, raw[pos - size:pos])[0]

        def name():
            nonlocal pos
            n = be(">H", 2)
            pos += n
            return raw[pos - n:pos]

        def skip(t):
            nonlocal pos
            if t in (1, 2, 3, 4, 5, 6):
                pos += (1, 2, 4, 8, 4, 8)[t - 1]
            elif t == 7:
                n = be(">i", 4)     # (`pos += be(...)` would read pos before be() advances it)
                pos += n
            elif t == 8:
                n = be(">H", 2)
                pos += n
            elif t == 9:
                et, n = u8(), be(">i", 4)
                for _ in range(n):
                    skip(et)
            elif t == 10:
                while True:
                    tt = u8()
                    if tt == 0:
                        return
                    name()
                    skip(tt)
            elif t == 11:
                n = be(">i", 4)
                pos += 4 * n
            elif t == 12:
                n = be(">i", 4)
                pos += 8 * n
            else:
                raise ValueError("bad tag")

        if u8() != 10:
            return None
        name()
        found = {}
        sections = {}
        while True:
            t = u8()
            if t == 0:
                break
            nm = name()
            if t == 10 and nm == b"Level":
                while True:
                    t2 = u8()
                    if t2 == 0:
                        break
                    n2 = name()
                    if t2 == 3 and n2 == b"xPos" and "x" not in found:
                        found["x"] = be(">i", 4)
                    elif t2 == 3 and n2 == b"zPos" and "z" not in found:
                        found["z"] = be(">i", 4)
                    elif t2 == 9 and n2 == b"Sections" and "s" not in found:
                        found["s"] = True
                        et, cnt = u8(), be(">i", 4)
                        if et != 10:
                            for _ in range(cnt):
                                skip(et)
                            continue
                        section_y = 0
                        for _ in range(cnt):
                            blocks = data = None
                            while True:
                                t3 = u8()
                                if t3 == 0:
                                    break
                                n3 = name()
                                if t3 == 7 and n3 == b"Blocks" and blocks is None:
                                    ln = be(">i", 4)
                                    if ln >= 4096:
                                        blocks = raw[pos:pos + 4096]
                                    pos += ln
                                elif t3 == 7 and n3 == b"Data" and data is None:
                                    ln = be(">i", 4)
                                    if ln >= 2048:
                                        data = raw[pos:pos + 2048]
                                    pos += ln
                                elif t3 == 1 and n3 == b"Y":
                                    section_y = struct.unpack("b", raw[pos:pos + 1])[0]
                                    pos += 1
                                else:
                                    skip(t3)
                            if blocks is not None:
                                sections[section_y + 128] = (blocks, data)
                            section_y = ((section_y + 1 + 128) % 256) - 128  # int8 wrap
                    else:
                        skip(t2)
            else:
                skip(t)
        if not ("x" in found and "z" in found and "s" in found):
            return None
        return found["x"], found["z"], sections


def ref_lpv_repropagate(blocks: np.ndarray, lights: np.ndarray, limit: int = 4, iterations: int = 3):
    """Core/VolumetricFloodFill.cpp as compiled, driven like Pipeline.cpp:1602-1611 / World::RepropogateLPV_ (384 x 128 x 384 only)."""
    assert blocks.shape == (DIMS[2], DIMS[1], DIMS[0])
    b = np.ascontiguousarray(blocks)
    l = np.ascontiguousarray(lights, dtype=np.int32).reshape(-1, 3)
    level, color = np.zeros_like(b), np.zeros_like(b)
    ref().vxref_lpv_repropagate(_p(b), _p(l), len(l), int(limit), int(iterations), _p(level), _p(color))
    return level, color


def ref_lpv_edit(blocks_after: np.ndarray, op: int, xyz, block: int, emissive: bool, limit: int, level: np.ndarray, color: np.ndarray):
    assert blocks_after.shape == (DIMS[2], DIMS[1], DIMS[0])
    b = np.ascontiguousarray(blocks_after)
    ref().vxref_lpv_edit(_p(b), int(op), int(xyz[0]), int(xyz[1]), int(xyz[2]), int(block), int(emissive), int(limit), _p(level), _p(color))


def ref_lpv_sample(level: np.ndarray, block_type: np.ndarray, avg: np.ndarray, points: np.ndarray, dither) -> np.ndarray:
    """SampleLPVData of the reference's ReflectionTraceFrag.glsl compiled in oracle/_ref/libvxrt_ref.so, called as a function"""
    from oracle import ref_binding as rb
    assert level.shape == (DIMS[2], DIMS[1], DIMS[0])
    l, b = np.ascontiguousarray(level, np.uint8), np.ascontiguousarray(block_type, np.uint8)
    a, p = np.ascontiguousarray(avg, np.float32).reshape(128, 4), np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(dither, np.float32).reshape(3)
    out = np.zeros_like(p)
    L = rb.lib()
    L.vxref_lpv_sample.argtypes = [C.c_void_p] * 4 + [C.c_int32, C.c_void_p, C.c_void_p]
    L.vxref_lpv_sample.restype = None
    L.vxref_lpv_sample(_p(l), _p(b), _p(a), _p(p), len(p), _p(d), _p(out))
    return out
