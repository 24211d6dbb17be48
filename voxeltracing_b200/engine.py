"""Host-side mirror of the reference's pass interface over the vxrt_cuda_* C ABI.

Names follow the reference: World::Buffer / GenerateDistanceField (Core/World.h, Core/World.cpp:69-113),
the InitialTrace / ShadowTrace / DiffuseTrace / ReflectionTrace / GenerateGBuffer / ColorPass dispatches
of Core/Pipeline.cpp.  Everything here marshals arguments; all arithmetic happens in libvxrt_cuda.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi


class VxrtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vxrt_cuda error {code}: {msg}")
        self.code = code


_ATT_DTYPES = {
    abi.ATT_INITIAL_T: (np.float16, 1),
    abi.ATT_INITIAL_NORMAL: (np.uint8, 1),
    abi.ATT_INITIAL_BLOCK: (np.uint8, 1),
    abi.ATT_INITIAL_INVT: (np.float32, 1),
    abi.ATT_SHADOW: (np.uint8, 1),
    abi.ATT_SHADOW_TRANSVERSAL: (np.float16, 1),
    abi.ATT_GBUF_ALBEDO: (np.float16, 3),
    abi.ATT_GBUF_NORMAL: (np.float16, 3),
    abi.ATT_GBUF_PBR: (np.uint8, 4),
    abi.ATT_GBUF_TEXAO: (np.uint8, 1),
    abi.ATT_DIRECT: (np.float16, 3),
    abi.ATT_GI_SH: (np.float16, 4),
    abi.ATT_GI_COCG: (np.float16, 2),
    abi.ATT_GI_UTILITY: (np.float16, 1),
    abi.ATT_GI_AOSKY: (np.uint8, 2),
    abi.ATT_REFL_COLOR: (np.float16, 4),
    abi.ATT_REFL_HITDIST: (np.float16, 1),
    abi.ATT_REFL_EMISSIVE: (np.uint8, 1),
    abi.ATT_PREV_INITIAL_T: (np.float16, 1),
    abi.ATT_PREV_INITIAL_NORMAL: (np.uint8, 1),
    abi.ATT_PREV_INITIAL_BLOCK: (np.uint8, 1),
}
_ATT_DTYPES.update({abi.ATT_SHADOW_TEMPORAL_A: (np.uint8, 1), abi.ATT_SHADOW_TEMPORAL_A + 1: (np.float16, 1),
                    abi.ATT_SHADOW_TEMPORAL_B: (np.uint8, 1), abi.ATT_SHADOW_TEMPORAL_B + 1: (np.float16, 1),
                    abi.ATT_SHADOW_FILTERED: (np.uint8, 1)})
for _s in (abi.ATT_REFL_TEMPORAL_A, abi.ATT_REFL_TEMPORAL_B):
    _ATT_DTYPES.update({_s: (np.float16, 4), _s + 1: (np.float16, 1), _s + 2: (np.float16, 1)})
_ATT_DTYPES[abi.ATT_PREV_REFL_HITDIST] = (np.float16, 1)
_ATT_DTYPES.update({abi.ATT_REFL_DENOISED_A: (np.float16, 4), abi.ATT_REFL_DENOISED_B: (np.float16, 4)})
# SVGF image sets (four consecutive ids): SH, CoCg, utility RGB16F (temporal sets) / variance R16F, AO + sky
for _s in (abi.ATT_SVGF_TEMPORAL_A, abi.ATT_SVGF_TEMPORAL_B, abi.ATT_SVGF_VARIANCE, abi.ATT_SVGF_DENOISE_A, abi.ATT_SVGF_DENOISE_B, abi.ATT_SVGF_PRESPATIAL):
    _ATT_DTYPES.update({_s: (np.float16, 4), _s + 1: (np.float16, 2),
                        _s + 2: (np.float16, 3 if _s in (abi.ATT_SVGF_TEMPORAL_A, abi.ATT_SVGF_TEMPORAL_B) else 1), _s + 3: (np.uint8, 2)})


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _fill(dst, src):
    src = np.asarray(src, dtype=np.float32).ravel()
    for i, v in enumerate(src):
        dst[i] = float(v)


class _DeviceArray:
    """Minimal __cuda_array_interface__ carrier so torch.as_tensor() can alias an attachment."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2, "strides": None,
        }


class Context:
    """One GPU's resources: grids, tables and attachments (the reference's GL context)."""

    def __init__(self, device: int = 0, dims=None):
        self._lib = abi.load_cuda()
        self._h = C.c_void_p()
        d = None
        if dims is not None:
            d = (C.c_int32 * 3)(*dims)
        self.dims = tuple(dims) if dims is not None else (384, 128, 384)
        self.device = int(device)
        self._check(self._lib.vxrt_cuda_create(C.byref(self._h), device, d))

    # -- plumbing --
    def _check(self, rc: int):
        if rc != abi.VXRT_OK:
            raise VxrtError(rc, self._lib.vxrt_cuda_last_error().decode())

    def close(self):
        if self._h:
            self._lib.vxrt_cuda_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self._lib.vxrt_cuda_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def set_option(self, name: str, value: int):
        self._check(self._lib.vxrt_cuda_set_option(self._h, name.encode(), int(value)))

    def synchronize(self):
        self._check(self._lib.vxrt_cuda_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.vxrt_cuda_launch_count(self._h))

    @property
    def nvox(self) -> int:
        return self.dims[0] * self.dims[1] * self.dims[2]

    # -- world (Core/World.h) --
    def upload_world(self, blocks: np.ndarray):
        b = np.ascontiguousarray(blocks, dtype=np.uint8)
        if b.size != self.nvox:
            raise ValueError(f"world has {b.size} voxels, context expects {self.nvox}")
        self._check(self._lib.vxrt_cuda_upload_world(self._h, _p(b)))

    def download_world(self) -> np.ndarray:
        nx, ny, nz = self.dims
        out = np.empty((nz, ny, nx), dtype=np.uint8)
        self._check(self._lib.vxrt_cuda_download_world(self._h, _p(out)))
        return out

    def edit_blocks(self, xyz_id: np.ndarray):
        e = np.ascontiguousarray(xyz_id, dtype=np.int32).reshape(-1, 4)
        self._check(self._lib.vxrt_cuda_edit_blocks(self._h, _p(e), e.shape[0]))

    def generate_distance_field(self):
        self._check(self._lib.vxrt_cuda_generate_distance_field(self._h))

    def download_distance_field(self) -> np.ndarray:
        nx, ny, nz = self.dims
        out = np.empty((nz, ny, nx), dtype=np.uint8)
        self._check(self._lib.vxrt_cuda_download_distance_field(self._h, _p(out)))
        return out

    def upload_distance_field(self, df: np.ndarray):
        d = np.ascontiguousarray(df, dtype=np.uint8)
        if d.size != self.nvox:
            raise ValueError("distance field size mismatch")
        self._check(self._lib.vxrt_cuda_upload_distance_field(self._h, _p(d)))

    # -- z-slab sharded regeneration (multi-GPU; see voxeltracing_b200/sharding.py) --
    def df_slab_phase_a(self, slab: int, slab_z0):
        z = (C.c_int32 * len(slab_z0))(*[int(v) for v in slab_z0])
        self._check(self._lib.vxrt_cuda_df_slab_phase_a(self._h, slab, len(slab_z0) - 1, z))

    def df_slab_phase_b(self, slab: int, slab_z0, first_planes_ptr: int, last_planes_ptr: int):
        z = (C.c_int32 * len(slab_z0))(*[int(v) for v in slab_z0])
        self._check(self._lib.vxrt_cuda_df_slab_phase_b(self._h, slab, len(slab_z0) - 1, z, C.c_void_p(first_planes_ptr),
                                                        C.c_void_p(last_planes_ptr)))

    def df_commit(self):
        self._check(self._lib.vxrt_cuda_df_commit(self._h))

    def df_device_array(self):
        """Zero-copy [nz, ny, nx] uint8 view of the distance field for torch.as_tensor (NCCL all-gather of slabs).
        The view is NOT ordered against this context's kernels unless the consumer runs on the context's stream: put the
        context on the consumer's stream first (set_stream / sharding.bind_streams) or synchronize() before using it."""
        blocks, df = C.c_void_p(), C.c_void_p()
        self._check(self._lib.vxrt_cuda_grid_device(self._h, C.byref(blocks), C.byref(df)))
        nx, ny, nz = self.dims
        return _DeviceArray(df.value, (nz, ny, nx), "|u1")

    # -- tables --
    def set_block_data(self, table: np.ndarray):
        t = np.ascontiguousarray(table, dtype=np.int32)
        if t.size != 6 * 128:
            raise ValueError("block data table must be 6 x 128 int32")
        self._check(self._lib.vxrt_cuda_set_block_data(self._h, _p(t)))

    def set_blue_noise(self, data: np.ndarray):
        d = np.ascontiguousarray(data, dtype=np.int32)
        self._check(self._lib.vxrt_cuda_set_blue_noise(self._h, _p(d), d.size))

    def set_blue_noise_texture(self, rgba: np.ndarray):
        t = np.ascontiguousarray(rgba, dtype=np.uint8)
        h, w = t.shape[0], t.shape[1]
        self._check(self._lib.vxrt_cuda_set_blue_noise_texture(self._h, _p(t), w, h))

    def set_texture_array(self, kind: int, rgba: np.ndarray):
        """rgba: uint8[layers, size, size, 4], file row 0 first (TextureArray::CreateArray)."""
        t = np.ascontiguousarray(rgba, dtype=np.uint8)
        layers, h, w = t.shape[0], t.shape[1], t.shape[2]
        self._check(self._lib.vxrt_cuda_set_texture_array(self._h, kind, layers, w, h, _p(t)))

    def set_skymap(self, faces: np.ndarray):
        """faces: float32[6, res, res, 3] in +X,-X,+Y,-Y,+Z,-Z order."""
        f = np.ascontiguousarray(faces, dtype=np.float32)
        self._check(self._lib.vxrt_cuda_set_skymap(self._h, f.shape[1], _p(f)))

    # -- passes --
    def initial_trace(self, cam, width: int, height: int, render_distance: int = 350, jitter=None, tile=(0, 0),
                      alpha_test: bool = False, fov: float = 90.0):
        p = abi.PrimaryParams()
        _fill(p.inv_view, cam.inv_view)
        _fill(p.inv_projection, cam.inv_projection)
        p.width, p.height = width, height
        if jitter is not None:
            p.jitter[0], p.jitter[1] = float(jitter[0]), float(jitter[1])
            p.jitter_on = 1
        p.render_distance = render_distance
        p.alpha_test = int(alpha_test)
        p.fov = float(fov)
        abi.set_tile(p.tile, tile)
        self._check(self._lib.vxrt_cuda_initial_trace(self._h, C.byref(p)))
        return p

    def shadow_trace(self, cam, width: int, height: int, light_direction, frame: int = 0, halton=(0.0, 0.0),
                     soft: bool = True, max_iterations: int = 350, tile=(0, 0), alpha_test: bool = False, fov: float = 90.0):
        p = abi.ShadowParams()
        _fill(p.inv_view, cam.inv_view)
        _fill(p.inv_projection, cam.inv_projection)
        p.width, p.height = width, height
        _fill(p.light_direction, light_direction)
        p.current_frame = frame
        p.halton[0], p.halton[1] = float(halton[0]), float(halton[1])
        p.soft_shadows = int(soft)
        p.alpha_test = int(alpha_test)
        p.fov = float(fov)
        p.max_iterations = max_iterations
        abi.set_tile(p.tile, tile)
        self._check(self._lib.vxrt_cuda_shadow_trace(self._h, C.byref(p)))
        return p

    def generate_gbuffer(self, params: "abi.GBufferParams"):
        self._check(self._lib.vxrt_cuda_generate_gbuffer(self._h, C.byref(params)))

    def shade_direct(self, params: "abi.DirectParams"):
        self._check(self._lib.vxrt_cuda_shade_direct(self._h, C.byref(params)))

    def diffuse_trace(self, params: "abi.GIParams"):
        self._check(self._lib.vxrt_cuda_diffuse_trace(self._h, C.byref(params)))

    def reflection_trace(self, params: "abi.ReflectionParams"):
        self._check(self._lib.vxrt_cuda_reflection_trace(self._h, C.byref(params)))

    # -- SVGF chain of the diffuse GI (Core/Pipeline.cpp:2428-2700) --
    def svgf_temporal(self, params: "abi.SvgfTemporalParams"):
        self._check(self._lib.vxrt_cuda_svgf_temporal(self._h, C.byref(params)))

    def svgf_prespatial(self, params: "abi.SvgfPreSpatialParams"):
        self._check(self._lib.vxrt_cuda_svgf_prespatial(self._h, C.byref(params)))

    def svgf_variance(self, params: "abi.SvgfVarianceParams"):
        self._check(self._lib.vxrt_cuda_svgf_variance(self._h, C.byref(params)))

    def svgf_spatial(self, params: "abi.SvgfSpatialParams"):
        self._check(self._lib.vxrt_cuda_svgf_spatial(self._h, C.byref(params)))

    def svgf_end_frame(self):
        self._check(self._lib.vxrt_cuda_svgf_end_frame(self._h))

    def end_frame(self):
        self._check(self._lib.vxrt_cuda_end_frame(self._h))

    # -- sun-shadow denoiser (Core/Pipeline.cpp:2947-3044) --
    def shadow_temporal(self, params: "abi.ShadowTemporalParams"):
        self._check(self._lib.vxrt_cuda_shadow_temporal(self._h, C.byref(params)))

    def shadow_filter(self, params: "abi.ShadowFilterParams"):
        self._check(self._lib.vxrt_cuda_shadow_filter(self._h, C.byref(params)))

    # -- reflection temporal filter (Core/Pipeline.cpp:3316-3400) --
    def specular_temporal(self, params: "abi.SpecularTemporalParams"):
        self._check(self._lib.vxrt_cuda_specular_temporal(self._h, C.byref(params)))

    def reflection_denoise(self, params: "abi.ReflectionDenoiseParams"):
        self._check(self._lib.vxrt_cuda_reflection_denoise(self._h, C.byref(params)))

    def select_shadow(self, att: int):
        """The image the reflection / colour passes sample as the shadow texture (raw trace by default)."""
        self._check(self._lib.vxrt_cuda_select_shadow(self._h, att))

    def read_set(self, first: int, with_ao: bool = True) -> dict:
        d = {"sh": self.read_attachment(first), "cocg": self.read_attachment(first + 1), "x": self.read_attachment(first + 2)}
        if with_ao:
            d["aosky"] = self.read_attachment(first + 3)
        return d

    def write_set(self, first: int, d: dict):
        for k, name in enumerate(("sh", "cocg", "x", "aosky")):
            if name in d:
                self.write_attachment(first + k, d[name])

    # -- attachments --
    def write_attachment(self, att: int, data: np.ndarray):
        """glTexImage2D: (h, w[, channels]) array in the attachment's format."""
        dt, ch = _ATT_DTYPES[att]
        a = np.ascontiguousarray(data, dtype=dt)
        h, w = a.shape[:2]
        assert a.size == h * w * ch, (a.shape, ch)
        self._check(self._lib.vxrt_cuda_write_attachment(self._h, att, w, h, a.itemsize * ch, _p(a)))

    def attachment_info(self, att: int):
        ptr, w, h, bpp = C.c_void_p(), C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self._lib.vxrt_cuda_attachment_device(self._h, att, C.byref(ptr), C.byref(w), C.byref(h), C.byref(bpp)))
        return ptr.value, w.value, h.value, bpp.value

    def bind_attachment(self, att: int, dev_ptr: int | None, capacity: int = 0):
        """Use caller-owned device memory for an attachment (None restores context-owned storage)."""
        self._check(self._lib.vxrt_cuda_bind_attachment(self._h, att, C.c_void_p(dev_ptr or 0), capacity))

    def read_attachment(self, att: int, out: np.ndarray | None = None) -> np.ndarray:
        _, w, h, bpp = self.attachment_info(att)
        dt, ch = _ATT_DTYPES[att]
        shape = (h, w) if ch == 1 else (h, w, ch)
        if out is None:
            out = np.empty(shape, dtype=dt)
        assert out.nbytes == w * h * bpp
        self._check(self._lib.vxrt_cuda_read_attachment(self._h, att, _p(out), out.nbytes))
        return out

    def read_attachment_async(self, att: int, out: np.ndarray):
        """Queues the read-back behind the passes issued so far and returns; `out` (page-locked for a truly asynchronous
        copy) is valid after wait_reads()."""
        _, w, h, bpp = self.attachment_info(att)
        assert out.nbytes == w * h * bpp and out.flags["C_CONTIGUOUS"]
        self._check(self._lib.vxrt_cuda_read_attachment_async(self._h, att, _p(out), out.nbytes))

    def copy_attachment_rows_async(self, att: int, dst_ptr: int, row0: int = 0, rows: int = 0):
        """Queues a DMA copy of rows [row0, row0 + rows) of an attachment (rows == 0: all of it) to `dst_ptr`, the address of those
        rows in the destination: device memory here or on a peer GPU (shared_open), or page-locked host memory."""
        self._check(self._lib.vxrt_cuda_copy_attachment_rows_async(self._h, att, row0, rows, C.c_void_p(dst_ptr)))

    def copy_attachment_rect_async(self, att: int, dst_image_ptr: int, tile=(0, 0, 0, 0)):
        """The same for a rectangle tile = (row0, rows, col0, cols): `dst_image_ptr` is the address of pixel (0, 0) of the destination
        image (same geometry as the attachment); one strided DMA copy."""
        t = tuple(tile) + (0, 0) * (len(tuple(tile)) == 2)
        self._check(self._lib.vxrt_cuda_copy_attachment_rect_async(self._h, att, t[0], t[1], t[2], t[3], C.c_void_p(dst_image_ptr)))

    # -- multi-GPU export: a buffer of this GPU that other processes' GPUs write into over NVLink --
    def shared_alloc(self, nbytes: int):
        """-> (device pointer, 64-byte handle to hand to the other ranks)"""
        ptr, handle = C.c_void_p(), (C.c_uint8 * 64)()
        self._check(self._lib.vxrt_cuda_shared_alloc(self._h, nbytes, C.byref(ptr), handle))
        return ptr.value, bytes(handle)

    def shared_free(self, ptr: int):
        self._check(self._lib.vxrt_cuda_shared_free(self._h, C.c_void_p(ptr)))

    def shared_open(self, handle: bytes) -> int:
        ptr, h = C.c_void_p(), (C.c_uint8 * 64).from_buffer_copy(handle)
        self._check(self._lib.vxrt_cuda_shared_open(self._h, h, C.byref(ptr)))
        return ptr.value

    def shared_close(self, ptr: int):
        self._check(self._lib.vxrt_cuda_shared_close(self._h, C.c_void_p(ptr)))

    def join_passes(self):
        """With set_option("pass_overlap", 1): the context's stream waits (on the device) for the passes queued on the second lane.  Only needed
        before the caller's own work on that stream (timing events, torch kernels); every Context call joins by itself."""
        self._check(self._lib.vxrt_cuda_join_passes(self._h))

    def join_reads(self):
        """The context's stream waits on the device for the copies queued so far (see vxrt_cuda_join_reads)."""
        self._check(self._lib.vxrt_cuda_join_reads(self._h))

    def wait_reads(self):
        self._check(self._lib.vxrt_cuda_wait_reads(self._h))

    def attachment_as_device_array(self, att: int):
        """Zero-copy view for torch.as_tensor(..., device='cuda') (NCCL tile gathers).  Same stream contract as
        df_device_array: consumers on another stream must be ordered by the caller (set_stream / sharding.bind_streams)."""
        ptr, w, h, bpp = self.attachment_info(att)
        dt, ch = _ATT_DTYPES[att]
        shape = (h, w) if ch == 1 else (h, w, ch)
        return _DeviceArray(ptr, shape, np.dtype(dt).str)

    RAY_HIT_DTYPE = np.dtype([("t", "<f4"), ("normal", "<f4", 3), ("end", "<f4", 3), ("block", "<i4"), ("intersection", "<i4"),
                              ("iterations", "<i4")])

    def trace_rays(self, origins, directions, max_iterations: int = 350) -> np.ndarray:
        """VoxelTraversalDF over a batch of caller-supplied rays ((n,3) float32 each); structured array of hits."""
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 3)
        if o.shape != d.shape:
            raise ValueError("origins and directions must have the same shape")
        hits = np.zeros(len(o), dtype=self.RAY_HIT_DTYPE)
        self._check(self._lib.vxrt_cuda_trace_rays(self._h, _p(o), _p(d), len(o), int(max_iterations), _p(hits)))
        return hits

    def raycast_detect(self, positions, directions) -> np.ndarray:
        """World::RaycastDetect over (n,3) rays; int32 (n,8): hit voxel x, y, z, block, face normal xyz, found."""
        o = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 3)
        if o.shape != d.shape:
            raise ValueError("positions and directions must have the same shape")
        out = np.zeros((len(o), 8), dtype=np.int32)
        self._check(self._lib.vxrt_cuda_raycast_detect(self._h, _p(o), _p(d), len(o), _p(out)))
        return out

    # -- world producers --
    def generate_world(self, gen_type: int, noise_seed: int, biome_seed: int, grass: int = 1, dirt: int = 2, stone: int = 3, sand: int = 5):
        """VoxelRT::GenerateWorld without structures, in device memory (Core/WorldGenerator.cpp:208-313)."""
        p = abi.WorldGenParams(int(gen_type), int(noise_seed), int(biome_seed), int(grass), int(dirt), int(stone), int(sand))
        self._check(self._lib.vxrt_cuda_generate_world(self._h, p))

    def import_sections(self, sections, import_origin, lut, clear_first: bool = True):
        """MCWorldImporter::ImportWorld over the chunk sections of host_api.RegionSections (Core/NBT/Importer.cpp:85-166)."""
        ids = np.ascontiguousarray(sections.block_ids, dtype=np.uint8)
        nib = np.ascontiguousarray(sections.data_nibbles, dtype=np.uint8)
        has = np.ascontiguousarray(sections.has_data, dtype=np.uint8)
        org = np.ascontiguousarray(sections.origins, dtype=np.int32)
        o = np.ascontiguousarray(import_origin, dtype=np.int32).reshape(3)
        l = np.ascontiguousarray(lut, dtype=np.uint8).reshape(256)
        self._check(self._lib.vxrt_cuda_import_sections(self._h, _p(ids), _p(nib), _p(has), _p(org), len(has), _p(o), _p(l), int(bool(clear_first))))

    def collect_lights(self, capacity: int = 1 << 20) -> np.ndarray:
        """LightLocations of LoadWorld (Core/WorldFileHandler.cpp:53-69): int32 (n,3) voxel coordinates in scan order."""
        out = np.zeros((max(int(capacity), 1), 3), dtype=np.int32)
        n = C.c_int32(0)
        self._check(self._lib.vxrt_cuda_collect_lights(self._h, _p(out), int(capacity), C.byref(n)))
        if n.value > capacity:
            return self.collect_lights(n.value)
        return out[: n.value].copy()

    # -- light propagation volume (Core/VolumetricFloodFill.cpp) --
    def lpv_repropagate(self, lights=None, distance_limit: int = 4):
        """Start-up sequence Core/Pipeline.cpp:1602-1611 / World::RepropogateLPV_ (Core/World.cpp:554-572).  lights: (n,3) int32 voxel
        coordinates in queue order, or None for the LightLocations scan of the device grid."""
        if lights is None:
            self._check(self._lib.vxrt_cuda_lpv_repropagate(self._h, None, 0, int(distance_limit)))
            return
        l = np.ascontiguousarray(lights, dtype=np.int32).reshape(-1, 3)
        self._check(self._lib.vxrt_cuda_lpv_repropagate(self._h, _p(l) if len(l) else _p(np.zeros(3, dtype=np.int32)), len(l), int(distance_limit)))

    def lpv_edit(self, op: int, xyz, block: int, distance_limit: int = 4):
        """The light-volume half of one block edit (Core/World.cpp:273-333 place, :395-446 break, :482-485); the grid already holds the edit."""
        self._check(self._lib.vxrt_cuda_lpv_edit(self._h, int(op), int(xyz[0]), int(xyz[1]), int(xyz[2]), int(block), int(distance_limit)))

    def lpv_average_colors(self) -> np.ndarray:
        """BlockAverageColorData (PrecomputeAverageBlockColor.comp, VolumetricFloodFill.cpp:102-123): (128, 4) float32."""
        out = np.zeros((128, 4), dtype=np.float32)
        self._check(self._lib.vxrt_cuda_lpv_average_colors(self._h, _p(out)))
        return out

    def lpv_set_average_colors(self, rgba: np.ndarray):
        t = np.ascontiguousarray(rgba, dtype=np.float32)
        assert t.size == 512
        self._check(self._lib.vxrt_cuda_lpv_set_average_colors(self._h, _p(t)))

    def lpv_sample(self, points, dither=(0.0, 0.0, 0.0)) -> np.ndarray:
        """SampleLPVData (ReflectionTraceFrag.glsl:1516-1528) at (n, 3) points in voxel units: (n, 3) float32."""
        p = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dither, dtype=np.float32).reshape(3)
        out = np.zeros_like(p)
        self._check(self._lib.vxrt_cuda_lpv_sample(self._h, _p(p), len(p), _p(d), _p(out)))
        return out

    def lpv_download(self):
        """(level, block_type) volumes indexed [z, y, x]."""
        nx, ny, nz = self.dims
        level = np.zeros((nz, ny, nx), dtype=np.uint8)
        color = np.zeros_like(level)
        self._check(self._lib.vxrt_cuda_lpv_download(self._h, _p(level), _p(color)))
        return level, color

    def lpv_upload(self, level: np.ndarray, block_type: np.ndarray):
        l, b = np.ascontiguousarray(level, dtype=np.uint8), np.ascontiguousarray(block_type, dtype=np.uint8)
        assert l.size == self.dims[0] * self.dims[1] * self.dims[2] and b.size == l.size
        self._check(self._lib.vxrt_cuda_lpv_upload(self._h, _p(l), _p(b)))

    # -- statistics --
    def stats_enable(self, on: bool):
        self._check(self._lib.vxrt_cuda_stats_enable(self._h, int(on)))

    def gather_peak(self, rounds: int = 256) -> float:
        """Measured 32-byte sectors / s for independent random 1-byte loads of the L2-resident distance field."""
        r = C.c_double(0.0)
        self._check(self._lib.vxrt_cuda_gather_peak(self._h, int(rounds), C.byref(r)))
        return r.value

    def probe_read(self, reset: bool = True) -> dict:
        """Summed CUDA-event duration / launch count / traversal statistics of the probed kernel (set_option('probe', 1))."""
        ms, n, s = C.c_double(0.0), C.c_int64(0), abi.TraceStats()
        self._check(self._lib.vxrt_cuda_probe_read(self._h, C.byref(ms), C.byref(n), C.byref(s), int(reset)))
        return {"ms": ms.value, "launches": n.value, "rays": s.rays, "iterations": s.iterations}

    def stats_read(self, reset: bool = True) -> dict:
        s = abi.TraceStats()
        self._check(self._lib.vxrt_cuda_stats_read(self._h, C.byref(s), int(reset)))
        return {"rays": s.rays, "iterations": s.iterations, "dda_steps": s.dda_steps, "hits": s.hits}
