"""ctypes view of include/vxrt_cuda.h (the C ABI) and voxeltracing_b200/host/vxrt_host.h.

This is the reference-side binding stub a maintainer would write; it contains no
compute.  Loading fails loudly if the CUDA library has not been built: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

_PKG = Path(__file__).resolve().parent
# VXRT_CUDA_LIB selects another build of the same library (A/B experiments, tools/debug); never a fallback
CUDA_LIB_PATH = Path(os.environ["VXRT_CUDA_LIB"]) if os.environ.get("VXRT_CUDA_LIB") else _PKG / "libvxrt_cuda.so"
HOST_LIB_PATH = _PKG / "libvxrt_host.so"

VXRT_OK = 0

# vxrt_attachment
ATT_INITIAL_T, ATT_INITIAL_NORMAL, ATT_INITIAL_BLOCK, ATT_INITIAL_INVT = 0, 1, 2, 3
ATT_SHADOW, ATT_SHADOW_TRANSVERSAL = 4, 5
ATT_GBUF_ALBEDO, ATT_GBUF_NORMAL, ATT_GBUF_PBR, ATT_GBUF_TEXAO, ATT_DIRECT = 6, 7, 8, 9, 10
ATT_GI_SH, ATT_GI_COCG, ATT_GI_UTILITY, ATT_GI_AOSKY = 11, 12, 13, 14
ATT_REFL_COLOR, ATT_REFL_HITDIST, ATT_REFL_EMISSIVE = 15, 16, 17
# SVGF image sets: +0 SH, +1 CoCg, +2 utility (temporal) / variance, +3 AO/sky
ATT_SVGF_TEMPORAL_A, ATT_SVGF_TEMPORAL_B, ATT_SVGF_VARIANCE, ATT_SVGF_DENOISE_A, ATT_SVGF_DENOISE_B = 18, 22, 26, 30, 34
ATT_PREV_INITIAL_T, ATT_PREV_INITIAL_NORMAL, ATT_PREV_INITIAL_BLOCK = 38, 39, 40
# shadow denoiser: temporal sets are (shadow R8, accumulated frames R16F)
ATT_SHADOW_TEMPORAL_A, ATT_SHADOW_TEMPORAL_B, ATT_SHADOW_FILTERED = 41, 43, 45
# reflection temporal filter: temporal sets are (colour RGBA16F, accumulation factor R16F, stabilised hit distance R16F)
ATT_REFL_TEMPORAL_A, ATT_REFL_TEMPORAL_B, ATT_PREV_REFL_HITDIST = 46, 49, 52
ATT_REFL_DENOISED_A, ATT_REFL_DENOISED_B = 53, 54
ATT_SVGF_PRESPATIAL = 55

TEX_ALBEDO, TEX_NORMAL, TEX_PBR, TEX_EMISSIVE = 0, 1, 2, 3


class Tile(C.Structure):
    _fields_ = [("row0", C.c_int32), ("rows", C.c_int32), ("col0", C.c_int32), ("cols", C.c_int32)]


def set_tile(t: "Tile", tile) -> None:
    """tile = (row0, rows) or (row0, rows, col0, cols); zeros mean every row / every column (vxrt_tile)"""
    tile = tuple(tile) + (0, 0) * (len(tuple(tile)) == 2)
    t.row0, t.rows, t.col0, t.cols = (int(v) for v in tile)


class PrimaryParams(C.Structure):
    _fields_ = [
        ("inv_view", C.c_float * 16),
        ("inv_projection", C.c_float * 16),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("jitter", C.c_float * 2),
        ("jitter_on", C.c_int32),
        ("render_distance", C.c_int32),
        ("alpha_test", C.c_int32),
        ("tile", Tile),
        ("fov", C.c_float),
    ]


class ShadowParams(C.Structure):
    _fields_ = [
        ("inv_view", C.c_float * 16),
        ("inv_projection", C.c_float * 16),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("light_direction", C.c_float * 3),
        ("current_frame", C.c_int32),
        ("halton", C.c_float * 2),
        ("soft_shadows", C.c_int32),
        ("alpha_test", C.c_int32),
        ("max_iterations", C.c_int32),
        ("tile", Tile),
        ("fov", C.c_float),
    ]


class GBufferParams(C.Structure):
    _fields_ = [
        ("inv_view", C.c_float * 16), ("inv_projection", C.c_float * 16), ("width", C.c_int32), ("height", C.c_int32),
        ("grass_props", C.c_int32 * 10), ("cactus_props", C.c_int32 * 10), ("tile", Tile),
    ]


class DirectParams(C.Structure):
    _fields_ = [
        ("inv_view", C.c_float * 16), ("inv_projection", C.c_float * 16), ("width", C.c_int32), ("height", C.c_int32),
        ("viewer_position", C.c_float * 3), ("sun_direction", C.c_float * 3), ("moon_direction", C.c_float * 3),
        ("sun_color", C.c_float * 3), ("moon_color", C.c_float * 3), ("texture_desat_amount", C.c_float),
        ("amplify_normal_map", C.c_int32), ("tile", Tile),
    ]


class GIParams(C.Structure):
    _fields_ = [
        ("inv_view", C.c_float * 16), ("inv_projection", C.c_float * 16), ("width", C.c_int32), ("height", C.c_int32),
        ("spp", C.c_int32), ("checker_spp", C.c_int32), ("checkerboard", C.c_int32), ("trace_length", C.c_int32),
        ("shadow_trace_length", C.c_int32), ("current_frame", C.c_int32), ("current_frame_mod128", C.c_int32),
        ("use_blue_noise", C.c_int32), ("supersample", C.c_int32), ("halton", C.c_float * 2),
        ("sun_direction", C.c_float * 3), ("moon_direction", C.c_float * 3), ("sun_visibility", C.c_float),
        ("gi_sun_strength", C.c_float), ("gi_sky_strength", C.c_float), ("diffuse_light_intensity", C.c_float),
        ("viewer_position", C.c_float * 3), ("apply_player_shadow", C.c_int32), ("tile", Tile),
    ]


class ReflectionParams(C.Structure):
    _fields_ = [
        ("inv_view", C.c_float * 16), ("inv_projection", C.c_float * 16), ("view", C.c_float * 16), ("projection", C.c_float * 16),
        ("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("checkerboard", C.c_int32), ("trace_length", C.c_int32),
        ("shadow_trace_length", C.c_int32), ("current_frame", C.c_int32), ("current_frame_mod128", C.c_int32),
        ("use_blue_noise", C.c_int32), ("rough_reflections", C.c_int32), ("roughness_bias", C.c_int32), ("temporal", C.c_int32),
        ("reproject_to_screen_space", C.c_int32), ("derive_from_diffuse_sh", C.c_int32), ("halton", C.c_float * 2),
        ("sun_direction", C.c_float * 3), ("moon_direction", C.c_float * 3), ("stronger_light_direction", C.c_float * 3),
        ("viewer_position", C.c_float * 3), ("sun_strength_modifier", C.c_float), ("moon_strength_modifier", C.c_float),
        ("grass_props", C.c_int32 * 10), ("tile", Tile),
        ("lpv_gi", C.c_int32), ("use_decoupled_gi", C.c_int32), ("screen_space_skylighting_valid", C.c_int32),
    ]


class SvgfTemporalParams(C.Structure):
    _fields_ = [("inv_view", C.c_float * 16), ("inv_projection", C.c_float * 16), ("prev_view", C.c_float * 16),
                ("prev_projection", C.c_float * 16), ("width", C.c_int32), ("height", C.c_int32), ("in_set", C.c_int32),
                ("history_set", C.c_int32), ("out_set", C.c_int32), ("be_useful", C.c_int32), ("tile", Tile)]


class SvgfVarianceParams(C.Structure):
    _fields_ = [("inv_view", C.c_float * 16), ("inv_projection", C.c_float * 16), ("width", C.c_int32), ("height", C.c_int32),
                ("in_set", C.c_int32), ("do_spatial", C.c_int32), ("aggressive_disocclusion", C.c_int32), ("tile", Tile)]


class SvgfSpatialParams(C.Structure):
    _fields_ = [("inv_view", C.c_float * 16), ("inv_projection", C.c_float * 16), ("width", C.c_int32), ("height", C.c_int32),
                ("in_set", C.c_int32), ("ao_set", C.c_int32), ("temporal_set", C.c_int32), ("out_set", C.c_int32), ("step", C.c_int32),
                ("large_kernel", C.c_int32), ("do_spatial", C.c_int32), ("aggressive_disocclusion", C.c_int32),
                ("color_phi_bias", C.c_float), ("time", C.c_float), ("resolution_scale", C.c_float), ("tile", Tile)]


class SvgfPreSpatialParams(C.Structure):
    """vxrt_svgf_prespatial_params"""
    _fields_ = [("inv_view", C.c_float * 16), ("inv_projection", C.c_float * 16), ("width", C.c_int32), ("height", C.c_int32),
                ("in_set", C.c_int32), ("time", C.c_float), ("tile", Tile)]


class ShadowTemporalParams(C.Structure):
    _fields_ = [("inv_view", C.c_float * 16), ("inv_projection", C.c_float * 16), ("prev_view", C.c_float * 16),
                ("prev_projection", C.c_float * 16), ("width", C.c_int32), ("height", C.c_int32), ("history_set", C.c_int32),
                ("out_set", C.c_int32), ("shadow_temporal", C.c_int32), ("tile", Tile)]


class SpecularTemporalParams(C.Structure):
    """vxrt_specular_temporal_params"""
    _fields_ = [("inv_view", C.c_float * 16), ("inv_projection", C.c_float * 16), ("prev_view", C.c_float * 16),
                ("prev_projection", C.c_float * 16), ("current_camera_pos", C.c_float * 3), ("prev_camera_pos", C.c_float * 3),
                ("width", C.c_int32), ("height", C.c_int32), ("history_set", C.c_int32), ("out_set", C.c_int32),
                ("temporal_spec", C.c_int32), ("firefly_rejection", C.c_int32), ("aggressive_firefly_rejection", C.c_int32),
                ("smart_clip", C.c_int32), ("roughness_weight", C.c_int32), ("stabilize_hit_distance", C.c_int32), ("tile", Tile)]


class ReflectionDenoiseParams(C.Structure):
    """vxrt_reflection_denoise_params"""
    _fields_ = [("inv_view", C.c_float * 16), ("inv_projection", C.c_float * 16), ("view", C.c_float * 16), ("width", C.c_int32), ("height", C.c_int32),
                ("in_attachment", C.c_int32), ("out_attachment", C.c_int32), ("temporal_set", C.c_int32), ("hit_distance_attachment", C.c_int32),
                ("dir", C.c_int32), ("roughness_bias", C.c_int32), ("normal_map_aware", C.c_int32), ("handle_lobe_deviation", C.c_int32),
                ("derive_from_diffuse_sh", C.c_int32), ("amplify_transversal_weight", C.c_int32), ("temporal_weight", C.c_int32),
                ("radius_bias", C.c_int32), ("normal_map_weight_strength", C.c_float), ("denoiser_scale", C.c_float),
                ("resolution_scale", C.c_float), ("roughness_normal_weight_bias_strength", C.c_float), ("tile", Tile)]


class ShadowFilterParams(C.Structure):
    _fields_ = [("inv_view", C.c_float * 16), ("inv_projection", C.c_float * 16), ("width", C.c_int32), ("height", C.c_int32),
                ("in_set", C.c_int32), ("filter_scale", C.c_float), ("tile", Tile)]


class RayHit(C.Structure):
    _fields_ = [("t", C.c_float), ("normal", C.c_float * 3), ("end", C.c_float * 3), ("block", C.c_int32),
                ("intersection", C.c_int32), ("iterations", C.c_int32)]


class WorldGenParams(C.Structure):
    """vxrt_worldgen_params"""
    _fields_ = [("gen_type", C.c_int32), ("noise_seed", C.c_int32), ("biome_seed", C.c_int32), ("grass_id", C.c_int32),
                ("dirt_id", C.c_int32), ("stone_id", C.c_int32), ("sand_id", C.c_int32)]


class TraceStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("iterations", C.c_uint64), ("dda_steps", C.c_uint64), ("hits", C.c_uint64)]


def declared_symbols(header: Path) -> list[str]:
    """Names of every function a C header declares (used by the export-coverage test)."""
    import re

    text = header.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:vxrt_cuda|vxh|vxo)_[a-z0-9_]+)\s*\(", text)))


_cuda = None
_host = None


def load_cuda() -> C.CDLL:
    """dlopen the product library.  Raises if it was not built (no fallback path exists)."""
    global _cuda
    if _cuda is not None:
        return _cuda
    if not CUDA_LIB_PATH.exists():
        raise RuntimeError(
            f"{CUDA_LIB_PATH} is missing: build it with `python -m voxeltracing_b200.build` "
            "(the hot path is CUDA-only; there is no CPU fallback)"
        )
    lib = C.CDLL(str(CUDA_LIB_PATH))
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    P = C.POINTER
    sig = {
        "vxrt_cuda_create": (C.c_int, [P(vp), C.c_int, P(i32)]),
        "vxrt_cuda_destroy": (C.c_int, [vp]),
        "vxrt_cuda_last_error": (C.c_char_p, []),
        "vxrt_cuda_set_stream": (C.c_int, [vp, vp]),
        "vxrt_cuda_synchronize": (C.c_int, [vp]),
        "vxrt_cuda_set_option": (C.c_int, [vp, C.c_char_p, i32]),
        "vxrt_cuda_launch_count": (i64, [vp]),
        "vxrt_cuda_upload_world": (C.c_int, [vp, vp]),
        "vxrt_cuda_download_world": (C.c_int, [vp, vp]),
        "vxrt_cuda_edit_blocks": (C.c_int, [vp, vp, i32]),
        "vxrt_cuda_generate_distance_field": (C.c_int, [vp]),
        "vxrt_cuda_download_distance_field": (C.c_int, [vp, vp]),
        "vxrt_cuda_upload_distance_field": (C.c_int, [vp, vp]),
        "vxrt_cuda_df_slab_phase_a": (C.c_int, [vp, i32, i32, P(i32)]),
        "vxrt_cuda_df_slab_phase_b": (C.c_int, [vp, i32, i32, P(i32), vp, vp]),
        "vxrt_cuda_df_plane_device": (C.c_int, [vp, i32, P(vp)]),
        "vxrt_cuda_grid_device": (C.c_int, [vp, P(vp), P(vp)]),
        "vxrt_cuda_df_commit": (C.c_int, [vp]),
        "vxrt_cuda_set_block_data": (C.c_int, [vp, vp]),
        "vxrt_cuda_set_blue_noise": (C.c_int, [vp, vp, i32]),
        "vxrt_cuda_set_blue_noise_texture": (C.c_int, [vp, vp, i32, i32]),
        "vxrt_cuda_set_texture_array": (C.c_int, [vp, i32, i32, i32, i32, vp]),
        "vxrt_cuda_set_skymap": (C.c_int, [vp, i32, vp]),
        "vxrt_cuda_generate_gbuffer": (C.c_int, [vp, P(GBufferParams)]),
        "vxrt_cuda_shade_direct": (C.c_int, [vp, P(DirectParams)]),
        "vxrt_cuda_diffuse_trace": (C.c_int, [vp, P(GIParams)]),
        "vxrt_cuda_reflection_trace": (C.c_int, [vp, P(ReflectionParams)]),
        "vxrt_cuda_svgf_temporal": (C.c_int, [vp, P(SvgfTemporalParams)]),
        "vxrt_cuda_svgf_variance": (C.c_int, [vp, P(SvgfVarianceParams)]),
        "vxrt_cuda_svgf_spatial": (C.c_int, [vp, P(SvgfSpatialParams)]),
        "vxrt_cuda_svgf_prespatial": (C.c_int, [vp, P(SvgfPreSpatialParams)]),
        "vxrt_cuda_svgf_end_frame": (C.c_int, [vp]),
        "vxrt_cuda_end_frame": (C.c_int, [vp]),
        "vxrt_cuda_shadow_temporal": (C.c_int, [vp, P(ShadowTemporalParams)]),
        "vxrt_cuda_shadow_filter": (C.c_int, [vp, P(ShadowFilterParams)]),
        "vxrt_cuda_select_shadow": (C.c_int, [vp, i32]),
        "vxrt_cuda_specular_temporal": (C.c_int, [vp, P(SpecularTemporalParams)]),
        "vxrt_cuda_reflection_denoise": (C.c_int, [vp, P(ReflectionDenoiseParams)]),
        "vxrt_cuda_write_attachment": (C.c_int, [vp, i32, i32, i32, i32, vp]),
        "vxrt_cuda_read_attachment": (C.c_int, [vp, i32, vp, sz]),
        "vxrt_cuda_read_attachment_async": (C.c_int, [vp, i32, vp, sz]),
        "vxrt_cuda_wait_reads": (C.c_int, [vp]),
        "vxrt_cuda_join_reads": (C.c_int, [vp]),
        "vxrt_cuda_join_passes": (C.c_int, [vp]),
        "vxrt_cuda_copy_attachment_rows_async": (C.c_int, [vp, i32, i32, i32, vp]),
        "vxrt_cuda_copy_attachment_rect_async": (C.c_int, [vp, i32, i32, i32, i32, i32, vp]),
        "vxrt_cuda_shared_alloc": (C.c_int, [vp, sz, C.POINTER(vp), vp]),
        "vxrt_cuda_shared_free": (C.c_int, [vp, vp]),
        "vxrt_cuda_shared_open": (C.c_int, [vp, vp, C.POINTER(vp)]),
        "vxrt_cuda_shared_close": (C.c_int, [vp, vp]),
        "vxrt_cuda_bind_attachment": (C.c_int, [vp, i32, vp, sz]),
        "vxrt_cuda_attachment_device": (C.c_int, [vp, i32, P(vp), P(i32), P(i32), P(i32)]),
        "vxrt_cuda_initial_trace": (C.c_int, [vp, P(PrimaryParams)]),
        "vxrt_cuda_shadow_trace": (C.c_int, [vp, P(ShadowParams)]),
        "vxrt_cuda_trace_rays": (C.c_int, [vp, vp, vp, i32, i32, vp]),
        "vxrt_cuda_raycast_detect": (C.c_int, [vp, vp, vp, i32, vp]),
        "vxrt_cuda_generate_world": (C.c_int, [vp, P(WorldGenParams)]),
        "vxrt_cuda_import_sections": (C.c_int, [vp, vp, vp, vp, vp, i32, vp, vp, i32]),
        "vxrt_cuda_collect_lights": (C.c_int, [vp, vp, i32, P(i32)]),
        "vxrt_cuda_lpv_repropagate": (C.c_int, [vp, vp, i32, i32]),
        "vxrt_cuda_lpv_edit": (C.c_int, [vp, i32, i32, i32, i32, i32, i32]),
        "vxrt_cuda_lpv_average_colors": (C.c_int, [vp, vp]),
        "vxrt_cuda_lpv_set_average_colors": (C.c_int, [vp, vp]),
        "vxrt_cuda_lpv_sample": (C.c_int, [vp, vp, i32, vp, vp]),
        "vxrt_cuda_lpv_download": (C.c_int, [vp, vp, vp]),
        "vxrt_cuda_lpv_upload": (C.c_int, [vp, vp, vp]),
        "vxrt_cuda_stats_enable": (C.c_int, [vp, i32]),
        "vxrt_cuda_stats_read": (C.c_int, [vp, P(TraceStats), i32]),
        "vxrt_cuda_gather_peak": (C.c_int, [vp, i32, P(C.c_double)]),
        "vxrt_cuda_probe_read": (C.c_int, [vp, P(C.c_double), P(i64), P(TraceStats), i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _cuda = lib
    return lib


def load_host() -> C.CDLL:
    global _host
    if _host is not None:
        return _host
    if not HOST_LIB_PATH.exists():
        raise RuntimeError(f"{HOST_LIB_PATH} is missing: build it with `python -m voxeltracing_b200.build`")
    lib = C.CDLL(str(HOST_LIB_PATH))
    vp, i32, u32, f = C.c_void_p, C.c_int32, C.c_uint32, C.c_float
    sig = {
        "vxh_perspective": (None, [f, f, f, f, vp]),
        "vxh_look_at": (None, [vp, vp, vp, vp]),
        "vxh_inverse": (None, [vp, vp]),
        "vxh_camera": (None, [vp, f, f, f, f, vp, vp, vp, vp]),
        "vxh_taa_jitter": (None, [i32, vp]),
        "vxh_sun_direction": (None, [f, vp, vp, vp]),
        "vxh_gen_plains": (None, [u32, i32, i32, i32, i32, vp]),
        "vxh_gen_rooms": (None, [u32, i32, i32, i32, vp]),
        "vxh_gen_town": (None, [u32, i32, i32, i32, vp]),
        "vxh_random_edits": (None, [u32, i32, i32, i32, i32, vp, vp]),
        "vxh_world_save": (i32, [C.c_char_p, vp, C.c_int64]),
        "vxh_world_load": (i32, [C.c_char_p, vp, C.c_int64]),
        "vxh_blockdb_parse": (vp, [C.c_char_p]),
        "vxh_blockdb_free": (None, [vp]),
        "vxh_blockdb_block_count": (i32, [vp]),
        "vxh_blockdb_block_id": (i32, [vp, C.c_char_p]),
        "vxh_blockdb_block_name": (C.c_char_p, [vp, i32]),
        "vxh_blockdb_layer_count": (i32, [vp, i32]),
        "vxh_blockdb_layer_path": (C.c_char_p, [vp, i32, i32]),
        "vxh_blockdb_texture": (i32, [vp, i32, i32, i32]),
        "vxh_blockdb_table": (None, [vp, vp]),
        "vxh_blockdb_face_props": (None, [vp, C.c_char_p, vp]),
        "vxh_blockdb_minecraft_lut": (None, [vp, vp]),
        "vxh_gen_texture_array": (None, [u32, i32, i32, i32, vp]),
        "vxh_mca_open": (vp, []),
        "vxh_mca_free": (None, [vp]),
        "vxh_mca_add_region_file": (i32, [vp, C.c_char_p]),
        "vxh_mca_add_region_dir": (i32, [vp, C.c_char_p]),
        "vxh_mca_section_count": (i32, [vp]),
        "vxh_mca_chunk_count": (i32, [vp]),
        "vxh_mca_palette_section_count": (i32, [vp]),
        "vxh_mca_bad_chunk_count": (i32, [vp]),
        "vxh_mca_block_ids": (vp, [vp]),
        "vxh_mca_data_nibbles": (vp, [vp]),
        "vxh_mca_has_data": (vp, [vp]),
        "vxh_mca_section_origins": (vp, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _host = lib
    return lib
