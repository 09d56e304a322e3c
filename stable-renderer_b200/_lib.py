"""ctypes binding of csrc/libsrx.so (the C ABI declared in include/srx.h).

There is no fallback: if the library is missing or cannot be loaded every call raises `SrxUnavailable`."""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libsrx.so")

# ---- enums (include/srx.h) ---------------------------------------------------------------------------------
SRX_OK, SRX_ERR_INVALID, SRX_ERR_INDEX, SRX_ERR_CUDA, SRX_ERR_UNSUPPORTED, SRX_ERR_KEY_RANGE = 0, -1, -2, -3, -4, -5
SRX_ERR_PEER_LOST = -6
SRX_F32, SRX_F16, SRX_BF16, SRX_I32, SRX_I16, SRX_U8 = 0, 1, 2, 10, 11, 20
SRX_KEY_VERTEX, SRX_KEY_TUPLE = 0, 1
SRX_STRATEGY = {"average": 0, "frame_distance": 1, "pixel_distance": 2, "perpendicular_view_normal": 3}
SRX_ACCUM_FAST, SRX_ACCUM_DETERMINISTIC, SRX_ACCUM_FAST_SPLIT = 0, 1, 2
SRX_BAKE_MODE = {"replace": 0, "replace_avg": 1, "first": 2, "first_avg": 3}
SRX_BAKE_WEIGHT = {None: 0, "none": 0, "uniform": 1, "view_normal": 2, "view_normal_depth": 3}


class SrxUnavailable(RuntimeError):
    pass


class SrxError(RuntimeError):
    pass


class srx_plan_desc(C.Structure):
    _fields_ = [("id_dtype", C.c_int), ("frames", C.c_int), ("height", C.c_int), ("width", C.c_int),
                ("batch", C.c_int), ("channels", C.c_int), ("lat_h", C.c_int), ("lat_w", C.c_int),
                ("key_mode", C.c_int), ("merge_len", C.c_int), ("accum_mode", C.c_int),
                ("key_capacity", C.c_int64), ("frame_map", C.POINTER(C.c_int32))]


class srx_plan_info(C.Structure):
    _fields_ = [("n_valid", C.c_int64), ("key_min", C.c_int64), ("key_max", C.c_int64), ("key_capacity", C.c_int64),
                ("workspace_bytes", C.c_int64), ("accum_offset", C.c_int64), ("accum_bytes", C.c_int64),
                ("accum_dtype", C.c_int), ("fast_path", C.c_int), ("fused", C.c_int), ("need_offset", C.c_int64)]


class srx_step_args(C.Structure):
    _fields_ = [("x_dev", C.c_void_p), ("x_dtype", C.c_int), ("ids_dev", C.c_void_p), ("ratio", C.c_float),
                ("adain", C.c_int), ("cache_slots", C.c_int)]


class srx_legacy_desc(C.Structure):
    _fields_ = [("id_dtype", C.c_int), ("frames", C.c_int), ("height", C.c_int), ("width", C.c_int),
                ("channels", C.c_int), ("lat_h", C.c_int), ("lat_w", C.c_int), ("merge_len", C.c_int),
                ("strategy", C.c_int)]


class srx_legacy_args(C.Structure):
    _fields_ = [("x_dev", C.c_void_p), ("x_dtype", C.c_int), ("ids_dev", C.c_void_p), ("alpha", C.c_float),
                ("view_normal_dev", C.c_void_p), ("workspace_dev", C.c_void_p), ("workspace_bytes", C.c_int64),
                ("defer_status", C.c_int)]


class srx_noise_args(C.Structure):
    _fields_ = [("ids_dev", C.c_void_p), ("id_dtype", C.c_int), ("frames", C.c_int), ("height", C.c_int), ("width", C.c_int),
                ("inv_frame_dev", C.c_void_p), ("rank_table_dev", C.c_void_p), ("key_latent_dev", C.c_void_p),
                ("key_noise_dev", C.c_void_p), ("base_latent_dev", C.c_void_p), ("base_noise_dev", C.c_void_p),
                ("latent_out_dev", C.c_void_p), ("noise_out_dev", C.c_void_p), ("mode", C.c_int), ("work_size", C.c_int),
                ("prev_frame_dev", C.c_void_p)]


class srx_bake_args(C.Structure):
    _fields_ = [("values_dev", C.c_void_p), ("writtens_dev", C.c_void_p), ("k2", C.c_int), ("texels", C.c_int),
                ("channels", C.c_int), ("colors_dev", C.c_void_p), ("color_dtype", C.c_int), ("color_channels", C.c_int),
                ("ids_dev", C.c_void_p), ("id_dtype", C.c_int), ("masks_dev", C.c_void_p), ("inverse_masks", C.c_int),
                ("frames", C.c_int), ("height", C.c_int), ("width", C.c_int), ("sprite_id", C.c_int),
                ("material_id", C.c_int), ("ignore_obj_mat_id", C.c_int), ("mode", C.c_int), ("weight_mode", C.c_int),
                ("normal_depth_dev", C.c_void_p), ("workspace_dev", C.c_void_p), ("workspace_bytes", C.c_int64), ("phase", C.c_int),
                ("frame_offset", C.c_int), ("frames_global", C.c_int), ("defer_status", C.c_int)]


class srx_gbuffer(C.Structure):
    _fields_ = [("color", C.c_void_p), ("ids", C.c_void_p), ("pos", C.c_void_p), ("normal_depth", C.c_void_p),
                ("noise", C.c_void_p), ("canny", C.c_void_p), ("canny_dtype", C.c_int)]


class srx_ingest_args(C.Structure):
    _fields_ = [("src", srx_gbuffer), ("height", C.c_int), ("width", C.c_int), ("flip_rows", C.c_int), ("frame_slot", C.c_int64),
                ("bg_noise", C.c_void_p), ("color_maps", C.c_void_p), ("masks", C.c_void_p), ("id_maps", C.c_void_p),
                ("pos_maps", C.c_void_p), ("normal_maps", C.c_void_p), ("depth_maps", C.c_void_p), ("canny_maps", C.c_void_p),
                ("noise_maps", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64)]


class srx_feature_args(C.Structure):
    _fields_ = [("ids_dev", C.c_void_p), ("id_dtype", C.c_int), ("frames", C.c_int), ("height", C.c_int), ("width", C.c_int),
                ("frame_map_dev", C.c_void_p), ("feat_dev", C.c_void_p), ("out_dev", C.c_void_p), ("x_dtype", C.c_int),
                ("batch", C.c_int), ("lat_h", C.c_int), ("lat_w", C.c_int), ("channels", C.c_int), ("map_height", C.c_int),
                ("map_width", C.c_int), ("ratio", C.c_float), ("key_capacity", C.c_int64), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_int64), ("reuse_buckets", C.c_int)]


class srx_gbuffer_arrays(C.Structure):
    _fields_ = [("color", C.c_void_p), ("ids", C.c_void_p), ("pos", C.c_void_p), ("normal_depth", C.c_void_p),
                ("noise", C.c_void_p), ("canny", C.c_void_p), ("canny_dtype", C.c_int)]


class srx_gbuffer_temp(C.Structure):
    _fields_ = [("color", C.c_void_p), ("ids", C.c_void_p), ("pos", C.c_void_p), ("normal", C.c_void_p), ("depth", C.c_void_p),
                ("noise", C.c_void_p), ("canny", C.c_void_p)]


# name -> (restype, argtypes); every symbol of include/srx.h is listed (tests check the export table against it)
_PROTOTYPES = {
    "srx_version": (C.c_int, []),
    "srx_last_error": (C.c_char_p, []),
    "srx_device_sm_count": (C.c_int, []),
    "srx_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(srx_plan_desc), C.c_void_p, C.c_void_p]),
    "srx_plan_get_info": (C.c_int, [C.c_void_p, C.POINTER(srx_plan_info)]),
    "srx_plan_bind_workspace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "srx_plan_bind_peers": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "srx_legacy_ordered_workspace_bytes": (C.c_int64, [C.POINTER(srx_legacy_desc)]),
    "srx_legacy_overlap_ordered": (C.c_int, [C.POINTER(srx_legacy_desc), C.POINTER(srx_legacy_args), C.c_int, C.c_void_p]),
    "srx_johnny_overlap": (C.c_int, [C.POINTER(srx_legacy_desc), C.POINTER(srx_legacy_args), C.c_float, C.c_void_p, C.c_void_p]),
    "srx_corrmap_keys_workspace_bytes": (C.c_int64, [C.c_int64]),
    "srx_corrmap_first_appearance": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_int64, C.c_void_p]),
    "srx_corrmap_drop_keys": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                        C.c_int64, C.c_void_p]),
    "srx_corrmap_trace_ranks_workspace_bytes": (C.c_int64, [C.c_int64]),
    "srx_corrmap_trace_ranks": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int64),
                                          C.c_void_p, C.c_int64, C.c_void_p]),
    "srx_corrmap_noise_fill": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "srx_bake_sharded_workspace_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "srx_atlas_quantize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "srx_atlas_dequantize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "srx_group_rank_workspace_ints": (C.c_int64, [C.c_int64]),
    "srx_group_rank": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
    "srx_ids_rank_table": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p,
                                     C.POINTER(C.c_int64), C.c_void_p]),
    "srx_group_broadcast": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "srx_noise_from_ids": (C.c_int, [C.POINTER(srx_noise_args), C.c_void_p]),
    "srx_plan_build_cache": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
    "srx_plan_cache_entries": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p]),
    "srx_plan_set_grid": (C.c_int, [C.c_void_p, C.c_int]),
    "srx_plan_bind_multicast": (C.c_int, [C.c_void_p, C.c_void_p]),
    "srx_feature_overlap_workspace_bytes": (C.c_int64, [C.POINTER(srx_feature_args)]),
    "srx_feature_overlap_bucket_bytes": (C.c_int64, [C.POINTER(srx_feature_args)]),
    "srx_feature_overlap": (C.c_int, [C.POINTER(srx_feature_args), C.c_void_p]),
    "srx_feature_overlap_rows": (C.c_int64, [C.POINTER(srx_feature_args), C.c_void_p]),
    "srx_feature_overlap_check": (C.c_int, [C.POINTER(srx_feature_args), C.c_void_p]),
    "srx_cells_overlap_workspace_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "srx_cells_overlap_keys": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p]),
    "srx_cells_overlap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "srx_plan_read_trace": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
    "srx_plan_read_step_ring": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]),
    "srx_plan_destroy": (C.c_int, [C.c_void_p]),
    "srx_plan_check": (C.c_int, [C.c_void_p, C.c_void_p]),
    "srx_vertex_screen_info": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32),
                                         C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
    "srx_overlap_step": (C.c_int, [C.c_void_p, C.POINTER(srx_step_args), C.c_void_p]),
    "srx_accum_reduce": (C.c_int, [C.c_void_p, C.POINTER(srx_step_args), C.c_void_p]),
    "srx_accum_finalize_gather": (C.c_int, [C.c_void_p, C.POINTER(srx_step_args), C.c_void_p]),
    "srx_group_by_then_average": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_int64, C.c_void_p]),
    "srx_legacy_workspace_bytes": (C.c_int64, [C.POINTER(srx_legacy_desc)]),
    "srx_legacy_overlap": (C.c_int, [C.POINTER(srx_legacy_desc), C.POINTER(srx_legacy_args), C.c_void_p]),
    "srx_legacy_check": (C.c_int, [C.POINTER(srx_legacy_desc), C.POINTER(srx_legacy_args), C.c_void_p]),
    "srx_bake_workspace_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "srx_bake_update": (C.c_int, [C.POINTER(srx_bake_args), C.c_void_p]),
    "srx_bake_check": (C.c_int, [C.POINTER(srx_bake_args), C.c_void_p]),
    "srx_gl_register_image": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint, C.c_uint, C.c_uint]),
    "srx_gl_map": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]),
    "srx_gl_unmap": (C.c_int, [C.c_void_p, C.c_void_p]),
    "srx_gl_unregister": (C.c_int, [C.c_void_p]),
    "srx_array_to_tensor": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "srx_tensor_to_array": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p]),
    "srx_array_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "srx_array_to_tensor_ch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "srx_tensor_to_array_ch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_uint, C.c_void_p]),
    "srx_atlas_to_array": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "srx_gl_mapped_layer": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "srx_array_free": (C.c_int, [C.c_void_p]),
    "srx_ingest_workspace_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "srx_frame_ingest": (C.c_int, [C.POINTER(srx_ingest_args), C.c_void_p]),
    "srx_gbuffer_merge_closer": (C.c_int, [C.POINTER(srx_gbuffer), C.c_int, C.c_int, C.c_int, C.POINTER(srx_gbuffer_temp), C.c_void_p]),
    "srx_frame_ingest_arrays": (C.c_int, [C.POINTER(srx_ingest_args), C.POINTER(srx_gbuffer_arrays), C.c_void_p]),
    "srx_gbuffer_merge_closer_arrays": (C.c_int, [C.POINTER(srx_gbuffer_arrays), C.c_int, C.c_int, C.c_int,
                                                  C.POINTER(srx_gbuffer_temp), C.c_void_p]),
}

_lock = threading.Lock()
_lib = None


def exported_symbols():
    return sorted(_PROTOTYPES)


def load() -> C.CDLL:
    """Loads libsrx.so (once).  Raises SrxUnavailable — never falls back to another implementation."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise SrxUnavailable(f"{LIB_PATH} is missing: build it with `python -m stable_renderer_b200.build` "
                                 f"(nvcc, sm_100a).  There is no CPU fallback.")
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise SrxUnavailable(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int):
    """Maps srx_status to the exception type the reference would raise for the same condition."""
    if rc == SRX_OK:
        return
    msg = load().srx_last_error().decode("utf-8", "replace")
    if rc == SRX_ERR_INDEX:
        raise IndexError(msg)
    if rc == SRX_ERR_INVALID:
        raise ValueError(msg)
    if rc == SRX_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise SrxError(f"[srx {rc}] {msg}")


def torch_dtype_code(dtype) -> int:
    import torch
    table = {torch.float32: SRX_F32, torch.float16: SRX_F16, torch.bfloat16: SRX_BF16, torch.int32: SRX_I32,
             torch.int16: SRX_I16, torch.uint8: SRX_U8}
    if dtype not in table:
        raise ValueError(f"unsupported dtype {dtype}")
    return table[dtype]


def current_stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream
