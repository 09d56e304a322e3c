"""Wide-channel feature overlap: the body of `OverlapCorresponder.post_atten_inject`
(reference source/common_utils/stable_render_utils/corresponder.py:236-295 — unreachable there behind an early return) on the
GPU: `[B, h*w, c]` post-attention features, same-key features averaged across frames and blended back, AdaIN of the original
features to the result's statistics.  One sort-based bucketing pass per call (`csrc/srx_feature.cu`)."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch

from . import _lib
from .corrmap import IDMap

POST_ATTN_SKIP_LAYERS = tuple(range(11))     # corresponder.py:232-234: "post atten inject does not seem effective" there


def feature_overlap(origin_values: torch.Tensor, id_map: IDMap, ratio: float = 0.6,
                    map_size: Optional[Tuple[int, int]] = None, key_capacity: int = 0, check: bool = True,
                    cache_buckets: bool = True, info: Optional[dict] = None) -> torch.Tensor:
    """origin_values [B, h*w, c] (h*w a perfect square) -> new tensor of the same shape and dtype.

    map_size: the (height, width) the features are up-sampled to before the per-pixel gather.  The reference passes
              `(id_map.height, id_map.width)`, which for `[F,H,W,4]` ids are `(W, 4)` (corrmap.py:85-93) — the default here,
              to stay a drop-in; `(H, W)` is what the comment in the reference intends ("match the size of the id map").
    key_capacity: exclusive upper bound of the vertex ids; 0 = one `max()` over the ids (syncs, first call per id batch only).
    cache_buckets: keep the result of the bucketing pass (one id pass + one sort; depends on the ids, the frame indices and the sizes,
              not on the features) on the IDMap — every attention layer of that size and every denoise step of the run reuses it;
              `IDMap.invalidate()` drops it.  False = bucket on every call (nothing is kept).
    info: if a dict, receives `rows_gathered` and `buckets_reused` (one sync)."""
    if not origin_values.is_cuda:
        raise _lib.SrxUnavailable("features must live on a CUDA device (there is no CPU path)")
    if origin_values.dim() != 3:
        raise ValueError(f"features must be [B, h*w, c], got {tuple(origin_values.shape)}")
    B, hw, c = (int(v) for v in origin_values.shape)
    h = int(round(math.sqrt(hw)))
    if h * h != hw:
        raise ValueError(f"Dimension hw={hw} is not a perfect square.")
    dev = origin_values.device
    ids = id_map.device_ids(dev)
    F, H, W = int(ids.shape[0]), int(ids.shape[1]), int(ids.shape[2])
    mh, mw = (int(id_map.height), int(id_map.width)) if map_size is None else (int(map_size[0]), int(map_size[1]))
    frame_indices = tuple(int(v) for v in id_map.frame_indices)
    buckets = id_map._feature_buckets if cache_buckets else {}
    if key_capacity <= 0:
        key_capacity = buckets.get(("max_key", dev))
        if key_capacity is None:
            key_capacity = int(ids[..., 3].max().item()) + 1
            buckets[("max_key", dev)] = key_capacity
    bkey = (dev, ids.data_ptr(), frame_indices, B, h, mh, mw, int(key_capacity))
    feat = origin_values.contiguous()
    out = torch.empty_like(feat)
    lib = _lib.load()
    a = _lib.srx_feature_args()
    a.ids_dev, a.id_dtype, a.frames, a.height, a.width = ids.data_ptr(), _lib.torch_dtype_code(ids.dtype), F, H, W
    a.feat_dev, a.out_dev, a.x_dtype = feat.data_ptr(), out.data_ptr(), _lib.torch_dtype_code(feat.dtype)
    a.batch, a.lat_h, a.lat_w, a.channels = B, h, h, c
    a.map_height, a.map_width, a.ratio, a.key_capacity = mh, mw, float(ratio), int(key_capacity)
    ws = buckets.get(bkey)
    reuse = ws is not None
    if not reuse:
        nbytes = int(lib.srx_feature_overlap_workspace_bytes(C.byref(a)))
        if nbytes < 0:
            _lib.check(_lib.SRX_ERR_INVALID if c * feat.element_size() % 16 == 0 and c <= 1280 else _lib.SRX_ERR_UNSUPPORTED)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        fmap = torch.tensor(frame_indices, dtype=torch.int32, device=dev)
        a.frame_map_dev = fmap.data_ptr()
    elif c * feat.element_size() % 16 != 0 or c > 1280:
        _lib.check(_lib.SRX_ERR_UNSUPPORTED)
    a.workspace, a.workspace_bytes, a.reuse_buckets = ws.data_ptr(), int(ws.numel()), int(reuse)
    with torch.cuda.device(dev):
        stream = _lib.current_stream_ptr(dev)
        _lib.check(lib.srx_feature_overlap(C.byref(a), stream))
        if check and not reuse:
            _lib.check(lib.srx_feature_overlap_check(C.byref(a), stream))      # range failures belong to the bucketing pass
        if info is not None:
            info["rows_gathered"] = int(lib.srx_feature_overlap_rows(C.byref(a), stream))
            info["buckets_reused"] = reuse
        if cache_buckets and not reuse:
            # keep the buckets only: the scratch of the sort (two thirds of the workspace) goes back to the allocator
            keep = int(lib.srx_feature_overlap_bucket_bytes(C.byref(a)))
            kept = torch.empty(keep, dtype=torch.uint8, device=dev)
            kept.copy_(ws[:keep])
            buckets[bkey] = kept
    return out


def taichi_cells_overlap(id_flatten_maps: torch.Tensor, origin_values: torch.Tensor, new_values: torch.Tensor,
                         contributions: torch.Tensor) -> None:
    """`taichi_cells_overlap(id_flatten_maps, origin_values, new_values, contributions)` (reference corr_utils.py:110-134): fills
    `new_values` [b, cells, c] (float32, in place — its content is added to, like the reference's placeholder) with the
    similarity-weighted mean over all cells.  id_flatten_maps [b, pixels, 4] (any integer-valued dtype), contributions [b, pixels].
    Linear in pixels and distinct (cell, key) pairs instead of the reference's loop over all pixel pairs of all cell pairs."""
    if not (origin_values.is_cuda and new_values.is_cuda):
        raise _lib.SrxUnavailable("values must live on a CUDA device (there is no CPU path)")
    if new_values.dtype != torch.float32 or not new_values.is_contiguous() or new_values.shape != origin_values.shape:
        raise ValueError("new_values must be a contiguous float32 tensor of origin_values' shape")
    dev = new_values.device
    b, cells, c = (int(v) for v in origin_values.shape)
    pixels = int(id_flatten_maps.shape[1])
    ids = id_flatten_maps.to(device=dev, dtype=torch.int32).contiguous()
    contrib = contributions.to(device=dev, dtype=torch.float32).contiguous()
    vals = origin_values.to(torch.float32).contiguous()
    lib = _lib.load()
    nbytes = int(lib.srx_cells_overlap_workspace_bytes(b, pixels))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    n_keys = C.c_int64(0)
    with torch.cuda.device(dev):
        stream = _lib.current_stream_ptr(dev)
        _lib.check(lib.srx_cells_overlap_keys(ids.data_ptr(), contrib.data_ptr(), b, pixels, cells, ws.data_ptr(), nbytes, C.byref(n_keys), stream))
        sums = torch.zeros(max(int(n_keys.value), 1) * (c + 1), dtype=torch.float32, device=dev)
        _lib.check(lib.srx_cells_overlap(vals.data_ptr(), new_values.data_ptr(), b, pixels, cells, c, ws.data_ptr(), sums.data_ptr(),
                                         int(n_keys.value), stream))


__all__ = ["feature_overlap", "taichi_cells_overlap", "POST_ATTN_SKIP_LAYERS"]
