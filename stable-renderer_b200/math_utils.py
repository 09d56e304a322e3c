"""GPU counterparts of the reference's `source/common_utils/math_utils.py` functions that sit on the hot path.

Same names, argument meaning and error behaviour; CUDA tensors only (no CPU fallback)."""
from __future__ import annotations

import ctypes as C
from typing import Literal

import torch

from . import _lib


def tensor_group_by_then_average(t: torch.Tensor, index_column: int, value_columns: list[int],
                                 return_unique: bool = False, key_capacity: int | None = None):
    """`tensor_group_by_then_average` (reference math_utils.py:86-161): group rows by `t[:, index_column]`, return the
    float32 mean of `value_columns` expanded back to every row (and the sorted unique keys when `return_unique`).

    Keys must be non-negative integers (vertex ids); `key_capacity` (default max key + 1, one sync) sizes the dense
    slot table that replaces the reference's `unique` sort."""
    if index_column >= t.shape[-1]:
        raise ValueError(f"Index column {index_column} is out of range.")
    if any(col >= t.shape[-1] for col in value_columns):
        raise ValueError(f"Value columns {value_columns} contain out of range values.")
    if not t.is_cuda:
        raise _lib.SrxUnavailable("tensor_group_by_then_average needs a CUDA tensor (there is no CPU path)")
    lib = _lib.load()
    keys = t[:, index_column].to(torch.float32).contiguous()
    values = t[:, value_columns].to(torch.float32).contiguous()
    n, c = values.shape
    if n == 0:
        out = torch.empty(0, c, dtype=torch.float32, device=t.device)
        return (out, keys.unique()) if return_unique else (out,)
    if key_capacity is None:
        key_capacity = int(keys.max().item()) + 1
    key_capacity = max(int(key_capacity), 1)
    out = torch.empty(n, c, dtype=torch.float32, device=t.device)
    ws = torch.empty(key_capacity * (c + 1), dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(lib.srx_group_by_then_average(values.data_ptr(), keys.data_ptr(), n, c, out.data_ptr(), ws.data_ptr(),
                                                 key_capacity, _lib.current_stream_ptr(t.device)))
    if return_unique:
        counts = ws[key_capacity * c:]
        unique_values = torch.nonzero(counts > 0).flatten().to(t.dtype)
        return out, unique_values
    return (out,)


def tensor_group_by_then_randn_init(t: torch.Tensor, index_column: int, value_columns: list[int],
                                    return_unique: bool = False, key_capacity: int | None = None,
                                    generator: torch.Generator | None = None):
    """`tensor_group_by_then_randn_init` (reference math_utils.py:164-229): rows that share `t[:, index_column]` receive
    the same random row; returns the expanded random values (and the sorted unique keys when `return_unique`).

    The reference sorts the keys (`unique(return_inverse=True)`), draws `torch.randn_like(unique.expand(-1, C), dtype=float)`
    and indexes it with the inverse.  Here the inverse comes from a dense rank table on the GPU (bit-identical to the
    sort's); the draw is the same torch call — one `[n_unique, C]` float32 normal draw on `t`'s device from `generator`
    (default: the device's global generator, as in the reference) — so the values match the reference's on the same device."""
    if index_column >= t.shape[-1]:
        raise ValueError(f"Index column {index_column} is out of range.")
    if any(col >= t.shape[-1] for col in value_columns):
        raise ValueError(f"Value columns {value_columns} contain out of range values.")
    if not t.is_cuda:
        raise _lib.SrxUnavailable("tensor_group_by_then_randn_init needs a CUDA tensor (there is no CPU path)")
    lib = _lib.load()
    keys = t[:, index_column].to(torch.float32).contiguous()
    n, c = int(keys.shape[0]), len(value_columns)
    if n == 0:
        out = torch.empty(0, c, dtype=torch.float32, device=t.device)
        return (out, keys.unique()) if return_unique else (out,)
    if key_capacity is None:
        key_capacity = int(keys.max().item()) + 1
    key_capacity = max(int(key_capacity), 1)
    table = torch.empty(int(lib.srx_group_rank_workspace_ints(key_capacity)), dtype=torch.int32, device=t.device)
    rank = torch.empty(n, dtype=torch.int32, device=t.device)
    n_unique = C.c_int64(0)
    with torch.cuda.device(t.device):
        stream = _lib.current_stream_ptr(t.device)
        _lib.check(lib.srx_group_rank(keys.data_ptr(), n, key_capacity, table.data_ptr(), rank.data_ptr(), C.byref(n_unique), stream))
        random_values = torch.randn(int(n_unique.value), c, dtype=torch.float32, device=t.device, generator=generator)
        out = torch.empty(n, c, dtype=torch.float32, device=t.device)
        _lib.check(lib.srx_group_broadcast(random_values.data_ptr(), rank.data_ptr(), n, c, out.data_ptr(), stream))
    if return_unique:
        unique_values = torch.nonzero(table[:key_capacity] >= 0).flatten().to(t.dtype)
        return out, unique_values
    return (out,)


def calc_map_mean_std(feat: torch.Tensor, eps: float = 1e-5):
    """`calc_map_mean_std` (reference math_utils.py:27-52): per (N, C) mean and sqrt(unbiased var + eps)."""
    assert feat.dim() == 4
    n, c = feat.shape[:2]
    flat = feat.reshape(n, c, -1).float()
    var = flat.var(dim=2) + eps
    return flat.mean(dim=2).view(n, c, 1, 1).to(feat.dtype), var.sqrt().view(n, c, 1, 1).to(feat.dtype)


def adaptive_instance_normalization(content_feat: torch.Tensor, style_feat: torch.Tensor, eps: float = 1e-5,
                                    mode: Literal["NCHW", "NHWC"] = "NCHW") -> torch.Tensor:
    """`adaptive_instance_normalization` (reference math_utils.py:55-80).  Stand-alone helper for callers outside the
    overlap step (inside the step AdaIN is fused into the finalize kernel)."""
    if mode == "NCHW":
        assert content_feat.shape[:2] == style_feat.shape[:2]
    elif mode == "NHWC":
        assert (content_feat.shape[0], content_feat.shape[3]) == (style_feat.shape[0], style_feat.shape[3])
        content_feat = content_feat.permute(0, 3, 1, 2)
        style_feat = style_feat.permute(0, 3, 1, 2)
    style_mean, style_std = calc_map_mean_std(style_feat, eps)
    content_mean, content_std = calc_map_mean_std(content_feat, eps)
    normalized = (content_feat - content_mean) / content_std
    return normalized * style_std + style_mean


__all__ = ["tensor_group_by_then_average", "tensor_group_by_then_randn_init", "calc_map_mean_std", "adaptive_instance_normalization"]
