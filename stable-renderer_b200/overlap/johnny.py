"""`johnny_overlap.overlap` — the experimental overlap variant of the legacy diffusers pipeline
(reference: legacy_codes/legacy_diffuser/modules/diffuser_pipelines/overlap/johnny_overlap.py:15-179).

Frame-distance weights `1 / (|t_i - t_j| + 1)`; every entry of a trace is rewritten before the next one is evaluated (the
reference loops over the entries and writes into the tensor it reads, :95-118); optionally each result is mixed with a base
colour — the noised ORIGINAL latent at the trace's first appearance (`beta`, :112-116).  The GPU kernel runs one warp per
trace (`csrc/srx_legacy_ordered.cu::k_johnny_sweep`).

As shipped the reference function cannot run: `beta = schedule(step, timestep, 'constant')` (:38) omits the required
`alpha_start` and raises TypeError, and `gamma` is hard-wired to 0 (:39).  Here `beta` is a keyword argument (default 0 = no
base-colour mix) and the extra-noise branch (`gamma > 0`, :139-143) does not exist."""
from __future__ import annotations

import ctypes as C
import math
from typing import Any, List, Literal, Optional

import torch

from .. import _lib
from .correspondence import CorrespondenceMap


def schedule(step: int, timestep: int, schedule_type: str, alpha_start: float, alpha_end: float = None, start_step: int = 0,
             end_step: int = 1000, every_step: int = 1, start_timestep: int = 1000, end_timestep: int = 0, **kwargs):
    """johnny_overlap.py:146-179 (host scalars)."""
    if step < start_step or step > end_step or step % every_step != 0 or timestep > start_timestep or timestep < end_timestep:
        return 0  # 0 means no overlap
    t = 1 - (timestep / 1000)
    if schedule_type == "constant":
        alpha = alpha_start
    elif schedule_type == "linear":
        alpha = alpha_start + (alpha_end - alpha_start) * t
    elif schedule_type == "cosine":
        alpha = alpha_start + (alpha_end - alpha_start) * (1 - math.cos(t * math.pi)) / 2
    elif schedule_type == "exponential":
        alpha = alpha_start * (alpha_end / alpha_start) ** t
    else:
        raise TypeError(f"Alpha schedule type `{schedule_type}` is not supported")
    return alpha


@torch.no_grad()
def overlap(frame_seq: List[torch.Tensor], corr_map: CorrespondenceMap, pipe: Any = None, interpolate_mode: str = 'nearest',
            weight_option: Literal['frame_distance', 'optical_flow'] = 'frame_distance', step: int = None, timestep: int = None,
            init_latents_orig_seq: Optional[List[torch.Tensor]] = None, noise_seq: Optional[List[torch.Tensor]] = None,
            beta: float = 0.0, **kwargs) -> List[torch.Tensor]:
    """frame_seq: T latents [B,C,h,w] -> T overlapped latents (new tensors)."""
    alpha = schedule(step, timestep, 'constant', 1)              # johnny_overlap.py:37
    beta = schedule(step, timestep, 'constant', beta)            # :38, with the missing argument supplied
    if alpha == 0:
        return frame_seq                                         # :44-45
    if interpolate_mode != 'nearest':
        raise NotImplementedError("only interpolate_mode='nearest' is supported")
    if weight_option != 'frame_distance':
        raise NotImplementedError("the reference implements 'frame_distance' only (johnny_overlap.py:100-104)")
    if getattr(corr_map, "merge_len", 0) > 1:
        raise NotImplementedError("johnny_overlap on a merged CorrespondenceMap: a merged trace keeps its sub-traces in dict order "
                                  "(correspondence_map.py:276-286) and the in-trace update is order dependent")
    if not frame_seq[0].is_cuda:
        raise _lib.SrxUnavailable("latents must be CUDA tensors (there is no CPU path)")
    lib = _lib.load()
    stack = torch.stack(frame_seq, dim=0).contiguous()           # [T,B,C,h,w]
    T, B, Cc, h, w = stack.shape
    dev = stack.device
    ids = corr_map.device_ids(dev)
    if ids.shape[0] < T:
        raise ValueError(f"correspondence map has {ids.shape[0]} frames, latents {T}")
    ids = ids[:T]
    base = None
    if beta > 0 and init_latents_orig_seq:
        # :63-65 — the noised original latents; the kernel samples them at the trace's first entry
        ts = torch.tensor([timestep])
        base = torch.stack([pipe.scheduler.add_noise(lat, noise, ts) for lat, noise in zip(init_latents_orig_seq, noise_seq)], dim=0)
        base = base.to(device=dev, dtype=torch.float32).reshape(T, B * Cc, h, w).contiguous()
    d = _lib.srx_legacy_desc()
    d.id_dtype = _lib.torch_dtype_code(ids.dtype)
    d.frames, d.height, d.width = T, ids.shape[1], ids.shape[2]
    d.channels, d.lat_h, d.lat_w = B * Cc, h, w
    d.merge_len = corr_map.merge_len
    d.strategy = _lib.SRX_STRATEGY["frame_distance"]
    need = int(lib.srx_legacy_ordered_workspace_bytes(C.byref(d)))
    if need < 0:
        _lib.check(_lib.SRX_ERR_INVALID)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    a = _lib.srx_legacy_args()
    a.x_dev, a.x_dtype = stack.data_ptr(), _lib.torch_dtype_code(stack.dtype)
    a.ids_dev, a.alpha = ids.data_ptr(), float(alpha)
    a.workspace_dev, a.workspace_bytes = ws.data_ptr(), ws.numel()
    with torch.cuda.device(dev):
        _lib.check(lib.srx_johnny_overlap(C.byref(d), C.byref(a), float(beta) if base is not None else 0.0,
                                          base.data_ptr() if base is not None else None, _lib.current_stream_ptr(dev)))
    return list(stack.unbind(0))


__all__ = ["overlap", "schedule"]
