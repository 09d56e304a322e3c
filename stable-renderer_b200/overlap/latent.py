"""GPU counterpart of the legacy `CorrMapLatentNoiseInitializer` node (reference: legacy_codes/nodes/latent.py:7-40):
latent and noise batches in which every pixel of a trace (a key of the `CorrespondenceMap` with at least two entries)
starts from the same random 4-vector.

The random numbers are the reference's own stream: both `torch.manual_seed` calls return the SAME default CPU generator
(latent.py:18-19), so after seeding with `seed + 1` the stream is base latent `[1,4,H,W]`, base noise `[1,4,H,W]`, then for
every trace in dict insertion order `randn(4)` for the latent and `randn(4)` for the noise (latent.py:28-35).  The
per-trace draws take torch's scalar sampling path; one `normal_()` on a strided view takes the same path, so all rows
come from a single call with bit-identical values (tests/test_latent_init.py checks this against the loop).
Which row belongs to which pixel (the dict's insertion order, singleton keys skipped) and the fill of exactly the pixels
the nearest down-sample keeps run in `csrc/srx_legacy.cu`; neither the dict nor the full-resolution batches exist."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from .correspondence import CorrespondenceMap


def trace_rows(n_traces: int) -> torch.Tensor:
    """`[n_traces, 2, 4]` rows from the default CPU generator: the values of `n_traces` x (`randn(4)`, `randn(4)`) calls."""
    if n_traces == 0:
        return torch.empty(0, 2, 4)
    buf = torch.empty(2 * 8 * n_traces)
    rows = buf[::2]                 # a non-contiguous view is sampled element by element, like the reference's randn(4)
    rows.normal_()
    return rows.contiguous().view(n_traces, 2, 4)


class CorrMapLatentNoiseInitializer:
    """Same call signature and return value (`({"samples", "noise"},)`) as the reference node."""

    Category = "Latent"

    @torch.no_grad()
    def __call__(self, width: int, height: int, batch_size: int, seed: int, correspondence_map: CorrespondenceMap):
        cm = correspondence_map
        if not torch.cuda.is_available():
            raise _lib.SrxUnavailable("CorrMapLatentNoiseInitializer runs on the GPU (there is no CPU path)")
        lib = _lib.load()
        ids = cm.ids if cm.ids.is_cuda else cm.device_ids("cuda")
        dev = ids.device
        F, H, W = cm.num_frames, cm.height, cm.width
        lat_h, lat_w = int(height) // 8, int(width) // 8
        if lat_h <= 0 or lat_w <= 0 or batch_size <= 0:
            raise ValueError("width, height and batch_size must be positive")
        npx = F * H * W
        rank = torch.empty(npx, dtype=torch.int32, device=dev)
        ws = torch.empty(int(lib.srx_corrmap_trace_ranks_workspace_bytes(npx)), dtype=torch.uint8, device=dev)
        n_traces = C.c_int64(0)
        with torch.cuda.device(dev):
            stream = _lib.current_stream_ptr(dev)
            _lib.check(lib.srx_corrmap_trace_ranks(ids.data_ptr(), _lib.torch_dtype_code(ids.dtype), F, H, W, cm.merge_len,
                                                   rank.data_ptr(), C.byref(n_traces), ws.data_ptr(), ws.numel(), stream))
            if batch_size < F and n_traces.value:
                raise IndexError(f"index {F - 1} is out of bounds for dimension 0 with size {batch_size}")
            torch.manual_seed(seed)
            torch.manual_seed(seed + 1)              # the generator both base draws and all rows come from (latent.py:18-19)
            base = torch.cat([torch.randn([1, 4, H, W], device="cpu"),       # latent base, then noise base: two draws from
                              torch.randn([1, 4, H, W], device="cpu")]).pin_memory()   # one stream (latent.py:22,25)
            rows = trace_rows(int(n_traces.value))
            base_d = base.to(dev, non_blocking=True)
            rows_d = rows.to(dev) if rows.numel() else torch.zeros(1, 2, 4, device=dev)
            latent = torch.empty(batch_size, 4, lat_h, lat_w, dtype=torch.float32, device=dev)
            noise = torch.empty_like(latent)
            frames = min(F, batch_size)              # batch_size < F is only reachable without traces: base values everywhere
            _lib.check(lib.srx_corrmap_noise_fill(rank.data_ptr(), frames, H, W, batch_size, lat_h, lat_w, base_d.data_ptr(),
                                                  rows_d.data_ptr(), latent.data_ptr(), noise.data_ptr(), stream))
        return ({"samples": latent, "noise": noise},)


__all__ = ["CorrMapLatentNoiseInitializer", "trace_rows"]
