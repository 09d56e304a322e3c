"""GPU counterpart of the reference's `CreateNoiseSequenceFromIdMap` node
(source/comfyUI/stable_rendering/_nodes/loaders.py:193-271): noise / latent sequences in which every pixel that shows the
same texel starts from the same random vector.

The random draws are the reference's own torch calls in the reference's order (two `torch.manual_seed`s, two CPU base
draws, two per-key draws on the GPU's global generator), so on the same device the result equals the reference's.
The arithmetic around them — ranking the keys (the reference's `unique` sort), scattering the per-key rows into the
frames and the 8x down-sampling — is one pass over the id buffers in `csrc/srx_group.cu`; neither the `[N,7]` entry list nor
the full-resolution `[F,4,H,W]` tensors are materialised."""
from __future__ import annotations

import ctypes as C
from typing import Literal

import torch

from . import _lib
from .corrmap import IDMap

_SIZES = {"SD15": 512, "SDXL": 1024}
_MODES = {"nearest": 0, "mean": 1, "max": 2, "min": 3}


class CreateNoiseSequenceFromIdMap:
    """Same call signature and return fields (`samples`, `noise`) as the reference node."""

    Category = "loader"

    def __call__(self, id_map: IDMap, seed: int, sd_version: Literal["SD15", "SDXL"] = "SD15",
                 downsample_option: Literal["mean", "max", "min", "nearest"] = "nearest",
                 key_capacity: int = 0, rng_device: str | None = None) -> dict:
        """key_capacity: dense key-table size (default: largest vertex id + 1, one sync).
        rng_device: where the per-key rows are drawn; default = the id map's device, as the reference does when it runs
        on a GPU; "cpu" reproduces a CPU run of the reference (the rows then come from the seeded CPU generator)."""
        if sd_version not in _SIZES:
            raise ValueError("sd_version should be either SD15 or SDXL")
        if downsample_option not in _MODES:
            raise ValueError("downsample_option should be either mean, max, min, or nearest")
        if not id_map:
            raise ValueError("ID map is empty.")
        ids = id_map.tensor
        if not ids.is_cuda:
            raise _lib.SrxUnavailable("the id map must live on a CUDA device (there is no CPU path)")
        ids = ids.contiguous()
        F, H, W = int(ids.shape[0]), int(ids.shape[1]), int(ids.shape[2])
        size = _SIZES[sd_version]
        if H != W:
            # x is divided by the height and y by the width (corrmap.py:239,249): a non-square map indexes out of the latent
            raise IndexError(f"id maps must be square, got {H}x{W}")
        # An id map of another size than the node's working size: entry (x, y) lands in pixel (trunc(fl32(x/H)*size),
        # trunc(fl32(y/W)*size)) (loaders.py:218-219) — several pixels per target when the map is larger (the last one in entry
        # order wins), holes when it is smaller (the base draw stays).  The kernels invert that mapping per target pixel.
        frame_values = [int(v) for v in id_map.frame_indices]
        inv, prev = [-1] * F, [-1] * F
        for g, v in enumerate(frame_values):
            if v < -F or v >= F:
                raise IndexError(f"index {v} is out of bounds for dimension 0 with size {F}")
            prev[g] = inv[v % F]                # several id frames on one latent frame: entries are written in id-frame order,
            inv[v % F] = g                      # so the latest id frame with an entry at the target wins (as index_put does)
        dev = ids.device
        lib = _lib.load()

        # the reference's draws, in its order (loaders.py:207-226, math_utils.py:219-222)
        latent_generator = torch.manual_seed(seed)
        noise_generator = torch.manual_seed(seed + 1)
        base_latent = torch.randn([1, 4, size, size], device="cpu", generator=latent_generator).to(dev)
        base_noise = torch.randn([1, 4, size, size], device="cpu", generator=noise_generator).to(dev)

        if key_capacity <= 0:
            key_capacity = int(ids[..., 3].max().item()) + 1
        key_capacity = max(int(key_capacity), 1)
        table = torch.empty(int(lib.srx_group_rank_workspace_ints(key_capacity)), dtype=torch.int32, device=dev)
        n_unique = C.c_int64(0)
        with torch.cuda.device(dev):
            stream = _lib.current_stream_ptr(dev)
            _lib.check(lib.srx_ids_rank_table(ids.data_ptr(), _lib.torch_dtype_code(ids.dtype), F, H, W, key_capacity,
                                              table.data_ptr(), C.byref(n_unique), stream))
            u = int(n_unique.value)
            rdev = dev if rng_device is None else torch.device(rng_device)
            key_latent = torch.randn(u, 4, dtype=torch.float32, device=rdev).to(dev)   # randn_like(expanded unique, dtype=float)
            key_noise = torch.randn(u, 4, dtype=torch.float32, device=rdev).to(dev)
            mode = _MODES[downsample_option]
            h, w = size // 8, size // 8
            if mode == 0:
                latent = torch.empty(F, 4, h, w, dtype=torch.float32, device=dev)
                noise = torch.empty(F, 4, h, w, dtype=torch.float32, device=dev)
            else:
                latent = None
                noise = torch.empty(F * 4 * size * size // 32, dtype=torch.float32, device=dev)   # [F,4,S,S] viewed as [-1,4,8,8]
            a = _lib.srx_noise_args()
            a.ids_dev, a.id_dtype = ids.data_ptr(), _lib.torch_dtype_code(ids.dtype)
            a.frames, a.height, a.width = F, H, W
            inv_t = torch.tensor(inv, dtype=torch.int32, device=dev)
            prev_t = torch.tensor(prev, dtype=torch.int32, device=dev)
            a.inv_frame_dev, a.rank_table_dev = inv_t.data_ptr(), table.data_ptr()
            a.work_size, a.prev_frame_dev = size, prev_t.data_ptr()
            a.key_latent_dev, a.key_noise_dev = key_latent.data_ptr(), key_noise.data_ptr()
            a.base_latent_dev, a.base_noise_dev = base_latent.data_ptr(), base_noise.data_ptr()
            a.latent_out_dev = latent.data_ptr() if latent is not None else None
            a.noise_out_dev, a.mode = noise.data_ptr(), mode
            _lib.check(lib.srx_noise_from_ids(C.byref(a), stream))
        if mode == 0:
            return {"samples": latent, "noise": noise}
        noise = noise.view(-1, 4, h, w)                       # 2F frames: the node's view arithmetic (loaders.py:267-268)
        return {"samples": torch.zeros_like(noise), "noise": noise}


__all__ = ["CreateNoiseSequenceFromIdMap"]
