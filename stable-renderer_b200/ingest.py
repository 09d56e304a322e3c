"""G-buffer frame ingest: the GPU counterpart of `RenderManager._save_frame_data`
(reference: source/engine/managers/renderManager.py:877-948) and of the "closer pixel wins" merge of identical-G-buffer
draws (renderManager.py:121-133, buffers :219-357).

The reference reads six attachments back one by one (map, Memcpy2D, device-wide sync, flip, clone), slices them and
`torch.cat`s every frame onto growing batches; the noise attachment is mixed with a fixed background noise where nothing
was drawn, mean-pooled and AdaIN-normalised against the raw attachment.  Here one kernel pass per frame writes all
outputs into slot `n` of preallocated batches (`csrc/srx_ingest.cu`), row flip included; the keys of `data` are the ones
the reference collects in `data_to_be_added_to_engineData`."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib
from .corrmap import IDMap


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _check(t: Optional[torch.Tensor], name: str, shape, dtype, dev) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.SrxUnavailable(f"{name} must live on a CUDA device (there is no CPU path)")
    if tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != dev:
        raise ValueError(f"{name} must be {dtype} {tuple(shape)} on {dev}, got {t.dtype} {tuple(t.shape)} on {t.device}")
    return t.contiguous()


class FrameIngest:
    """Collects frames like `_save_frame_data`.  `capacity` frames are preallocated and doubled when exceeded."""

    def __init__(self, height: int, width: int, device="cuda", capacity: int = 16, bg_noise: Optional[torch.Tensor] = None):
        if height % 8 or width % 8:
            raise ValueError("frame size must be a multiple of 8")
        if not torch.cuda.is_available():
            raise _lib.SrxUnavailable("frame ingest runs on the GPU (there is no CPU path)")
        self.height, self.width = int(height), int(width)
        self.device = torch.device(device if torch.device(device).index is not None else f"cuda:{torch.cuda.current_device()}")
        self._lib = _lib.load()
        self._cap = 0
        self._n = 0
        self._buf: Dict[str, torch.Tensor] = {}
        self.frame_indices: List[int] = []
        self._bg = bg_noise
        self._canny_dtype = torch.float16          # cannyFBOTex is read as a HALF tensor (renderManager.py:353); f32 also accepted
        self._ws = torch.empty(int(self._lib.srx_ingest_workspace_bytes(self.height, self.width)), dtype=torch.uint8, device=self.device)
        self._reserve(max(1, int(capacity)))

    @property
    def GlobalBGNoise(self) -> torch.Tensor:
        """`[1,H,W,4]` float32, drawn once (renderManager.py:869-875)."""
        if self._bg is None:
            self._bg = torch.randn((1, self.height, self.width, 4), dtype=torch.float32, device=self.device)
        return self._bg

    def _shapes(self, F: int):
        H, W = self.height, self.width
        return {"color_maps": ((F, H, W, 3), torch.float16), "masks": ((F, H, W), torch.float16),
                "id_maps": ((F, H, W, 4), torch.int32), "pos_maps": ((F, H, W, 3), torch.float32),
                "normal_maps": ((F, H, W, 3), torch.float16), "depth_maps": ((F, H, W, 3), torch.float16),
                "canny_maps": ((F, H, W, 3), self._canny_dtype), "noise_maps": ((F, 4, H // 8, W // 8), torch.float32)}

    def _reserve(self, cap: int):
        if cap <= self._cap:
            return
        new = {k: torch.empty(s, dtype=d, device=self.device) for k, (s, d) in self._shapes(cap).items()}
        for k, t in self._buf.items():
            new[k][: self._n].copy_(t[: self._n])
        self._buf, self._cap = new, cap

    def __len__(self) -> int:
        return self._n

    def clear(self):
        """`data_to_be_added_to_engineData.clear()` after a prompt was submitted (renderManager.py:1025).  The batches handed
        out through `data` stay valid (the reference builds EngineData from its own `torch.cat` results): the next frames go
        into a fresh set of buffers instead of overwriting tensors a sampling run may still be reading."""
        if self._n:
            self._buf = {k: torch.empty(s, dtype=d, device=self.device) for k, (s, d) in self._shapes(self._cap).items()}
        self._n = 0
        self.frame_indices = []

    def _array_handle(self, a):
        """cudaArray_t of an attachment: a `Texture` (mapped here, unmapped by `unmap_all`), or a raw handle (int)."""
        if a is None:
            return None
        if hasattr(a, "map_array"):
            return a.map_array()
        return int(a)

    def save_frame_arrays(self, frame_count: int, color, ids, pos=None, normal_depth=None, noise=None, canny=None,
                          flip: bool = True, canny_dtype: torch.dtype = torch.float32):
        """Zero-copy form of `save_frame_data`: the attachments are the mapped cudaArrays of the GL textures (`Texture`
        objects of this package, or raw `cudaArray_t` handles) and the ingest kernel reads them through surface objects —
        no staging tensor per attachment, no device-wide sync, no flip pass (replaces the seven `Texture.tensor()` reads of
        renderManager.py:882-943).  Arrays keep the GL row order, hence `flip=True` by default."""
        H, W, dev = self.height, self.width, self.device
        if color is None or ids is None:
            raise ValueError("the colour and id attachments are required")
        if canny is not None and canny_dtype != self._canny_dtype:
            if self._n or canny_dtype not in (torch.float16, torch.float32):
                raise ValueError(f"canny must be {self._canny_dtype} like the frames collected so far")
            self._canny_dtype = canny_dtype
            self._buf["canny_maps"] = torch.empty(self._shapes(self._cap)["canny_maps"][0], dtype=canny_dtype, device=dev)
        if self._n == self._cap:
            self._reserve(self._cap * 2)
        textures = [t for t in (color, ids, pos, normal_depth, noise, canny) if hasattr(t, "map_array")]
        try:
            arr = _lib.srx_gbuffer_arrays(self._array_handle(color), self._array_handle(ids), self._array_handle(pos),
                                          self._array_handle(normal_depth), self._array_handle(noise), self._array_handle(canny),
                                          _lib.torch_dtype_code(self._canny_dtype))
            a = _lib.srx_ingest_args()
            a.height, a.width, a.flip_rows, a.frame_slot = H, W, int(bool(flip)), self._n
            self._fill_outputs(a, pos is not None, normal_depth is not None, canny is not None, noise is not None)
            with torch.cuda.device(dev):
                _lib.check(self._lib.srx_frame_ingest_arrays(C.byref(a), C.byref(arr), _lib.current_stream_ptr(dev)))
        finally:
            for t in textures:
                t.unmap_array()
        self.frame_indices.append(int(frame_count))
        self._n += 1

    def _fill_outputs(self, a, has_pos: bool, has_nd: bool, has_canny: bool, has_noise: bool):
        b, n = self._buf, self._n
        a.bg_noise = self.GlobalBGNoise.data_ptr() if has_noise else None
        a.color_maps, a.masks, a.id_maps = b["color_maps"].data_ptr(), b["masks"].data_ptr(), b["id_maps"].data_ptr()
        a.pos_maps = b["pos_maps"].data_ptr() if has_pos else None
        a.normal_maps = b["normal_maps"].data_ptr() if has_nd else None
        a.depth_maps = b["depth_maps"].data_ptr() if has_nd else None
        a.canny_maps = b["canny_maps"].data_ptr() if has_canny else None
        a.noise_maps = b["noise_maps"].data_ptr() if has_noise else None
        a.workspace, a.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        # an attachment that is absent for this frame leaves zeros in its slot, not uninitialised memory
        for key, present in (("pos_maps", has_pos), ("normal_maps", has_nd), ("depth_maps", has_nd), ("canny_maps", has_canny),
                             ("noise_maps", has_noise)):
            if not present:
                b[key][n].zero_()

    def save_frame_data(self, frame_count: int, color: torch.Tensor, ids: torch.Tensor, pos: Optional[torch.Tensor] = None,
                        normal_depth: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                        canny: Optional[torch.Tensor] = None, flip: bool = False):
        """One frame of attachments (`[H,W,C]` device tensors in `Texture.tensor()` layout; `flip=True` when they are still
        in GL row order).  Attachments that are None leave their batch untouched for this frame."""
        H, W, dev = self.height, self.width, self.device
        color = _check(color, "color", (H, W, 4), torch.float16, dev)
        ids = _check(ids, "ids", (H, W, 4), torch.int32, dev)
        pos = _check(pos, "pos", (H, W, 3), torch.float32, dev)
        normal_depth = _check(normal_depth, "normal_depth", (H, W, 4), torch.float16, dev)
        noise = _check(noise, "noise", (H, W, 4), torch.float16, dev)
        if canny is not None and canny.dtype != self._canny_dtype:
            if self._n or canny.dtype not in (torch.float16, torch.float32):
                raise ValueError(f"canny must be {self._canny_dtype} like the frames collected so far")
            self._canny_dtype = canny.dtype
            self._buf["canny_maps"] = torch.empty(self._shapes(self._cap)["canny_maps"][0], dtype=canny.dtype, device=dev)
        canny = _check(canny, "canny", (H, W, 3), self._canny_dtype, dev)
        if color is None or ids is None:
            raise ValueError("the colour and id attachments are required")
        if self._n == self._cap:
            self._reserve(self._cap * 2)
        a = _lib.srx_ingest_args()
        a.src = _lib.srx_gbuffer(_ptr(color), _ptr(ids), _ptr(pos), _ptr(normal_depth), _ptr(noise), _ptr(canny),
                                 _lib.torch_dtype_code(self._canny_dtype))
        a.height, a.width, a.flip_rows, a.frame_slot = H, W, int(bool(flip)), self._n
        self._fill_outputs(a, pos is not None, normal_depth is not None, canny is not None, noise is not None)
        with torch.cuda.device(dev):
            _lib.check(self._lib.srx_frame_ingest(C.byref(a), _lib.current_stream_ptr(dev)))
        self.frame_indices.append(int(frame_count))
        self._n += 1

    @property
    def data(self) -> dict:
        """What the reference hands to `EngineData(**data_to_be_added_to_engineData)` (renderManager.py:1002-1011): batches
        of the frames collected so far (views, no copy), `id_maps` as an `IDMap` carrying `frame_indices`."""
        n = self._n
        out = {k: t[:n] for k, t in self._buf.items()}
        out["frame_indices"] = list(self.frame_indices)
        if n:
            out["id_maps"] = IDMap(frame_indices=list(self.frame_indices), tensor=out["id_maps"])
        return out


class GBufferTemp:
    """`RenderManager._*_buffer_temp` (renderManager.py:219-357) with the merge of renderManager.py:121-133: draws of
    identical-G-buffer tasks are merged pixel by pixel, the closer one (larger reversed depth) wins."""

    def __init__(self, height: int, width: int, device="cuda"):
        if not torch.cuda.is_available():
            raise _lib.SrxUnavailable("the G-buffer merge runs on the GPU (there is no CPU path)")
        H, W = int(height), int(width)
        self.height, self.width = H, W
        dev = torch.device(device if torch.device(device).index is not None else f"cuda:{torch.cuda.current_device()}")
        self.device = dev
        self.color = torch.zeros((H, W, 4), dtype=torch.float16, device=dev)
        self.ids = torch.zeros((H, W, 4), dtype=torch.int32, device=dev)
        self.pos = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
        self.normal = torch.zeros((H, W, 3), dtype=torch.float16, device=dev)
        self.depth = torch.zeros((H, W), dtype=torch.float16, device=dev)
        self.noise = torch.zeros((H, W, 4), dtype=torch.float16, device=dev)
        self.canny = torch.zeros((H, W, 3), dtype=torch.float16, device=dev)
        self._lib = _lib.load()

    def clear(self):
        for t in (self.color, self.ids, self.pos, self.normal, self.depth, self.noise, self.canny):
            t.zero_()

    def merge_closer(self, color, ids, pos, normal_depth, noise, canny, flip: bool = False):
        H, W, dev = self.height, self.width, self.device
        color = _check(color, "color", (H, W, 4), torch.float16, dev)
        ids = _check(ids, "ids", (H, W, 4), torch.int32, dev)
        pos = _check(pos, "pos", (H, W, 3), torch.float32, dev)
        normal_depth = _check(normal_depth, "normal_depth", (H, W, 4), torch.float16, dev)
        noise = _check(noise, "noise", (H, W, 4), torch.float16, dev)
        canny_dtype = canny.dtype if canny is not None else torch.float16
        if canny_dtype not in (torch.float16, torch.float32):
            raise ValueError("canny must be float16 or float32")
        canny = _check(canny, "canny", (H, W, 3), canny_dtype, dev)
        if normal_depth is None:
            raise ValueError("the normal+depth attachment is required")
        cur = _lib.srx_gbuffer(_ptr(color), _ptr(ids), _ptr(pos), _ptr(normal_depth), _ptr(noise), _ptr(canny),
                               _lib.torch_dtype_code(canny_dtype))
        tmp = _lib.srx_gbuffer_temp(self.color.data_ptr(), self.ids.data_ptr(), self.pos.data_ptr(), self.normal.data_ptr(),
                                    self.depth.data_ptr(), self.noise.data_ptr(), self.canny.data_ptr())
        with torch.cuda.device(dev):
            _lib.check(self._lib.srx_gbuffer_merge_closer(C.byref(cur), H, W, int(bool(flip)), C.byref(tmp),
                                                          _lib.current_stream_ptr(dev)))

    def merge_closer_arrays(self, color, ids, pos, normal_depth, noise, canny, flip: bool = True,
                            canny_dtype: torch.dtype = torch.float32):
        """`merge_closer` reading the mapped cudaArrays of the attachments (`Texture` objects or raw `cudaArray_t` handles)."""
        if normal_depth is None:
            raise ValueError("the normal+depth attachment is required")
        H, W, dev = self.height, self.width, self.device
        srcs = (color, ids, pos, normal_depth, noise, canny)
        textures = [t for t in srcs if hasattr(t, "map_array")]
        try:
            h = [None if t is None else (t.map_array() if hasattr(t, "map_array") else int(t)) for t in srcs]
            cur = _lib.srx_gbuffer_arrays(h[0], h[1], h[2], h[3], h[4], h[5], _lib.torch_dtype_code(canny_dtype))
            tmp = _lib.srx_gbuffer_temp(self.color.data_ptr(), self.ids.data_ptr(), self.pos.data_ptr(), self.normal.data_ptr(),
                                        self.depth.data_ptr(), self.noise.data_ptr(), self.canny.data_ptr())
            with torch.cuda.device(dev):
                _lib.check(self._lib.srx_gbuffer_merge_closer_arrays(C.byref(cur), H, W, int(bool(flip)), C.byref(tmp),
                                                                     _lib.current_stream_ptr(dev)))
        finally:
            for t in textures:
                t.unmap_array()


__all__ = ["FrameIngest", "GBufferTemp"]
