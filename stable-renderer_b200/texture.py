"""`Texture` — the tensor side of the engine's GL textures (reference: source/engine/static/texture/texture.py:166-254,
326-408; the cuda-python wrappers of source/common_utils/cuda_utils.py:101-190).

The reference keeps a persistent `[H,W,C]` CUDA tensor per shared texture and moves texels with pycuda `Memcpy2D` between
the mapped `cudaArray` and that tensor, followed by `torch.cuda.synchronize()` and a separate `.flip(0)`.  Here the mapped
array is a CUDA-graphics-resource mapping used in place: `tensor()` / `set_data()` are one kernel each on the current
stream (row flip, offsets, RGB -> RGBA padding fused; no device-wide sync), and the ingest / bake kernels can read the
mapped array directly (`map_array()`), so that no staging copy exists at all (`FrameIngest.save_frame_arrays`).

The GL object itself belongs to the engine: pass its texture id (`gl_texture=`) and the wrapper registers it with
`cudaGraphicsGLRegisterImage` (needs the GL context current on this thread).  Without a GL context — tests, headless
tools — `cuda_array=` wraps a plain `cudaArray_t`, or `Texture.offscreen()` allocates one."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Union

import numpy as np
import torch

from . import _lib

GL_TEXTURE_2D = 0x0DE1
GL_TEXTURE_2D_ARRAY = 0x8C1A
_ONE_BITS = {torch.float16: 0x3C00, torch.bfloat16: 0x3F80, torch.float32: 0x3F800000, torch.int32: 1, torch.int16: 1,
             torch.uint8: 1, torch.int8: 1}
_KIND = {torch.float16: 2, torch.float32: 2, torch.int32: 0, torch.int16: 0, torch.int8: 0, torch.uint8: 1}


class Texture:
    def __init__(self, width: int, height: int, channel_count: int, dtype: torch.dtype = torch.float16, *,
                 gl_texture: Optional[int] = None, gl_target: int = GL_TEXTURE_2D, cuda_array: Optional[int] = None,
                 name: Optional[str] = None, share_to_torch: bool = True, device=None):
        if dtype not in _ONE_BITS:
            raise ValueError(f"unsupported texture element type {dtype}")
        self.width, self.height, self.channel_count = int(width), int(height), int(channel_count)
        self.dtype, self.name, self.share_to_torch = dtype, name, bool(share_to_torch)
        self.texID = gl_texture
        self._gl_target = gl_target
        self._array = cuda_array          # a plain cudaArray_t (no GL): always "mapped"
        self._owns_array = False
        self._res = None                  # srx_gl_resource*
        self._mapped = 0
        self._tensor: Optional[torch.Tensor] = None
        self._tensor_flipped = False
        self.device = torch.device(device) if device is not None else None
        self._lib = None

    # -- construction helpers ------------------------------------------------------------------------------------
    @classmethod
    def offscreen(cls, width: int, height: int, channel_count: int, dtype: torch.dtype = torch.float16, **kw) -> "Texture":
        """A texture backed by a plain CUDA array (no GL context): what a registered GL texture looks like to CUDA.
        Three-channel textures get a four-channel array, like RGB32F under `cudaGraphicsGLRegisterImage`."""
        lib = _lib.load()
        arr = C.c_void_p()
        ch = 4 if channel_count == 3 else channel_count
        _lib.check(lib.srx_array_alloc(C.byref(arr), width, height, ch, torch.empty(0, dtype=dtype).element_size() * 8, _KIND[dtype]))
        t = cls(width, height, channel_count, dtype, cuda_array=arr.value, **kw)
        t._owns_array = True
        return t

    def _dev(self) -> torch.device:
        if self.device is None:
            if not torch.cuda.is_available():
                raise _lib.SrxUnavailable("texture <-> tensor interop runs on the GPU (there is no CPU path)")
            self.device = torch.device("cuda", torch.cuda.current_device())
        return self.device

    def _init_tensor(self):
        """texture.py:166-202: the persistent `[H,W,C]` CUDA tensor and the CUDA registration of the GL texture."""
        assert self.share_to_torch, "This texture is not set to be sharing to torch."
        self._lib = _lib.load()
        if self._tensor is None:
            self._tensor = torch.zeros((self.height, self.width, self.channel_count), dtype=self.dtype, device=self._dev())
        if self._array is None and self._res is None:
            assert self.texID is not None, "This texture is not yet sent to GPU."
            res = C.c_void_p()
            _lib.check(self._lib.srx_gl_register_image(C.byref(res), int(self.texID), int(self._gl_target), 0))
            self._res = res

    # -- the mapping itself ----------------------------------------------------------------------------------------
    def map_array(self, layer: int = 0) -> int:
        """`cudaArray_t` of the texture (mapped for CUDA until `unmap_array`); kernels read / write it in place."""
        if self._array is not None:
            return int(self._array)
        if self._res is None:
            self._init_tensor()
        arr = C.c_void_p()
        _lib.check(self._lib.srx_gl_map(self._res, C.byref(arr), _lib.current_stream_ptr(self._dev())))
        self._mapped += 1
        if layer:
            _lib.check(self._lib.srx_gl_mapped_layer(self._res, int(layer), C.byref(arr)))
        return int(arr.value)

    def unmap_array(self) -> None:
        if self._array is not None or self._res is None or self._mapped == 0:
            return
        self._mapped -= 1
        if self._mapped == 0:
            _lib.check(self._lib.srx_gl_unmap(self._res, _lib.current_stream_ptr(self._dev())))

    # -- reference API -----------------------------------------------------------------------------------------
    def tensor(self, update: bool = True, flip: bool = True) -> torch.Tensor:
        """texture.py:221-254.  `update=False` returns the tensor of the last update; `flip` = top-left origin (GL's is
        bottom-left).  One kernel on the current stream; the returned tensor is the persistent one (overwritten by the next
        update — `.clone()` it to keep a frame, as the reference's callers do, renderManager.py:884)."""
        if self._tensor is None or self._lib is None:
            self._init_tensor()
        if not update:
            return self._tensor if flip == self._tensor_flipped else self._tensor.flip(0)
        dev = self._dev()
        arr = self.map_array()
        try:
            with torch.cuda.device(dev):
                _lib.check(self._lib.srx_array_to_tensor_ch(arr, self._tensor.data_ptr(), self.width, self.height,
                                                            self._tensor.element_size(), self.channel_count, int(bool(flip)),
                                                            _lib.current_stream_ptr(dev)))
        finally:
            self.unmap_array()
        self._tensor_flipped = bool(flip)
        return self._tensor

    def set_data(self, data: Union[torch.Tensor, np.ndarray], xOffset: int = 0, yOffset: int = 0, width: Optional[int] = None,
                 height: Optional[int] = None, flip: bool = False):
        """texture.py:326-408 for tensors / arrays: `[height,width,C]` data into the region at (xOffset, yOffset).  Like the
        reference: a transposed `[width,height,C]` input is transposed back, single-channel data is repeated, RGB data for
        an RGBA texture gets alpha = 1, wider data is truncated, the dtype is cast.  The offsets address the TEXTURE region
        (the reference's GPU path applies them to the source, texture.py:393-394 — its glTexSubImage2D path, :338, to the
        texture)."""
        if self._lib is None:
            self._init_tensor()
        width = width or (self.width - xOffset)
        height = height or (self.height - yOffset)
        if isinstance(data, np.ndarray):
            data = torch.from_numpy(np.ascontiguousarray(data))
        if not isinstance(data, torch.Tensor):
            raise Exception("Invalid data type: {}".format(type(data)))
        if data.dim() == 2:
            data = data.unsqueeze(-1)
        if data.shape[0] != height or data.shape[1] != width:
            if data.shape[1] == height and data.shape[0] == width:
                data = data.transpose(0, 1)
            else:
                raise Exception("The data shape should be [height, width, channel_count]. Got: {}".format(tuple(data.shape)))
        src_ch = data.shape[2]
        if src_ch != self.channel_count:
            if src_ch < self.channel_count:
                if not (src_ch == 1 or (src_ch == 3 and data.shape[0] == self.height and data.shape[1] == self.width)):
                    raise Exception("Invalid data shape: {}".format(tuple(data.shape)))
            else:
                data = data[:, :, :self.channel_count]
                src_ch = self.channel_count
        dev = self._dev()
        data = data.to(device=dev, dtype=self.dtype).contiguous()
        arr = self.map_array()
        try:
            with torch.cuda.device(dev):
                _lib.check(self._lib.srx_tensor_to_array_ch(arr, data.data_ptr(), int(width), int(height), data.element_size(),
                                                            int(src_ch), int(bool(flip)), int(xOffset), int(yOffset),
                                                            _ONE_BITS[self.dtype], _lib.current_stream_ptr(dev)))
        finally:
            self.unmap_array()

    def clear(self):
        """Releases the CUDA side (registration / offscreen array); the GL object belongs to the engine."""
        lib = self._lib or (_lib.load() if (self._res is not None or self._owns_array) else None)
        if self._res is not None:
            lib.srx_gl_unregister(self._res)
            self._res = None
        if self._owns_array and self._array is not None:
            lib.srx_array_free(C.c_void_p(self._array))
        self._array, self._owns_array, self._tensor = None, False, None

    def __del__(self):  # pragma: no cover
        try:
            self.clear()
        except Exception:
            pass


__all__ = ["Texture", "GL_TEXTURE_2D", "GL_TEXTURE_2D_ARRAY"]
