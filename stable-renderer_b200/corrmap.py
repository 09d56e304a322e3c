"""`IDMap` and `CorrespondMap` with the reference's fields and method signatures
(reference: source/engine/static/corrmap.py:49-280 and :373-736), backed by the CUDA kernels of libsrx.so.

Differences that are deliberate:
  * tensors live on a CUDA device; nothing here runs arithmetic on the CPU;
  * OpenGL upload / binding methods (`load`, `bind`, `set_data`, corrmap.py:443-529) belong to the renderer, not to
    this path, and are not provided;
  * a 3-D id tensor handed to `CorrespondMap.update` is one frame (the reference loops forever, corrmap.py:629-630);
  * mask + sprite/material filters are applied consistently (the reference re-indexes a compacted colour array with
    original pixel indices, corrmap.py:715).
"""
from __future__ import annotations

import ctypes as C
import os
from functools import partial
from pathlib import Path
from typing import Literal, Optional, Sequence, TypeAlias

import numpy as np
import torch
from attr import attrib, attrs
from torch import Tensor

from . import _lib

UpdateMode: TypeAlias = Literal["replace", "replace_avg", "first", "first_avg"]
BakeWeight: TypeAlias = Literal["none", "uniform", "view_normal", "view_normal_depth"]
NO_ID_MAP_INDEX = 2048


def extract_index(file_path, i):
    """File-name ordering rule of the reference's dumps (source/common_utils/path_utils.py:173-179)."""
    name = os.path.basename(file_path)
    stem = name.split(".")[0]
    if stem.split("_")[-1].isdigit():   # e.g. id_12.npy
        return int(stem.split("_")[-1])
    if name.split("_")[0].isdigit():    # e.g. 12_id.npy
        return int(name.split("_")[0])
    return i


def _default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.SrxUnavailable("no CUDA device: the overlap / bake path has no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device())


@attrs
class IDMap:
    """Per-frame correspondence ids (reference corrmap.py:49-136).

    tensor [N,H,W,4] = (spriteID, materialID, map_index, vertexID); `masks` [N,H,W] float32, 1 = no id."""

    tensor: Tensor = attrib()
    frame_indices: list = attrib(default=None)
    masks: Tensor = attrib(default=None)
    _vertex_screen_info_cache: Tensor = attrib(default=None, init=False)
    _plans: dict = attrib(factory=dict, init=False, repr=False)
    _feature_buckets: dict = attrib(factory=dict, init=False, repr=False)     # feature.py: bucketing passes of the feature overlap
    _device_ids: dict = attrib(factory=dict, init=False, repr=False)

    @property
    def frame_count(self) -> int:
        return len(self.frame_indices)

    def __getitem__(self, index: int) -> Tensor:
        return self.tensor[index]

    def __len__(self):
        return self.frame_count

    @property
    def height(self) -> int:  # same (surprising) definition as corrmap.py:85-88 for NHWC ids
        return self.tensor.shape[-2]

    @property
    def width(self) -> int:   # corrmap.py:90-93
        return self.tensor.shape[-1]

    def __attrs_post_init__(self):
        if self.frame_indices is None and self.tensor is not None:
            if not isinstance(self.tensor, torch.Tensor):
                raise ValueError("Invalid type of real ID tensor.")
            if self.tensor.dim() == 4:
                n = self.tensor.shape[0]
            elif self.tensor.dim() == 3:
                n = 1
            else:
                raise ValueError("Invalid shape of real ID tensor.")
            self.frame_indices = list(range(n))
        elif isinstance(self.frame_indices, int):
            self.frame_indices = [self.frame_indices]
        if isinstance(self.tensor, torch.Tensor) and self.tensor.dim() == 3:
            self.tensor = torch.stack([self.tensor] * len(self.frame_indices), dim=0)
        if self.masks is None:
            if self.tensor is not None:   # corrmap.py:119-124
                no_id = (self.tensor[..., 2] == NO_ID_MAP_INDEX) | torch.all(self.tensor == 0, dim=-1)
                self.masks = no_id.to(torch.float32)
        elif self.masks.dim() == 2:
            self.masks = torch.stack([self.masks] * self.frame_count, dim=0)

    def __deepcopy__(self, memo=None):
        return IDMap(frame_indices=list(self.frame_indices), tensor=self.tensor.clone(),
                     masks=self.masks.clone() if self.masks is not None else None)

    @classmethod
    def from_directory(cls, directory, frame_start: int | None = None, num_frames: int | None = None,
                       use_frame_indices_from_filename: bool = True, device=None) -> "IDMap":
        """Loads `id_N.npy` dumps ordered by `extract_index` (reference corrmap.py:138-198)."""
        assert os.path.exists(directory)
        frame_start = frame_start or 0
        filenames = [f for f in os.listdir(directory) if f.endswith(".npy")]
        ordered = sorted(filenames, key=lambda x: extract_index(x, filenames.index(x)))
        if use_frame_indices_from_filename:
            frame_indices = list(map(partial(extract_index, i=-1), ordered))
        else:
            frame_indices = list(range(len(ordered)))
        num_frames = num_frames or len(frame_indices)
        frame_indices = frame_indices[frame_start: frame_start + num_frames]
        assert all(i != -1 for i in frame_indices), "Illegal filename(s) found."
        tensors = []
        for name in ordered[frame_start: frame_start + num_frames]:
            t = torch.from_numpy(np.load(os.path.join(directory, name))).squeeze()
            if t.dim() != 3:
                raise ValueError(f"Invalid shape of id tensor: {t.shape}.")
            if not (t.shape[-1] == 4 or t.shape[1] == 4):
                raise ValueError(f"Invalid id tensor shape: {t.shape}.")
            tensors.append(t)
        if len(tensors) == 0:
            raise ValueError("No valid id data found.")
        if any(t.shape != tensors[0].shape for t in tensors):
            raise ValueError("Tensor data has inconsistent shapes.")
        t = torch.stack(tensors, dim=0)
        if device is not None:
            t = t.to(device)
        return cls(frame_indices=frame_indices, tensor=t)

    @classmethod
    def from_tensor(cls, frame_indices: list, tensor: Tensor) -> "IDMap":
        if tensor.dim() != 4:
            raise ValueError(f"Tensor should be in (B, H, W, C), got shape {tensor.shape}")
        if len(frame_indices) != tensor.shape[0]:
            raise ValueError(f"Frame indices count should be equal to the batch size of the tensor, got "
                             f"{len(frame_indices)} and {tensor.shape[0]}")
        return cls(frame_indices=frame_indices, tensor=tensor)

    # -- device plumbing -------------------------------------------------------------------------------------
    def device_ids(self, device: torch.device) -> Tensor:
        """The id buffers on `device` as contiguous int32/int16 (cached; the reference moves the derived
        vertex_screen_info instead, corresponder.py:309)."""
        device = torch.device(device)
        t = self._device_ids.get(device)
        if t is None:
            t = self.tensor
            if t.dtype not in (torch.int32, torch.int16):
                t = t.to(torch.int32)
            t = t.to(device).contiguous()
            self._device_ids[device] = t
        return t

    def invalidate(self) -> None:
        """The id buffers were overwritten in place (a new batch of frames in the same tensor): drop everything derived
        from the old content.  The reference has the same staleness hazard in `_vertex_screen_info_cache`
        (corrmap.py:226,278) and avoids it by building a new IDMap per batch; this keeps plans (and, frame-sharded, their
        symmetric-memory workspaces) alive across batches."""
        self._vertex_screen_info_cache = None
        self._feature_buckets.clear()
        same = self._device_ids.get(self.tensor.device) if isinstance(self.tensor, Tensor) else None
        self._device_ids = {self.tensor.device: same} if same is not None and same.data_ptr() == self.tensor.data_ptr() else {}
        for key, plan in list(self._plans.items()):
            plan.cached = False
            plan._calls = 0
            plan._checked = False          # a capacity hint must be re-validated against the new ids (first step checks)
            # the plan streams `plan._ids`: point it at the NEW content.  When the device copy aliases `tensor` it already
            # is; otherwise (CPU tensor, int64 ids, another device) a fresh device copy is made here.
            fresh = self.device_ids(plan.device)
            if tuple(fresh.shape) != tuple(plan.id_shape) or fresh.dtype != plan.id_dtype:
                plan.close()
                del self._plans[key]
            else:
                plan._ids = fresh

    def create_vertex_screen_info(self) -> Tensor:
        """[N,7] float32 (object, material, map_index, vertex_id, x/H, y/W, frame_index) in (frame,y,x) order —
        bit-identical to the reference's tensor (corrmap.py:220-280), built by `srx_vertex_screen_info`."""
        if self._vertex_screen_info_cache is None:
            dev = self.tensor.device if self.tensor.is_cuda else _default_device()
            ids = self.device_ids(dev)
            F, H, W, _ = ids.shape
            out = torch.empty(F * H * W, 7, dtype=torch.float32, device=dev)
            n = C.c_int64(0)
            fv = (C.c_int32 * F)(*[int(v) for v in self.frame_indices])
            with torch.cuda.device(dev):
                _lib.check(_lib.load().srx_vertex_screen_info(ids.data_ptr(), _lib.torch_dtype_code(ids.dtype), F, H, W, fv,
                                                              out.data_ptr(), C.byref(n), _lib.current_stream_ptr(dev)))
            self._vertex_screen_info_cache = out[: n.value].clone().to(self.tensor.device)
        return self._vertex_screen_info_cache


@attrs(repr=False, eq=False)
class CorrespondMap:
    """The UV-texture atlas (reference corrmap.py:373-411): `_values` fp16 [k*k, height*width, C] and
    `_writtens` bool [k*k, height*width], vertexID = width*y + x."""

    name: Optional[str] = attrib(default=None)
    k: int = attrib(default=3, kw_only=True)
    height: int = attrib(default=512, kw_only=True)
    width: int = attrib(default=512, kw_only=True)
    channel_count: int = attrib(default=4, kw_only=True)
    device: Optional[torch.device] = attrib(default=None, kw_only=True)
    _values: Tensor = attrib(init=False)
    _writtens: Tensor = attrib(init=False)
    vertex_screen_info: Optional[Tensor] = attrib(default=None, init=False)
    _workspace: Optional[Tensor] = attrib(default=None, init=False)

    def __attrs_post_init__(self):
        self.device = torch.device(self.device) if self.device is not None else _default_device()
        self._values = torch.zeros(self.k * self.k, self.height * self.width, self.channel_count, dtype=torch.float16,
                                   device=self.device)
        self._writtens = torch.zeros(self.k * self.k, self.height * self.width, dtype=torch.bool, device=self.device)

    def __repr__(self):
        return f"<CorrespondMap: {self.name or 'untitled'}, k={self.k}, size={self.height}x{self.width}, channel_count={self.channel_count}>"

    def __getitem__(self, index):
        return self._values[index]

    def clear(self):
        self._values.zero_()
        self._writtens.zero_()

    def numpy_data(self, map_index: int, width: int | None = None, height: int | None = None, dtype=np.float16):
        return self.get_map(map_index, height, width).cpu().numpy().astype(dtype)

    def get_map(self, index: int, height: int | None = None, width: int | None = None,
                order: Literal["whc", "hwc"] = "hwc"):
        height = height or self.height
        width = width or self.width
        assert height * width <= self.height * self.width, "The given size is larger than the original size."
        if height * width < self.height * self.width:
            return self._values[index, : width * height].view(height, width, self.channel_count)
        data = self._values[index].view(height, width, self.channel_count)
        if order == "whc":
            data = data.permute(1, 0, 2)
        return data

    def get_maps(self, height: int | None = None, width: int | None = None):
        width = self.width if width is None else width
        height = self.height if height is None else height
        if width * height < self.width * self.height:
            return self._values[:, : width * height].view(self.k * self.k, height, width, self.channel_count)
        return self._values.view(self.k * self.k, height, width, self.channel_count)

    def get_written_flag_map(self, index: int, height: int | None = None, width: int | None = None):
        height = self.height if height is None else height
        width = self.width if width is None else width
        if height * width < self.height * self.width:
            return self._writtens[index, : width * height].view(height, width)
        return self._writtens[index].view(height, width)

    # -- the bake ------------------------------------------------------------------------------------------------
    def update(self, color_frames, id_maps, spriteID: int | None = None, materialID: int | None = None,
               mode: UpdateMode = "first_avg", masks=None, inverse_masks: bool = False, ignore_obj_mat_id: bool = False,
               weight_mode: BakeWeight = "none", normal_depth: Optional[Tensor] = None, process_group=None,
               phase: int = 0, frame_offset: int = 0, frames_global: int = 0, defer_check: bool = False):
        """`CorrespondMap.update` (reference corrmap.py:578-670): same arguments; all frames go to the GPU in one call.

        defer_check: by default a pixel that addresses a texel outside the atlas raises IndexError from this call, as in the
        reference — which costs one host sync per call.  With `defer_check=True` the call only enqueues its kernels (no host
        sync, CUDA-graph capturable); the status stays on the device until `check()` raises for everything since the last check.

        weight_mode / normal_depth select the depth/normal-weighted multi-view bake, which the
        reference lists as TODO (README.md:18-19); "none" is the reference behaviour.
        process_group: view-sharded multi-GPU bake; every rank passes its own contiguous block of views (rank order = view
        order) and ends with the same atlas.  Weighted modes: the weighted sums are all-reduced (`phase` 1 accumulate /
        2 finalise for callers that do the exchange themselves).  Reference modes: the order keys number the views of all
        ranks, the claims are MAX-reduced, the ranks' winning texels SUM-reduced (`phase` 1 claim / 2 write / 3 merge with
        `frame_offset`, `frames_global` for callers that do the exchange themselves)."""
        if mode not in ("replace", "replace_avg", "first", "first_avg"):
            raise ValueError(f"unknown update mode {mode}")
        colors = self._stack(color_frames, "color_frames")
        if isinstance(id_maps, IDMap):
            id_maps = id_maps.tensor
        if isinstance(id_maps, (list, tuple)):
            id_maps = [m.tensor if isinstance(m, IDMap) else m for m in id_maps]
        ids = self._stack(id_maps, "id_maps")
        if colors.shape[0] != ids.shape[0]:
            raise ValueError(f"The length of color_frames and id_maps should be the same, but got: {colors.shape[0]} and {ids.shape[0]}")
        if colors.shape[1:3] != ids.shape[1:3]:
            raise ValueError(f"colour frames {tuple(colors.shape)} and id maps {tuple(ids.shape)} differ in size")
        mk = None
        if masks is not None:
            if isinstance(masks, (list, tuple)):
                masks = torch.stack([m.squeeze() for m in masks], dim=0)
            if not isinstance(masks, Tensor):
                raise ValueError("Invalid type of masks. Got: ", type(masks))
            if masks.dim() == 4 and masks.shape[-1] == 1:
                masks = masks.squeeze(-1)
            elif masks.dim() == 3 and masks.shape[-1] == 1 and ids.shape[0] == 1:
                masks = masks.squeeze(-1).unsqueeze(0)
            if masks.dim() == 2:
                masks = masks.unsqueeze(0)
            if masks.dim() != 3:
                raise ValueError("The shape of masks is invalid. Got: ", masks.shape)
            if masks.shape[0] != colors.shape[0]:
                raise ValueError(f"The length of masks should be the same as color_frames, but got: {masks.shape[0]} and {colors.shape[0]}")
            mk = masks.to(device=self.device, dtype=torch.float32).contiguous()
        if ids.dtype not in (torch.int32, torch.int16):
            ids = ids.to(torch.int32)
        ids = ids.to(self.device).contiguous()
        if colors.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            colors = colors.to(torch.float32)
        colors = colors.to(self.device).contiguous()
        nd = None
        if weight_mode in ("view_normal", "view_normal_depth"):
            if normal_depth is None:
                raise ValueError(f"weight_mode={weight_mode} needs the normal+depth attachment")
            nd = normal_depth.to(device=self.device, dtype=torch.float16).contiguous()
            if nd.shape[:3] != ids.shape[:3] or nd.shape[-1] != 4:
                raise ValueError("normal_depth must be [F,H,W,4]")
        lib = _lib.load()
        wm = _lib.SRX_BAKE_WEIGHT[weight_mode]
        sharded = process_group is not None
        if sharded:
            import torch.distributed as dist
            sharded = dist.get_world_size(process_group) > 1
        need = int(lib.srx_bake_workspace_bytes(self.k * self.k, self.height * self.width, self.channel_count, wm))
        if wm == 0 and (sharded or phase):
            need = int(lib.srx_bake_sharded_workspace_bytes(self.k * self.k, self.height * self.width, self.channel_count))
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.zeros(need, dtype=torch.uint8, device=self.device)    # zeroed: the deferred status word is sticky
        a = _lib.srx_bake_args()
        a.defer_status = 1 if defer_check else 0
        self._last_args = a
        a.values_dev = self._values.data_ptr()
        a.writtens_dev = self._writtens.data_ptr()
        a.k2, a.texels, a.channels = self.k * self.k, self.height * self.width, self.channel_count
        a.colors_dev, a.color_dtype, a.color_channels = colors.data_ptr(), _lib.torch_dtype_code(colors.dtype), colors.shape[-1]
        a.ids_dev, a.id_dtype = ids.data_ptr(), _lib.torch_dtype_code(ids.dtype)
        a.masks_dev = mk.data_ptr() if mk is not None else None
        a.inverse_masks = 1 if inverse_masks else 0
        a.frames, a.height, a.width = ids.shape[0], ids.shape[1], ids.shape[2]
        a.sprite_id = -1 if spriteID is None else int(spriteID)
        a.material_id = -1 if materialID is None else int(materialID)
        a.ignore_obj_mat_id = 1 if ignore_obj_mat_id else 0
        a.mode = _lib.SRX_BAKE_MODE[mode]
        a.weight_mode = wm
        a.normal_depth_dev = nd.data_ptr() if nd is not None else None
        a.workspace_dev, a.workspace_bytes = self._workspace.data_ptr(), self._workspace.numel()
        a.frame_offset, a.frames_global = int(frame_offset), int(frames_global)
        with torch.cuda.device(self.device):
            stream = _lib.current_stream_ptr(self.device)
            if sharded and wm == 0:
                # reference modes: "last pixel in view order wins" is a maximum over order keys, not a sum
                ntex = self.k * self.k * self.height * self.width
                counts = torch.zeros(dist.get_world_size(process_group), dtype=torch.int64, device=self.device)
                counts[dist.get_rank(process_group)] = ids.shape[0]
                dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=process_group)
                counts = counts.tolist()
                a.frame_offset, a.frames_global = int(sum(counts[:dist.get_rank(process_group)])), int(sum(counts))
                a.phase = 1
                _lib.check(lib.srx_bake_update(C.byref(a), stream))
                dist.all_reduce(self._workspace[:ntex * 4].view(torch.int32), op=dist.ReduceOp.MAX, group=process_group)
                a.phase = 2
                _lib.check(lib.srx_bake_update(C.byref(a), stream))
                lo = (ntex * 4 + 255) // 256 * 256 + 256
                dist.all_reduce(self._workspace[lo:need].view(torch.int32), op=dist.ReduceOp.SUM, group=process_group)
                a.phase = 3
                _lib.check(lib.srx_bake_update(C.byref(a), stream))
            elif sharded:
                a.phase = 1
                _lib.check(lib.srx_bake_update(C.byref(a), stream))
                ntex = self.k * self.k * self.height * self.width
                # RGB into an RGBA atlas: the weight sums live in the accumulator's alpha channel (csrc/srx_bake.cu), the
                # separate [ntex] weight array stays zero and is not exchanged
                nbytes = ntex * 16 if (self.channel_count == 4 and colors.shape[-1] == 3) else need - 256
                dist.all_reduce(self._workspace[:nbytes].view(torch.float32), op=dist.ReduceOp.SUM, group=process_group)
                a.phase = 2
                _lib.check(lib.srx_bake_update(C.byref(a), stream))
            else:
                a.phase = int(phase)
                _lib.check(lib.srx_bake_update(C.byref(a), stream))

    def check(self) -> None:
        """Raises IndexError if an `update(..., defer_check=True)` since the last check addressed a texel outside the atlas
        (syncs the current stream)."""
        a = getattr(self, "_last_args", None)
        if a is None:
            return
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().srx_bake_check(C.byref(a), _lib.current_stream_ptr(self.device)))

    @staticmethod
    def _stack(frames, what: str) -> Tensor:
        if isinstance(frames, Tensor):
            if frames.dim() == 4:
                return frames
            if frames.dim() == 3:
                return frames.unsqueeze(0)
            raise ValueError(f"The shape of {what} is invalid. Got: ", frames.shape)
        if isinstance(frames, (list, tuple)):
            parts = []
            for f in frames:
                if f.dim() == 4:
                    parts.append(f)
                elif f.dim() == 3:
                    parts.append(f.unsqueeze(0))
                else:
                    raise ValueError(f"The shape of {what} is invalid. Got: ", f.shape)
            return torch.cat(parts, dim=0)
        raise ValueError(f"Invalid type of {what}. Got: ", type(frames))

    # -- on-disk format (reference corrmap.py:738-872) -------------------------------------------------------------
    def dump(self, path, name: Optional[str] = None, zip: bool = False, force: bool = False):
        """`CorrespondMap.dump`: k*k `{i}.png` (uint8 = clip(255*value)) + `{i}_written.png` + `meta.json` in a folder
        (or a zip).  Same naming rules as the reference: without `force` an existing name gets a `_N` suffix, with
        `force` the old folder / file is removed first.  The quantisation runs on the GPU; only bytes are copied back."""
        import json
        import shutil
        import tempfile
        import zipfile
        from PIL import Image
        name = name or self.name
        real_name = name
        suffix = ".zip" if zip else ""
        path = str(path)
        if not force:
            count = 1
            while os.path.exists(os.path.join(path, real_name + suffix)):
                real_name = f"{name}_{count}"
                count += 1
        else:
            old = os.path.join(path, real_name + suffix)
            if os.path.exists(old):
                if os.path.isdir(old):
                    if zip:
                        raise ValueError(f"Folder with the same name {real_name + suffix} already exists in {path}. "
                                         "It is not allowed to delete a whole folder in zip mode.")
                    shutil.rmtree(old)
                else:
                    os.remove(old)
        real_path = os.path.join(path, real_name + suffix)
        working = tempfile.mkdtemp(prefix="srx_corrmap_") if zip else real_path
        os.makedirs(path, exist_ok=True)
        os.makedirs(working, exist_ok=True)
        k2, texels, ch = self.k * self.k, self.height * self.width, self.channel_count
        q = torch.empty(k2 * texels * ch, dtype=torch.uint8, device=self.device)
        fl = torch.empty(k2 * texels, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().srx_atlas_quantize(self._values.data_ptr(), self._writtens.data_ptr(), q.data_ptr(), fl.data_ptr(),
                                                      q.numel(), fl.numel(), _lib.current_stream_ptr(self.device)))
        q = q.view(k2, self.height, self.width, ch).cpu().numpy()
        fl = fl.view(k2, self.height, self.width).cpu().numpy()
        mode_map = {1: "L", 3: "RGB", 4: "RGBA"}
        for i in range(k2):
            img = q[i, :, :, 0] if ch == 1 else q[i]
            Image.fromarray(img, mode=mode_map[ch]).save(os.path.join(working, f"{i}.png"))
            Image.fromarray(fl[i], mode="L").save(os.path.join(working, f"{i}_written.png"))
        with open(os.path.join(working, "meta.json"), "w") as f:
            json.dump({"k": self.k, "height": self.height, "width": self.width, "channel_count": ch, "name": name}, f)
        if zip:
            with zipfile.ZipFile(real_path, "w") as z:
                for i in range(k2):
                    z.write(os.path.join(working, f"{i}.png"), f"{i}.png")
                    z.write(os.path.join(working, f"{i}_written.png"), f"{i}_written.png")
                z.write(os.path.join(working, "meta.json"), "meta.json")
            shutil.rmtree(working)
        return real_path

    @classmethod
    def Load(cls, path, name: Optional[str] = None, device=None):
        """`CorrespondMap.Load`: folder or zip written by `dump` (values come back as uint8 / 255 in float16)."""
        import json
        import shutil
        import tempfile
        import zipfile
        from PIL import Image
        path = str(path)
        tmp = None
        real_path = path
        if os.path.isfile(path):
            tmp = tempfile.mkdtemp(prefix="srx_corrmap_")
            with zipfile.ZipFile(path, "r") as z:
                z.extractall(tmp)
            real_path = tmp
        try:
            with open(os.path.join(real_path, "meta.json")) as f:
                meta = json.load(f)
            kw = {} if device is None else {"device": device}
            m = cls(name=name or meta["name"], k=meta["k"], height=meta["height"], width=meta["width"],
                    channel_count=meta["channel_count"], **kw)
            k2 = m.k * m.k
            vals = np.stack([np.array(Image.open(os.path.join(real_path, f"{i}.png"))).reshape(m.height, m.width, m.channel_count)
                             for i in range(k2)])
            flags = np.stack([np.array(Image.open(os.path.join(real_path, f"{i}_written.png"))) for i in range(k2)])
        finally:
            if tmp is not None:
                shutil.rmtree(tmp)
        vq = torch.from_numpy(np.ascontiguousarray(vals, dtype=np.uint8)).to(m.device)
        fq = torch.from_numpy(np.ascontiguousarray(flags, dtype=np.uint8)).to(m.device)
        with torch.cuda.device(m.device):
            _lib.check(_lib.load().srx_atlas_dequantize(vq.data_ptr(), fq.data_ptr(), m._values.data_ptr(), m._writtens.data_ptr(),
                                                        vq.numel(), fq.numel(), _lib.current_stream_ptr(m.device)))
        return m

    # -- atlas -> GL_TEXTURE_2D_ARRAY (reference CorrespondMap.load / set_data, corrmap.py:443-489) --------------------
    def load(self, texture=None, arrays=None, transpose: bool = True):
        """Uploads the k*k atlas layers into the `GL_TEXTURE_2D_ARRAY` the G-buffer shader samples
        (default_Gbuffer.frag.glsl:186-200) without leaving the GPU.  The reference round-trips every layer through the host
        (`get_map(i, order='whc').cpu().numpy()` -> `glTexSubImage3D`, corrmap.py:470-480); here one kernel per layer writes
        the layer's mapped cudaArray from the fp16 atlas, with the reference's 'whc' transpose (texel (x, y) =
        `_values[i, x * width + y]`) done through a shared-memory tile.

        texture: a `stable_renderer_b200.texture.Texture` wrapping the engine's GL_TEXTURE_2D_ARRAY (RGBA16F / RG16F / R16F
                 with k*k layers; RGB16F cannot be registered with CUDA, renderManager.py:268), or
        arrays:  a list of k*k `cudaArray_t` handles (one per layer) — tests and headless tools."""
        k2 = self.k * self.k
        if self.channel_count not in (1, 2, 4):
            raise NotImplementedError("CUDA arrays hold 1, 2 or 4 channels: a 3-channel atlas (RGB16F) has no CUDA mapping")
        lib = _lib.load()
        if (texture is None) == (arrays is None):
            raise ValueError("pass either the Texture of the GL array texture or the list of per-layer cudaArray handles")
        if arrays is not None and len(arrays) != k2:
            raise ValueError(f"{k2} layer arrays expected, got {len(arrays)}")
        layer_bytes = self.height * self.width * self.channel_count * 2
        try:
            for i in range(k2):
                arr = texture.map_array(layer=i) if texture is not None else int(arrays[i])
                with torch.cuda.device(self.device):
                    _lib.check(lib.srx_atlas_to_array(arr, self._values.data_ptr() + i * layer_bytes, self.height, self.width,
                                                      self.channel_count, 1 if transpose else 0, _lib.current_stream_ptr(self.device)))
        finally:
            if texture is not None:
                for _ in range(k2):
                    texture.unmap_array()

    def load_vertex_screen_info(self, id_map: IDMap):
        self.vertex_screen_info = id_map.create_vertex_screen_info()

    @property
    def unique_vertex_ids(self) -> Tensor:
        assert self.vertex_screen_info is not None, "Vertex screen positions are not loaded."
        return self.vertex_screen_info[..., 3].unique()


__all__ = ["IDMap", "UpdateMode", "CorrespondMap", "extract_index"]
