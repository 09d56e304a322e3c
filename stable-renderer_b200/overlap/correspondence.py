"""`CorrespondenceMap` of the legacy generation
(reference: legacy_codes/stable_rendering_algo/data_classes/correspondence_map.py:25-286).

The reference builds a Python dict {id tuple -> [([row, col], frame), ...]} with a triple loop over every pixel and
pickles it.  Here the map IS the id buffers on the GPU plus the merge distance: the kernels key on the id tuple
directly (exact 64-bit packing), so nothing is built, cached or pickled.  `Map` materialises the reference's dict on
demand for inspection."""
from __future__ import annotations

import os
import re
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch


class CorrespondenceMap:
    def __init__(self, ids: torch.Tensor, merge_len: int = 0, num_frames: Optional[int] = None):
        if ids.dim() != 4 or ids.shape[-1] != 4:
            raise ValueError(f"id buffers must be [T,H,W,4], got {tuple(ids.shape)}")
        if num_frames is not None:
            ids = ids[:num_frames]
        if ids.dtype not in (torch.int16, torch.int32):
            ids = ids.to(torch.int32)
        self._ids = ids.contiguous()
        self._merge_len = int(merge_len)
        self._device_ids: dict = {}

    # -- reference properties (correspondence_map.py:54-71) --------------------------------------------------------
    @property
    def width(self) -> int:
        return int(self._ids.shape[2])

    @property
    def height(self) -> int:
        return int(self._ids.shape[1])

    @property
    def size(self) -> tuple:
        return (self.width, self.height)

    @property
    def num_frames(self) -> int:
        return int(self._ids.shape[0])

    @property
    def merge_len(self) -> int:
        return self._merge_len

    @property
    def ids(self) -> torch.Tensor:
        return self._ids

    def device_ids(self, device) -> torch.Tensor:
        device = torch.device(device)
        t = self._device_ids.get(device)
        if t is None:
            t = self._ids.to(device).contiguous()
            self._device_ids[device] = t
        return t

    def merge_nearby(self, distance: int):
        """Key -> (obj, mat, texX // d, texY // d) (correspondence_map.py:276-286); successive calls compose."""
        d = int(distance)
        if d > 1:
            self._merge_len = max(self._merge_len, 1) * d

    # -- dropouts and cache (correspondence_map.py:177-274) ----------------------------------------------------------------
    def _device(self) -> torch.device:
        if self._ids.is_cuda:
            return self._ids.device
        from .. import _lib
        if not torch.cuda.is_available():
            raise _lib.SrxUnavailable("CorrespondenceMap maintenance runs on the GPU (there is no CPU path)")
        self._ids = self._ids.cuda()
        self._device_ids = {}
        return self._ids.device

    def _first_appearance(self) -> torch.Tensor:
        """Linear pixel indices (frame, row, col order) that introduce a key: the reference dict's keys in insertion order."""
        import ctypes as C
        from .. import _lib
        dev = self._device()
        lib = _lib.load()
        F, H, W = self.num_frames, self.height, self.width
        npx = F * H * W
        mask = torch.empty(npx, dtype=torch.uint8, device=dev)
        ws = torch.empty(int(lib.srx_corrmap_keys_workspace_bytes(npx)), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.srx_corrmap_first_appearance(self._ids.data_ptr(), _lib.torch_dtype_code(self._ids.dtype), F, H, W,
                                                        self._merge_len, mask.data_ptr(), ws.data_ptr(), ws.numel(),
                                                        _lib.current_stream_ptr(dev)))
        return torch.nonzero(mask).flatten()

    def _drop(self, seeds: torch.Tensor) -> None:
        import ctypes as C
        from .. import _lib
        dev = self._device()
        lib = _lib.load()
        seeds = seeds.to(device=dev, dtype=torch.int64).contiguous()
        n = int(seeds.numel())
        if n == 0:
            return
        ws = torch.empty(int(lib.srx_corrmap_keys_workspace_bytes(n)), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.srx_corrmap_drop_keys(self._ids.data_ptr(), _lib.torch_dtype_code(self._ids.dtype), self.num_frames,
                                                 self.height, self.width, self._merge_len, seeds.data_ptr(), n, ws.data_ptr(),
                                                 ws.numel(), _lib.current_stream_ptr(dev)))
        self._device_ids = {}

    def dropout_index(self, probability: float, seed: int):
        """`dropout_index` (correspondence_map.py:207-223): every key is deleted with `probability`; the decisions come from
        Python's `random` seeded with `seed`, one draw per key in the dict's insertion order — replayed here over the
        first-appearance order of the keys.  Irreversible, like the reference."""
        import random
        assert 0 <= probability <= 1
        first = self._first_appearance()
        random.seed(seed)
        flags = [random.random() < probability for _ in range(int(first.numel()))]
        self._drop(first[torch.tensor(flags, dtype=torch.bool, device=first.device)] if flags else first[:0])

    def dropout_in_rectangle(self, rectangle, at_frame: int):
        """`dropout_in_rectangle` (correspondence_map.py:225-274): deletes every key that has a pixel strictly inside the
        rectangle `((r0, c0), (r1, c1))` (or an object with `.top_left` / `.bottom_right`) at frame `at_frame`."""
        assert 0 <= at_frame <= self.num_frames
        if isinstance(rectangle, tuple):
            top_left, bottom_right = rectangle
        elif hasattr(rectangle, "top_left") and hasattr(rectangle, "bottom_right"):
            top_left, bottom_right = rectangle.top_left, rectangle.bottom_right
        else:
            raise ValueError(f"Data type of {type(rectangle)} is not supported for rectangle")
        assert self.width >= max(top_left[0], bottom_right[0])
        assert self.height >= max(top_left[1], bottom_right[1])
        if at_frame >= self.num_frames:
            return
        dev = self._device()
        rows = torch.arange(max(top_left[0] + 1, 0), min(bottom_right[0], self.height), device=dev, dtype=torch.int64)
        cols = torch.arange(max(top_left[1] + 1, 0), min(bottom_right[1], self.width), device=dev, dtype=torch.int64)
        if rows.numel() == 0 or cols.numel() == 0:
            return
        seeds = (at_frame * self.height + rows.view(-1, 1)) * self.width + cols.view(1, -1)
        self._drop(seeds.reshape(-1))

    def save_cache(self, path: str):
        """`save_cache` (correspondence_map.py:194-205): pickles the map ('<dir>/corr_map.pkl' when a directory is given)."""
        import pickle
        path = str(path)
        if os.path.exists(path) and os.path.isdir(path):
            path = os.path.join(path, "corr_map.pkl")
        with open(path, "wb") as f:
            pickle.dump({"ids": self._ids.cpu(), "merge_len": self._merge_len}, f)

    @classmethod
    def LoadFromCache(cls, path: str, device=None) -> "CorrespondenceMap":
        """`LoadFromCache` (correspondence_map.py:177-192).  Reads this package's cache (the id buffers) and the REFERENCE's
        `corr_map.pkl` — a pickle of its dict-based `CorrespondenceMap` object (`save_cache`, :194-205): the dict
        {id tuple: [([row, col], frame), ...]} is turned back into `[F,H,W,4]` id buffers with array operations, so the plan
        builder is fed without the reference's Python triple loop (:148-168).  Only builtins and numpy scalars / arrays are
        unpickled; the reference class is mapped by name onto an attribute holder (no reference code runs).  A map that was merged
        before it was cached (`merge_nearby`) keeps its merged keys; its entries are then ordered by position, not by sub-trace."""
        import pickle
        path = str(path)
        if os.path.exists(path) and os.path.isdir(path):
            path = os.path.join(path, "corr_map.pkl")
        if not os.path.exists(path):
            raise FileNotFoundError(f"Correspondence map cache file not found at {path}")

        class _Holder:                                   # stands in for the pickled reference object
            pass

        class _Restricted(pickle.Unpickler):
            def find_class(self, module, name):
                if name == "CorrespondenceMap":
                    return _Holder
                root = module.split(".")[0]
                if root in ("builtins", "collections", "numpy", "torch", "_codecs", "copyreg"):
                    return super().find_class(module, name)
                raise pickle.UnpicklingError(f"refusing to unpickle {module}.{name}")
        with open(path, "rb") as f:
            d = _Restricted(f).load()
        if isinstance(d, dict) and "ids" in d:
            ids = d["ids"] if device is None else d["ids"].to(device)
            return cls(ids, merge_len=d["merge_len"])
        if not isinstance(d, _Holder) or not isinstance(getattr(d, "_correspondence_map", None), dict):
            raise ValueError(f"{path} holds neither this package's cache nor a pickled reference CorrespondenceMap")
        m = d._correspondence_map
        if not m:
            raise ValueError(f"{path}: empty correspondence map")
        keys = np.array([[int(c) for c in k] for k in m.keys()], dtype=np.int64)
        if keys.ndim != 2 or keys.shape[1] != 4:
            raise ValueError(f"{path}: id keys must be 4-tuples, got shape {keys.shape}")
        lens = np.fromiter((len(v) for v in m.values()), dtype=np.int64, count=len(m))
        flat = np.array([(p[0], p[1], f) for v in m.values() for (p, f) in v], dtype=np.int64).reshape(-1, 3)
        W, H, F = getattr(d, "_width", None), getattr(d, "_height", None), getattr(d, "_num_frames", None)
        H = int(H) if H else int(flat[:, 0].max()) + 1
        W = int(W) if W else int(flat[:, 1].max()) + 1
        F = int(F) if F else int(flat[:, 2].max()) + 1
        if flat.min() < 0 or flat[:, 0].max() >= H or flat[:, 1].max() >= W or flat[:, 2].max() >= F:
            raise ValueError(f"{path}: pixel positions outside the {F} x {H} x {W} id buffers (a position callback was used)")
        lo, hi = np.iinfo(np.int32).min, np.iinfo(np.int32).max
        if keys.min() < lo or keys.max() > hi:
            raise ValueError(f"{path}: id components outside int32")
        ids = np.zeros((F, H, W, 4), dtype=np.int32)
        ids[flat[:, 2], flat[:, 0], flat[:, 1]] = np.repeat(keys, lens, axis=0).astype(np.int32)
        t = torch.from_numpy(ids)
        return cls(t if device is None else t.to(device), num_frames=F)

    # -- construction ---------------------------------------------------------------------------------------------------
    @classmethod
    def from_ids(cls, ids: torch.Tensor, num_frames: Optional[int] = None, merge_len: int = 0) -> "CorrespondenceMap":
        return cls(ids, merge_len=merge_len, num_frames=num_frames)

    @classmethod
    def FromExisting(cls, directory: str, num_frames: Optional[int] = None, device=None, enable_cache: bool = True,
                     **_ignored) -> "CorrespondenceMap":
        """Loads the `*.npy` id dumps of a directory ordered by the first number in each file name
        (correspondence_map.py:122-143).  With `enable_cache` an existing `corr_map.pkl` — this package's or the reference's —
        is used instead, looked for where the reference looks (:103-116): the path itself, the directory, the parent of an `id`
        directory.  No cache is written (building from the dumps is one `np.stack`); the position callback does not apply."""
        directory = str(directory)
        if enable_cache:
            cache_path = None
            if os.path.isfile(directory) and directory.endswith(".pkl"):
                cache_path = directory
            elif os.path.isdir(directory) and "corr_map.pkl" in os.listdir(directory):
                cache_path = os.path.join(directory, "corr_map.pkl")
            elif os.path.isdir(directory) and directory.rstrip("/\\").endswith("id") and \
                    os.path.isfile(os.path.join(directory, "..", "corr_map.pkl")):
                cache_path = os.path.join(directory, "..", "corr_map.pkl")
            if cache_path is not None:
                return cls.LoadFromCache(cache_path, device=device)
        assert os.path.isdir(directory), f"{directory} is not a directory"
        if not directory.endswith("id") and "id" in os.listdir(directory):
            directory = os.path.join(directory, "id")
        items = []
        for name in os.listdir(directory):
            if not name.endswith(".npy"):
                continue
            m = re.search(r"\d+", name)
            if not m:
                raise RuntimeError(f"{name} has no numeric component in filename.")
            items.append((int(m.group()), name))
        items.sort()
        if num_frames is not None:
            items = items[:num_frames]
        if not items:
            raise FileNotFoundError(f"no id dumps in {directory}")
        ids = torch.from_numpy(np.stack([np.load(os.path.join(directory, n)) for _, n in items]))
        if device is not None:
            ids = ids.to(device)
        return cls(ids)

    # -- inspection (host side, not used by the kernels) ------------------------------------------------------------
    @property
    def Map(self) -> Dict[tuple, List[Tuple[List[int], int]]]:
        ids = self._ids.cpu().numpy().astype(np.int64)
        nz = np.any(ids != 0, axis=-1)
        f_idx, r_idx, c_idx = np.nonzero(nz)
        keys = ids[f_idx, r_idx, c_idx]
        d = max(self._merge_len, 1)
        if d > 1:
            keys = keys.copy()
            keys[:, 2] //= d
            keys[:, 3] //= d
        out: Dict[tuple, List[Tuple[List[int], int]]] = {}
        for k, f, r, c in zip(map(tuple, keys.tolist()), f_idx.tolist(), r_idx.tolist(), c_idx.tolist()):
            out.setdefault(k, []).append(([r, c], f))
        return out

    def __len__(self) -> int:
        ids = self._ids.reshape(-1, 4).to(torch.int64)
        ids = ids[(ids != 0).any(dim=1)]
        d = max(self._merge_len, 1)
        if d > 1:
            ids = torch.stack([ids[:, 0], ids[:, 1], ids[:, 2] // d, ids[:, 3] // d], dim=1)
        return int(torch.unique(ids, dim=0).shape[0])
