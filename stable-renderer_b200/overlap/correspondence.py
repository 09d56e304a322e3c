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

    # -- construction ---------------------------------------------------------------------------------------------------
    @classmethod
    def from_ids(cls, ids: torch.Tensor, num_frames: Optional[int] = None, merge_len: int = 0) -> "CorrespondenceMap":
        return cls(ids, merge_len=merge_len, num_frames=num_frames)

    @classmethod
    def FromExisting(cls, directory: str, num_frames: Optional[int] = None, device=None, **_ignored) -> "CorrespondenceMap":
        """Loads the `*.npy` id dumps of a directory ordered by the first number in each file name
        (correspondence_map.py:122-143).  The reference's pickle cache and position callback do not apply."""
        directory = str(directory)
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
