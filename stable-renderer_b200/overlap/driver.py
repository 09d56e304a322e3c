"""`Overlap` and `ResizeOverlap` (reference: legacy_codes/stable_rendering_algo/overlap/overlap.py:18-222).

Same constructor and call signatures; one `srx_legacy_overlap` call replaces the per-vertex Python loop.
With kernel_radius > 0 the reference is an in-place (Gauss–Seidel) update in dict order (overlap.py:97,136-145):
`srx_legacy_overlap_ordered` walks the traces in that order on the GPU (entries of a trace in parallel, traces in
sequence) — see DESIGN.md §3."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from .. import _lib
from .algorithms import OverlapAlgorithm
from .correspondence import CorrespondenceMap
from .scheduler import Scheduler


class Overlap:
    def __init__(self, alpha_scheduler: Scheduler, kernel_radius_scheduler: Scheduler, algorithm: OverlapAlgorithm,
                 verbose: bool = True):
        self._verbose = verbose
        self.algorithm = algorithm
        self.alpha_scheduler = alpha_scheduler
        self.kernel_radius_scheduler = kernel_radius_scheduler
        self._workspace: Optional[torch.Tensor] = None
        self.defer_check = False
        '''True: calls only enqueue kernels (no host sync); ids that do not fit the packed key are reported by `check()`'''
        self._last = None

    @property
    def verbose(self):
        return self._verbose

    @verbose.setter
    def verbose(self, value: bool):
        self._verbose = value

    # ------------------------------------------------------------------------------------------------------------------
    def _run(self, stack: torch.Tensor, corr_map: CorrespondenceMap, alpha: float, view_normal_map, radius: int = 0) -> None:
        """stack [T, B*C, h, w] contiguous CUDA tensor, updated in place."""
        if not stack.is_cuda:
            raise _lib.SrxUnavailable("latents must be CUDA tensors (there is no CPU path)")
        lib = _lib.load()
        strategy = getattr(self.algorithm, "strategy", None)
        if strategy not in _lib.SRX_STRATEGY:
            raise ValueError(f"Unknown algorithm {strategy}")
        ids = corr_map.device_ids(stack.device)
        T, CC, h, w = stack.shape
        if ids.shape[0] < T:
            raise ValueError(f"correspondence map has {ids.shape[0]} frames, latents {T}")
        ids = ids[:T]
        d = _lib.srx_legacy_desc()
        d.id_dtype = _lib.torch_dtype_code(ids.dtype)
        d.frames, d.height, d.width = T, ids.shape[1], ids.shape[2]
        d.channels, d.lat_h, d.lat_w = CC, h, w
        d.merge_len = corr_map.merge_len
        d.strategy = _lib.SRX_STRATEGY[strategy]
        need = int((lib.srx_legacy_ordered_workspace_bytes if radius > 0 else lib.srx_legacy_workspace_bytes)(C.byref(d)))
        if need < 0:
            _lib.check(_lib.SRX_ERR_INVALID)
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != stack.device:
            self._workspace = torch.zeros(need, dtype=torch.uint8, device=stack.device)
        a = _lib.srx_legacy_args()
        a.defer_status = 1 if (self.defer_check and radius == 0) else 0
        a.x_dev, a.x_dtype = stack.data_ptr(), _lib.torch_dtype_code(stack.dtype)
        a.ids_dev = ids.data_ptr()
        a.alpha = float(alpha)
        vn = None
        if strategy == "perpendicular_view_normal":
            if view_normal_map is None:
                raise TypeError("overlap() missing 1 required positional argument: 'view_normal_map'")
            vn = view_normal_map.reshape(view_normal_map.shape[0], ids.shape[1], ids.shape[2])[:T]
            vn = vn.to(device=stack.device, dtype=torch.float32).contiguous()
            a.view_normal_dev = vn.data_ptr()
        a.workspace_dev, a.workspace_bytes = self._workspace.data_ptr(), self._workspace.numel()
        with torch.cuda.device(stack.device):
            stream = _lib.current_stream_ptr(stack.device)
            if radius > 0:
                _lib.check(lib.srx_legacy_overlap_ordered(C.byref(d), C.byref(a), int(radius), stream))
            else:
                _lib.check(lib.srx_legacy_overlap(C.byref(d), C.byref(a), stream))
                if a.defer_status:
                    self._last = (d, a, stack.device)

    def check(self) -> None:
        """Raises for ids that did not fit the packed key in calls made with `defer_check = True` (syncs)."""
        if self._last is None:
            return
        d, a, dev = self._last
        with torch.cuda.device(dev):
            _lib.check(_lib.load().srx_legacy_check(C.byref(d), C.byref(a), _lib.current_stream_ptr(dev)))

    def _schedule(self, step, timestep):
        """(alpha, kernel_radius) for this call (overlap.py:100-101: the radius is `int()`-truncated)."""
        alpha = self.alpha_scheduler(step, timestep)
        radius = int(self.kernel_radius_scheduler(step, timestep))
        return float(alpha), max(radius, 0)

    @torch.no_grad()
    def __call__(self, frame_seq: List[torch.Tensor], corr_map: CorrespondenceMap, step: int = None,
                 timestep: int = None, apply_corr_map_decay: bool = False, **kwargs) -> torch.Tensor:
        """frame_seq: T tensors [B,C,H,W] at the correspondence map's resolution -> stack [T,B,C,H,W] (overlap.py:83-152)."""
        assert frame_seq[0].shape[2:] == (corr_map.height, corr_map.width), \
            f"frame shape {frame_seq[0].shape[2:]} does not match corr_map shape {(corr_map.height, corr_map.width)}"
        alpha, radius = self._schedule(step, timestep)
        stack = torch.stack(frame_seq, dim=0).contiguous()      # [T,B,C,H,W], a fresh tensor like the reference's
        T, B, Cc, H, W = stack.shape
        self._run(stack.view(T, B * Cc, H, W), corr_map, alpha, kwargs.get("view_normal_map"), radius)
        return stack


class ResizeOverlap(Overlap):
    def __init__(self, alpha_scheduler: Scheduler, kernel_radius_scheduler: Scheduler, algorithm: OverlapAlgorithm,
                 verbose: bool = True, interpolate_mode: str = "nearest"):
        super().__init__(alpha_scheduler, kernel_radius_scheduler, algorithm, verbose)
        self._interpolate_mode = interpolate_mode

    @property
    def interpolate_mode(self):
        return self._interpolate_mode

    @interpolate_mode.setter
    def interpolate_mode(self, value: str):
        self._interpolate_mode = value

    @torch.no_grad()
    def __call__(self, frame_seq: List[torch.Tensor], corr_map: CorrespondenceMap, step: int = None,
                 timestep: int = None, **kwargs) -> List[torch.Tensor]:
        """frame_seq: T latents [B,C,h,w]; returns T overlapped latents (overlap.py:180-222).  The up-sample /
        overlap / down-sample / where() chain is evaluated directly on the latent cells."""
        alpha = self.alpha_scheduler(step, timestep)
        if alpha == 0:
            return frame_seq                                    # overlap.py:200-201
        if self._interpolate_mode != "nearest":
            return self._call_interpolated(frame_seq, corr_map, step, timestep, **kwargs)
        alpha, radius = self._schedule(step, timestep)
        stack = torch.stack(frame_seq, dim=0).contiguous()      # [T,B,C,h,w]
        T, B, Cc, h, w = stack.shape
        self._run(stack.view(T, B * Cc, h, w), corr_map, alpha, kwargs.get("view_normal_map"), radius)
        return list(stack.unbind(0))


    def _call_interpolated(self, frame_seq, corr_map, step, timestep, **kwargs) -> List[torch.Tensor]:
        """interpolate_mode != 'nearest' (overlap.py:205-221): the latents are really resampled — up to the map size with
        `F.interpolate`, overlapped there at full resolution (the same kernels with latent size = map size), resampled back, and
        merged with `torch.where(ovlp != 0, ovlp, original)`.  The nearest mode never builds the 64x larger tensors; the smooth
        modes have no such shortcut, every map pixel carries its own interpolated value."""
        import torch.nn.functional as F
        mode = self._interpolate_mode
        align_corners = False if mode in ("linear", "bilinear", "bicubic", "trilinear") else None
        screen_w, screen_h = corr_map.size
        frame_h, frame_w = frame_seq[0].shape[-2:]
        up = [F.interpolate(latents, size=(screen_h, screen_w), mode=mode, align_corners=align_corners) for latents in frame_seq]
        stack = Overlap.__call__(self, up, corr_map, step=step, timestep=timestep, **kwargs)
        down = [F.interpolate(latents, size=(frame_h, frame_w), mode=mode, align_corners=align_corners) for latents in stack]
        return [torch.where(down[i] != 0, down[i], frame_seq[i]) for i in range(len(frame_seq))]
