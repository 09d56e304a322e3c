"""OverlapPlan — thin owner of an `srx_plan` handle plus its torch-allocated workspace.

One plan serves one batch of id buffers (one `IDMap`) and one latent shape; it is what the reference caches as
`IDMap._vertex_screen_info_cache` (source/engine/static/corrmap.py:218-280), except that nothing per-entry is
materialised: keying happens inside the streaming kernel."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import _lib


class OverlapPlan:
    def __init__(self, ids: Optional[torch.Tensor], latent_shape: Sequence[int], *,
                 frame_indices: Optional[Sequence[int]] = None, id_shape: Optional[Sequence[int]] = None,
                 id_dtype: Optional[torch.dtype] = None, key_capacity: int = 0, deterministic: bool = False,
                 device: Optional[torch.device] = None, process_group=None, exchange: str = "auto",
                 split_kernels: bool = False):
        """process_group: frame-sharded run — every rank holds its own frames and the same key
        capacity.  exchange: "peer" = the accumulator exchange runs inside the step kernel over NVLink peer memory
        (workspace allocated as torch symmetric memory), "nccl" = reduce / all_reduce / gather as three launches,
        "auto" = peer when it can be set up, else nccl."""
        lib = _lib.load()
        if ids is not None:
            if not ids.is_cuda:
                raise _lib.SrxUnavailable("id buffers must live on a CUDA device (there is no CPU path)")
            if ids.dim() != 4 or ids.shape[-1] != 4:
                raise ValueError(f"id buffers must be [F,H,W,4], got {tuple(ids.shape)}")
            ids = ids.contiguous()
            id_shape, id_dtype, device = ids.shape, ids.dtype, ids.device
        elif id_shape is None or id_dtype is None or key_capacity <= 0 or device is None:
            raise ValueError("without ids, id_shape, id_dtype, device and key_capacity are required")
        F, H, W = int(id_shape[0]), int(id_shape[1]), int(id_shape[2])
        B, Cc, h, w = (int(v) for v in latent_shape)
        self.device = torch.device(device)
        self.id_shape = (F, H, W, 4)
        self.id_dtype = id_dtype
        self.latent_shape = (B, Cc, h, w)
        self.deterministic = bool(deterministic)
        self._ids = ids
        desc = _lib.srx_plan_desc()
        desc.id_dtype = _lib.torch_dtype_code(id_dtype)
        desc.frames, desc.height, desc.width = F, H, W
        desc.batch, desc.channels, desc.lat_h, desc.lat_w = B, Cc, h, w
        desc.key_mode = _lib.SRX_KEY_VERTEX
        desc.merge_len = 0
        desc.accum_mode = (_lib.SRX_ACCUM_DETERMINISTIC if deterministic else
                           _lib.SRX_ACCUM_FAST_SPLIT if split_kernels else _lib.SRX_ACCUM_FAST)
        desc.key_capacity = int(key_capacity)
        fm = None
        if frame_indices is not None:
            if len(frame_indices) != F:
                raise ValueError(f"frame_indices has {len(frame_indices)} entries for {F} id frames")
            fm = (C.c_int32 * F)(*[int(v) for v in frame_indices])
            desc.frame_map = C.cast(fm, C.POINTER(C.c_int32))
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            stream = _lib.current_stream_ptr(self.device)
            _lib.check(lib.srx_plan_create(C.byref(self._handle), C.byref(desc), ids.data_ptr() if ids is not None else None,
                                           stream))
            info = _lib.srx_plan_info()
            _lib.check(lib.srx_plan_get_info(self._handle, C.byref(info)))
            self.info = info
            self.key_capacity_hint = int(info.key_capacity)
            self.nvls = False
            self.fused = bool(info.fused)
            self.group = process_group
            self.world = 1
            self.exchange = "none"
            self._symm = None
            nbytes = int(info.workspace_bytes)
            if process_group is not None:
                import torch.distributed as dist
                self.world = dist.get_world_size(process_group)
                self.exchange = "nccl"
            self._want_nvls = exchange == "nvls" or bool(os.environ.get("SRX_NVLS"))
            if exchange == "nvls":
                exchange = "peer"
            want_peer = self.world > 1 and exchange in ("auto", "peer")
            if exchange == "peer" and self.world > 1 and not self.fused:
                raise _lib.SrxError("exchange='peer' needs the persistent step kernel (8x8 pixels per cell, 4 channels, "
                                    "non-deterministic accumulators)")
            self.workspace = None
            if want_peer and self.fused:
                self.workspace = self._alloc_symmetric(nbytes, exchange == "peer")
            if self.workspace is None:
                self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            _lib.check(lib.srx_plan_bind_workspace(self._handle, self.workspace.data_ptr(), nbytes, stream))
            if self._symm is not None:
                self._bind_peers(lib)
        self.key_capacity = int(info.key_capacity)
        self.n_valid = int(info.n_valid)
        self.fast_path = bool(info.fast_path)

    # -- frame-sharded peer mode -----------------------------------------------------------------------------
    def _alloc_symmetric(self, nbytes: int, required: bool) -> Optional[torch.Tensor]:
        """Workspace as torch symmetric memory (plumbing only: it hands every rank the peers' device pointers).
        All ranks of the group must succeed or fail together, so the outcome is agreed with one all_reduce."""
        import torch.distributed as dist
        ws, err = None, None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            ws = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._symm = symm_mem.rendezvous(ws, group=self.group)
        except Exception as e:  # noqa: BLE001 - any failure means "no peer mapping on this box"
            ws, err, self._symm = None, e, None
        ok = torch.tensor([1 if ws is not None else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            self._symm = None
            if required:
                raise _lib.SrxError(f"peer exchange unavailable: symmetric memory rendezvous failed ({err})")
            return None
        return ws

    def _bind_peers(self, lib) -> None:
        import torch.distributed as dist
        rank = dist.get_rank(self.group)
        ptrs = [int(v) for v in self._symm.buffer_ptrs]
        assert len(ptrs) == self.world
        arr = (C.c_void_p * self.world)(*ptrs)
        torch.cuda.synchronize(self.device)          # the clears issued by bind_workspace have landed ...
        dist.barrier(group=self.group)               # ... on every rank before anyone signals a peer
        _lib.check(lib.srx_plan_bind_peers(self._handle, rank, self.world, arr))
        self.exchange = "peer"
        # NVLS (opt-in: SRX_NVLS=1, or exchange="nvls"): when the symmetric allocation has a multicast mapping on EVERY rank,
        # the in-kernel exchange reduces inside the NVSwitch (multimem.ld_reduce) and broadcasts the totals (multimem.st)
        # instead of pulling world-1 slices.  Measured on 8 x B200 (profiles/r2_exchange_nvls_vs_pull.txt) it is correct but
        # not faster than the pull exchange (N=8: 105 vs 101 us per cfg3 step; N=4: 166 vs 145; N=2: 247 vs 236): the
        # in-switch reduction returns after ~9 us and the broadcast of all non-empty slots moves more bytes than the pull
        # of the needed records — so the pull form stays the default.
        self.nvls = False
        mc = 0
        if self._want_nvls and not os.environ.get("SRX_NO_NVLS"):
            try:
                mc = int(self._symm.multicast_ptr or 0)
            except Exception:  # noqa: BLE001 - no multicast support in this torch build / on this fabric
                mc = 0
            if self.key_capacity_hint % 4 != 0 or self.key_capacity_hint > (4 << 20):
                mc = 0
        ok = torch.tensor([1 if mc else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 1:
            _lib.check(lib.srx_plan_bind_multicast(self._handle, C.c_void_p(mc)))
            self.nvls = True

    def bind_peers(self, rank: int, peer_workspaces: Sequence[int]) -> None:
        """Low-level peer binding from raw workspace device pointers (tests emulate two ranks on one GPU with it)."""
        world = len(peer_workspaces)
        arr = (C.c_void_p * world)(*[int(v) for v in peer_workspaces])
        _lib.check(_lib.load().srx_plan_bind_peers(self._handle, int(rank), world, arr))
        self.world, self.exchange = world, ("peer" if world > 1 else "none")

    def read_trace(self) -> dict:
        """Phase durations (microseconds at `sm_mhz`-independent SM ticks -> caller scales) of the last persistent step."""
        buf = (C.c_int64 * 16)()
        _lib.check(_lib.load().srx_plan_read_trace(self._handle, buf, _lib.current_stream_ptr(self.device)))
        v = list(buf)
        return {"first_cta": v[:8], "last_cta": v[8:]}

    def read_step_ring(self) -> list:
        buf = (C.c_uint64 * 320)()
        _lib.check(_lib.load().srx_plan_read_step_ring(self._handle, buf, _lib.current_stream_ptr(self.device)))
        return [(int(buf[2 * i]), int(buf[2 * i + 1]), [int(buf[64 + 8 * i + j]) for j in range(8)]) for i in range(32)]

    def set_grid(self, ctas: int) -> None:
        _lib.check(_lib.load().srx_plan_set_grid(self._handle, int(ctas)))

    # -- accumulator view for the multi-GPU exchange ---------------------------------------
    @property
    def accumulator(self) -> torch.Tensor:
        """The key-indexed accumulator [K*(C+1)] (float32; int64 in deterministic mode): K*C sums then K counts."""
        dt = torch.int64 if self.deterministic else torch.float32
        raw = self.workspace[int(self.info.accum_offset): int(self.info.accum_offset + self.info.accum_bytes)]
        return raw.view(dt)

    def _args(self, x: torch.Tensor, ids: Optional[torch.Tensor], ratio: float, adain: bool,
              need_ids: bool = True) -> _lib.srx_step_args:
        if tuple(x.shape) != self.latent_shape:
            raise ValueError(f"latents {tuple(x.shape)} do not match the plan {self.latent_shape}")
        if not x.is_cuda or x.device != self.device:
            raise _lib.SrxUnavailable("latents must live on the plan's CUDA device (there is no CPU path)")
        if not x.is_contiguous():
            raise ValueError("latents must be contiguous")
        a = _lib.srx_step_args()
        a.x_dev = x.data_ptr()
        a.x_dtype = _lib.torch_dtype_code(x.dtype)
        a.ids_dev = None
        if need_ids:
            ids = self._ids if ids is None else ids
            if ids is None:
                raise ValueError("this plan was built without ids: pass them to every call")
            if (tuple(ids.shape) != self.id_shape or ids.dtype != self.id_dtype or ids.device != self.device
                    or not ids.is_contiguous()):
                raise ValueError("id buffers do not match the plan (shape / dtype / device / contiguity)")
            a.ids_dev = ids.data_ptr()
        a.ratio = float(ratio)
        a.adain = 1 if adain else 0
        a.cache_slots = 0
        return a

    def build_cache(self, ids: Optional[torch.Tensor] = None) -> None:
        """Bucketing pass (cached-plan regime): two streaming passes over the ids store, per CTA, the (key, cell,
        multiplicity) pairs of the keys that win a cell; later `step(..., cached=True)` calls run from that pool without
        touching the ids — the regime of denoise steps 2..N of one sampling run.  Syncs once (pool sizing)."""
        ids = self._ids if ids is None else ids
        if ids is None:
            raise ValueError("this plan was built without ids: pass them")
        if (tuple(ids.shape) != self.id_shape or ids.dtype != self.id_dtype or ids.device != self.device
                or not ids.is_contiguous()):
            raise ValueError("id buffers do not match the plan (shape / dtype / device / contiguity)")
        self.cache_mark(ids)
        if self.world > 1 and self.group is not None:
            # a key matters if it wins a cell on ANY rank: its mean needs every rank's contributions
            import torch.distributed as dist
            dist.all_reduce(self.need_map, op=dist.ReduceOp.MAX, group=self.group)
        self.cache_emit(ids)

    @property
    def need_map(self) -> torch.Tensor:
        """[key_capacity] uint8: 1 where the key wins a cell (valid after cache_mark)."""
        o = int(self.info.need_offset)
        return self.workspace[o:o + self.key_capacity]

    def cache_mark(self, ids: torch.Tensor) -> None:
        lib = _lib.load()
        nbytes = C.c_int64(0)
        with torch.cuda.device(self.device):
            _lib.check(lib.srx_plan_build_cache(self._handle, ids.data_ptr(), None, C.byref(nbytes), _lib.current_stream_ptr(self.device)))
            self._pool = torch.empty(int(nbytes.value), dtype=torch.uint8, device=self.device)

    def cache_emit(self, ids: torch.Tensor) -> None:
        nbytes = C.c_int64(self._pool.numel())
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().srx_plan_build_cache(self._handle, ids.data_ptr(), self._pool.data_ptr(), C.byref(nbytes),
                                                        _lib.current_stream_ptr(self.device)))
        self.cached = True

    def cache_entries(self) -> tuple:
        kept, cap = C.c_int64(0), C.c_int64(0)
        _lib.check(_lib.load().srx_plan_cache_entries(self._handle, C.byref(kept), C.byref(cap), _lib.current_stream_ptr(self.device)))
        return int(kept.value), int(cap.value)

    def step(self, x: torch.Tensor, ratio: float, *, adain: bool = True, ids: Optional[torch.Tensor] = None,
             cached: bool = False) -> None:
        """One overlap step in place on x (reduce + gather on this GPU; with a peer group also the exchange).
        cached=True runs from the pool of `build_cache` instead of streaming the ids."""
        if cached:
            if not getattr(self, "cached", False):
                raise _lib.SrxError("step(cached=True) needs build_cache() first")
            a = self._args(x, None, ratio, adain, need_ids=False)
            _lib.check(_lib.load().srx_overlap_step(self._handle, C.byref(a), _lib.current_stream_ptr(self.device)))
            return
        a = self._args(x, ids, ratio, adain)
        _lib.check(_lib.load().srx_overlap_step(self._handle, C.byref(a), _lib.current_stream_ptr(self.device)))

    def reduce(self, x: torch.Tensor, *, ids: Optional[torch.Tensor] = None) -> None:
        a = self._args(x, ids, 0.0, True)
        _lib.check(_lib.load().srx_accum_reduce(self._handle, C.byref(a), _lib.current_stream_ptr(self.device)))

    def gather(self, x: torch.Tensor, ratio: float, *, adain: bool = True) -> None:
        a = self._args(x, None, ratio, adain, need_ids=False)
        _lib.check(_lib.load().srx_accum_finalize_gather(self._handle, C.byref(a), _lib.current_stream_ptr(self.device)))

    def check(self) -> None:
        """Raises for device-side failures of earlier launches (syncs)."""
        _lib.check(_lib.load().srx_plan_check(self._handle, _lib.current_stream_ptr(self.device)))

    def close(self) -> None:
        if getattr(self, "_handle", None) is not None and self._handle:
            _lib.load().srx_plan_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
