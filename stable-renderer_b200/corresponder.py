"""Drop-in `Corresponder` objects for the sampler loop
(reference: source/common_utils/stable_render_utils/corresponder.py:29-376).

`OverlapCorresponder.step_finished(engine_data, sampling_context)` keeps the reference's contract — it mutates
`sampling_context.noise` in place and returns None — but runs as two CUDA kernels on the current stream instead of
`clone / unique / scatter_add_ / index_put_ / var / mean`.  `DefaultCorresponder.finished` bakes the decoded
frames into the `CorrespondMap` atlases on the GPU."""
from __future__ import annotations

from typing import TYPE_CHECKING, Any, Optional, Protocol

import torch
from attr import attrib, attrs

from . import _lib
from .corrmap import IDMap, UpdateMode
from .plan import OverlapPlan

if TYPE_CHECKING:  # duck-typed: EngineData / SamplingCallbackContext are the host application's classes
    EngineData = Any
    SamplingCallbackContext = Any
    IMAGE = Any


class Corresponder(Protocol):
    """Same five hooks as the reference protocol (corresponder.py:29-98); callers probe them with hasattr."""

    def prepare(self, engine_data: "EngineData"): ...

    def pre_atten_inject(self, block, engine_data: "EngineData", q_context: torch.Tensor, k_context: torch.Tensor,
                         v_context: torch.Tensor, layer: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]: ...

    def post_atten_inject(self, block, engine_data: "EngineData", origin_values: torch.Tensor, layer: int) -> torch.Tensor: ...

    def step_finished(self, engine_data: "EngineData", sampling_context: "SamplingCallbackContext"): ...

    def finished(self, engine_data: "EngineData", images: "IMAGE"): ...


def _bake_all(corrmaps, images, id_maps: IDMap, mode, ignore_obj_mat_id, weight_mode="none", normal_depth=None):
    ids = id_maps.tensor
    masks = id_maps.masks   # IDMap.masks (1 = no id), not EngineData.masks — corresponder.py:139
    if masks is None:
        masks = torch.ones(ids.shape[:-1], dtype=torch.float32, device=ids.device)
    for (spriteID, materialID), corrmap in corrmaps.items():
        corrmap.update(color_frames=images, id_maps=ids, mode=mode, masks=masks, spriteID=spriteID,
                       materialID=materialID, ignore_obj_mat_id=ignore_obj_mat_id, inverse_masks=True,
                       weight_mode=weight_mode, normal_depth=normal_depth)


@attrs
class DefaultCorresponder:
    """Reference corresponder.py:100-155: bakes at `finished`; the attention hooks are inert."""

    layer_range: tuple = attrib(default=(6,))
    update_corrmap: bool = attrib(default=True)
    update_corrmap_mode: UpdateMode = attrib(default="first_avg")
    post_attn_inject_ratio: float = attrib(default=0.6)
    ignore_obj_mat_id_when_update: bool = attrib(default=False)
    bake_weight_mode: str = attrib(default="none")
    '''"none" = reference behaviour; "uniform" / "view_normal" / "view_normal_depth" = weighted multi-view bake'''

    def post_atten_inject(self, block, engine_data, origin_values: torch.Tensor, layer: int) -> torch.Tensor:
        return origin_values   # the reference returns before doing anything (corresponder.py:124)

    def finished(self, engine_data: "EngineData", images: "IMAGE"):
        if not self.update_corrmap or images is None or engine_data.id_maps is None:
            return
        corrmaps = engine_data.correspond_maps
        if corrmaps:
            _bake_all(corrmaps, images, engine_data.id_maps, self.update_corrmap_mode,
                      self.ignore_obj_mat_id_when_update, self.bake_weight_mode,
                      getattr(engine_data, "normal_maps", None) if self.bake_weight_mode.startswith("view") else None)


@attrs
class OverlapCorresponder:
    """Reference corresponder.py:157-376.

    Extra knobs (all default to the reference behaviour):
      key_capacity   dense slot count of the key table; 0 derives it from `engine_data.correspond_maps`
                     (height*width of the largest map) or, failing that, from one scan of the ids;
      deterministic  order-independent fixed-point accumulation (bit reproducible run to run);
      adain          False writes the blended latents instead of re-standardising the originals."""

    layer_range: tuple = attrib(default=(6,))
    update_corrmap: bool = attrib(default=True)
    update_corrmap_mode: UpdateMode = attrib(default="first")
    pre_attn_inject_num_random_frames: int = attrib(default=1)
    _random_frame_indices: torch.Tensor = attrib(default=None, init=False)
    post_attn_inject_ratio: float = attrib(default=0.6)
    step_finished_inject_ratio: float = attrib(default=0.1)
    step_finished_stop_inject_timestep: int = attrib(default=500)
    key_capacity: int = attrib(default=0, kw_only=True)
    deterministic: bool = attrib(default=False, kw_only=True)
    adain: bool = attrib(default=True, kw_only=True)
    process_group: Any = attrib(default=None, kw_only=True)
    exchange: str = attrib(default="auto", kw_only=True)
    enable_post_attn_inject: bool = attrib(default=False, kw_only=True)
    post_attn_map_size: Any = attrib(default=None, kw_only=True)
    '''None = the reference's (IDMap.height, IDMap.width) = (W, 4) for [F,H,W,4] ids; or an explicit (height, width)'''
    cache_plan: bool = attrib(default=True, kw_only=True)
    '''bucket the ids once per id batch and run later denoise steps from the cached (key, cell) pairs'''
    '''frame-sharded multi-GPU runs (SURVEY.md §8e): when set, every rank reduces its own frames into the key-indexed
    accumulator, the accumulators are summed across ranks, then each rank gathers its own frames.  Pass
    `torch.distributed.group.WORLD` (or True) for the default group.  `exchange`: "peer" = the sum runs inside the step
    kernel over NVLink peer memory, "nccl" = one NCCL all-reduce between two kernels, "auto" = peer when available.'''

    def prepare(self, engine_data: "EngineData"):
        pass

    def pre_atten_inject(self, block, engine_data, q_context, k_context, v_context, layer: int):
        """K/V replacement by a few random frames' contexts (corresponder.py:188-220).  Pure view/expand work on the
        attention side — not part of the scatter-reduce path; kept for interface parity."""
        if self.pre_attn_inject_num_random_frames < 0:
            return q_context, k_context, v_context
        if self._random_frame_indices is None:
            self._random_frame_indices = torch.randint(1, k_context.shape[0], (self.pre_attn_inject_num_random_frames,))
        idx = self._random_frame_indices.to(k_context.device)
        n = k_context.shape[0]
        k_new = k_context[idx].reshape(1, -1, k_context.shape[-1]).expand(n, -1, -1)
        v_new = v_context[idx].reshape(1, -1, v_context.shape[-1]).expand(n, -1, -1)
        return q_context, k_new, v_new

    def post_atten_inject(self, block, engine_data, origin_values: torch.Tensor, layer: int) -> torch.Tensor:
        """The reference returns `origin_values` before doing anything (corresponder.py:228) — the default here too.  With
        `enable_post_attn_inject=True` the body behind that return (:230-295) runs on the GPU: the wide-channel feature
        overlap of `feature.py` for layers above 10, blended with `post_attn_inject_ratio`."""
        if not self.enable_post_attn_inject:
            return origin_values
        from .feature import POST_ATTN_SKIP_LAYERS, feature_overlap
        if layer in POST_ATTN_SKIP_LAYERS:
            return origin_values
        return feature_overlap(origin_values, engine_data.id_maps, self.post_attn_inject_ratio,
                               map_size=self.post_attn_map_size, key_capacity=int(self.key_capacity))

    # -- the hot path ----------------------------------------------------------------------------------------------
    def _plan(self, engine_data, id_map: IDMap, x: torch.Tensor) -> OverlapPlan:
        key = (x.device, tuple(x.shape), self.deterministic, int(self.key_capacity), self.exchange,
               self.process_group is not None)
        plan = id_map._plans.get(key)
        if plan is None:
            cap = int(self.key_capacity)
            if cap <= 0:
                cms = getattr(engine_data, "correspond_maps", None)
                if cms:
                    cap = max(int(cm.height) * int(cm.width) for cm in cms.values())
            ids_dev = id_map.device_ids(x.device)
            if cap <= 0 and self.process_group is not None:
                # every rank must index the same slot table: agree on the largest vertex id once per id batch
                import torch.distributed as dist
                kmax = ids_dev[..., 3].max().to(torch.int64).reshape(1)
                dist.all_reduce(kmax, op=dist.ReduceOp.MAX, group=None if self.process_group is True else self.process_group)
                cap = int(kmax.item()) + 1
            group = None
            if self.process_group is not None:
                import torch.distributed as dist
                group = dist.group.WORLD if self.process_group is True else self.process_group
            plan = OverlapPlan(ids_dev, x.shape, frame_indices=id_map.frame_indices,
                               key_capacity=cap, deterministic=self.deterministic, process_group=group,
                               exchange=self.exchange)
            id_map._plans[key] = plan
        return plan

    def step_finished(self, engine_data: "EngineData", sampling_context: "SamplingCallbackContext"):
        timestep = sampling_context.timestep
        if timestep < self.step_finished_stop_inject_timestep:   # corresponder.py:299-303
            return
        id_map = engine_data.id_maps
        noise = sampling_context.noise
        if not noise.is_cuda:
            raise _lib.SrxUnavailable("sampling_context.noise must be a CUDA tensor (there is no CPU path)")
        x = noise if noise.is_contiguous() else noise.contiguous()
        plan = self._plan(engine_data, id_map, x)
        if self.process_group is None or plan.exchange == "peer":
            # One kernel, exchange included.  The ids of a sampling run do not change between denoise steps: once a
            # second step on the same id batch is certain, bucket them (two id passes) and run the remaining steps from
            # the cached pairs instead of streaming 16 B per pixel every step.
            if plan.fused and self.cache_plan and not getattr(plan, "cached", False):
                steps_left = None
                try:
                    steps_left = int(sampling_context.total_steps) - int(sampling_context.step_index)
                except (AttributeError, TypeError, ValueError):
                    pass
                plan._calls = getattr(plan, "_calls", 0) + 1
                if (steps_left is not None and steps_left >= 3) or plan._calls >= 2:
                    plan.build_cache()
            plan.step(x, self.step_finished_inject_ratio, adain=self.adain, cached=getattr(plan, "cached", False))
        else:
            import torch.distributed as dist
            group = None if self.process_group is True else self.process_group
            plan.reduce(x)
            dist.all_reduce(plan.accumulator, op=dist.ReduceOp.SUM, group=group)
            plan.gather(x, self.step_finished_inject_ratio, adain=self.adain)
        if plan.n_valid < 0 and not getattr(plan, "_checked", False):
            # capacity came from a hint instead of a scan: verify once per id batch that no key fell outside the table
            plan._checked = True
            plan.check()
        if x is not noise:
            noise.copy_(x)   # the reference writes frame by frame into the sampler's tensor (corresponder.py:375-376)

    # NB: like the reference class, OverlapCorresponder has no `finished` hook — its node installs a no-op VAE
    # callback (source/comfyUI/stable_rendering/_nodes/samplers.py:112-122).


__all__ = ["Corresponder", "DefaultCorresponder", "OverlapCorresponder"]
