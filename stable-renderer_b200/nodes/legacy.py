"""Legacy-generation nodes: `StableRenderSampler`, `OverlapScheduler` (and `CorrMapLatentNoiseInitializer`, which lives in
overlap/latent.py).  Reference: legacy_codes/nodes/samplers.py:16-144, legacy_codes/nodes/schedulers.py:7-39.

`StableRenderSampler` applies `ResizeOverlap` — here the CUDA legacy overlap — to the sampler's noise and / or denoised
latents after every step, in place, through the sampler's callback list."""

import warnings
from typing import Callable, Literal

from ..overlap import CorrespondenceMap, ResizeOverlap, Scheduler, overlap_algorithm_factory
from ._base import (COMFY_SAMPLERS, COMFY_SCHEDULERS, FLOAT, INT, LATENT, MODEL, SamplingCallbackContext,
                    StableRenderingNode, get_ksampler)

_default_sampler = COMFY_SAMPLERS.__args__[0]   # type: ignore
_default_scheduler = COMFY_SCHEDULERS.__args__[0]   # type: ignore

OverlapAlgorithm = Literal["average", "frame_distance", "pixel_distance", "perpendicular_view_normal"]


def estimated_denoising_timestep(step_index: int, total_steps: int) -> int:
    """samplers.py:83,99 — the node has no access to the sampler's sigma schedule and estimates the timestep linearly."""
    return 1000 - int(((step_index + 1) / total_steps) * 1000)


def make_overlap_callback(overlap: Callable, correspondence_map, apply_overlap_option: str, sampler_name: str) -> Callable:
    """The `execute_overlap` callback of samplers.py:79-129: overlap on `context.noise`, on `context.denoised` (ddpm only:
    other samplers fall back to the noise with a warning) or on both, written back frame by frame, in place."""

    def _apply(context: SamplingCallbackContext, target):
        frame_seq = [frame.unsqueeze(0) for frame in target]
        overlapped_frame_seq = overlap(frame_seq,
                                       corr_map=correspondence_map,
                                       step=context.step_index,
                                       timestep=estimated_denoising_timestep(context.step_index, context.total_steps))
        for i, frame in enumerate(overlapped_frame_seq):
            target[i] = frame.squeeze()

    def execute_overlap_on_noise(context: SamplingCallbackContext):
        _apply(context, context.noise)

    def execute_overlap_on_denoised(context: SamplingCallbackContext):
        _apply(context, context.denoised)

    def execute_overlap(context: SamplingCallbackContext):
        if apply_overlap_option == 'noise':
            execute_overlap_on_noise(context)
        elif apply_overlap_option == 'denoised':
            if sampler_name != "ddpm":
                warnings.warn("apply_overlap_option is set to 'denoised' but sampler_name is not 'ddpm'. "
                              "Using on noise version instead.")
                execute_overlap_on_noise(context)
            else:
                execute_overlap_on_denoised(context)
        elif apply_overlap_option == 'both':
            if sampler_name != "ddpm":
                warnings.warn("apply_overlap_option is set to 'both' but sampler_name is not 'ddpm'. "
                              "Using on noise version instead.")
                execute_overlap_on_noise(context)
            else:
                execute_overlap_on_noise(context)
                execute_overlap_on_denoised(context)
        else:
            raise ValueError(f"Unknown apply_overlap_option: {apply_overlap_option}")

    return execute_overlap


class StableRenderSampler(StableRenderingNode):

    Category = "sampling"

    def __call__(self,
                 model: MODEL,
                 positive: "CONDITIONING",  # type: ignore  # noqa: F821
                 negative: "CONDITIONING",  # type: ignore  # noqa: F821
                 latent_image: LATENT,
                 correspondence_map: CorrespondenceMap,
                 alpha_scheduler: Scheduler,
                 kernel_radius_scheduler: Scheduler,
                 overlap_algorithm: OverlapAlgorithm = "average",
                 noise_option: Literal['disable', 'default', 'incoming'] = 'default',
                 apply_overlap_option: Literal['noise', 'denoised', 'both'] = 'noise',
                 noise_seed: INT(0, 0xffffffffffffffff) = 0,  # type: ignore
                 steps: INT(1, 10000) = 20,  # type: ignore
                 cfg: FLOAT(0.0, 100.0, 0.01, round=0.01) = 8.0,  # type: ignore
                 sampler_name: COMFY_SAMPLERS = _default_sampler,
                 scheduler: COMFY_SCHEDULERS = _default_scheduler,
                 denoise: FLOAT(0, 1) = 1.0  # type: ignore
                 ) -> LATENT:
        """Sampling with a correspondence map: see the reference docstring (samplers.py:38-62) for the arguments."""
        SUPPORTED_SAMPLERS = ["ddim", "ddpm"]
        if sampler_name not in SUPPORTED_SAMPLERS:
            warnings.warn(f"Scheduling with {sampler_name} is not supported. Using default sampler instead.")
            sampler_name = "ddpm"

        overlap = ResizeOverlap(
            alpha_scheduler=alpha_scheduler,
            kernel_radius_scheduler=kernel_radius_scheduler,
            algorithm=overlap_algorithm_factory(overlap_algorithm),
        )
        callbacks = [make_overlap_callback(overlap, correspondence_map, apply_overlap_option, sampler_name)]

        return get_ksampler()(model,
                              noise_seed,
                              steps,
                              cfg,
                              sampler_name,
                              scheduler,
                              positive,
                              negative,
                              latent_image,
                              denoise=denoise,
                              noise_option=noise_option,
                              callbacks=callbacks)


class OverlapScheduler(StableRenderingNode):

    Category = "scheduler"

    def __call__(self,
                 every_step: INT(1, 1000) = 1,  # type: ignore

                 start_step: INT(1, 1000) = 1,  # type: ignore
                 end_step: INT(1, 1000) = 1000,  # type: ignore

                 start_timestep: INT(0, 1000) = 0,  # type: ignore
                 end_timestep: INT(0, 1000) = 1000,  # type: ignore

                 interpolate_begin: FLOAT(0, 1) = 0.0,  # type: ignore
                 interpolate_end: FLOAT(0, 1) = 1.0,  # type: ignore
                 power: float = 1.0,
                 interpolate_type: Literal["constant", "linear", "cosine", "exponential"] = 'constant',

                 no_interpolate_return: FLOAT(0, 1) = 0.0,) -> Scheduler:  # type: ignore
        return Scheduler(
            every_step=every_step,
            start_step=start_step,
            end_step=end_step,
            start_timestep=start_timestep,
            end_timestep=end_timestep,
            interpolate_begin=interpolate_begin,
            interpolate_end=interpolate_end,
            power=power,
            interpolate_type=interpolate_type,
            no_interpolate_return=no_interpolate_return,
        )


__all__ = ["StableRenderSampler", "OverlapScheduler", "make_overlap_callback", "estimated_denoising_timestep"]
