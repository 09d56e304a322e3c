"""Current-generation nodes: `DefaultCorresponder`, `OverlapCorresponder`, `CorrespondSampler`
(reference: source/comfyUI/stable_rendering/_nodes/samplers.py:20-201).

Same class names, parameter names, defaults and return annotations as the reference; what they build are this package's
GPU corresponders, so the sampler loop calls the CUDA overlap step / bake instead of the torch op chain."""

from functools import partial
from typing import Optional

from ..corresponder import Corresponder
from ..corresponder import DefaultCorresponder as _DefaultCorresponder
from ..corresponder import OverlapCorresponder as _OverlapCorresponder
from ..corrmap import UpdateMode
from ._base import (COMFY_SAMPLERS, COMFY_SCHEDULERS, FLOAT, INT, LATENT, MODEL, EngineData, SamplingCallbackContext,
                    StableRenderingNode, VAEDecodeCallback, dev_looping, get_ksampler, is_empty_method)

_default_sampler = COMFY_SAMPLERS.__args__[0]   # type: ignore
_default_scheduler = COMFY_SCHEDULERS.__args__[0]   # type: ignore


def _vae_callback(corresponder, engine_data):
    """`finished` hook as the VAE-decode callback (samplers.py:57-67); a corresponder without one gets a no-op."""
    if hasattr(corresponder, "finished") and not is_empty_method(corresponder.finished):
        return partial(corresponder.finished, engine_data)
    return lambda *args, **kwargs: None


class DefaultCorresponder(StableRenderingNode):

    Category = "sampling"

    def __call__(self,
                 engine_data: EngineData,   # hidden value
                 update_corrmap: bool = True,
                 update_mode: UpdateMode = 'first_avg',
                 post_attn_inject_ratio: float = 0.6,
                 ) -> tuple[Corresponder, VAEDecodeCallback]:
        """The equal-contribution corresponder: bakes the decoded frames into the atlases at `finished`."""
        corresponder = _DefaultCorresponder(update_corrmap=update_corrmap,
                                            update_corrmap_mode=update_mode,
                                            post_attn_inject_ratio=post_attn_inject_ratio)
        return corresponder, _vae_callback(corresponder, engine_data)   # type: ignore


class OverlapCorresponder(StableRenderingNode):

    Category = "sampling"

    def __call__(self,
                 engine_data: EngineData,   # hidden value
                 update_corrmap: bool = True,
                 update_mode: UpdateMode = 'first_avg',
                 pre_attn_inject_num_of_random_frames: int = 1,
                 post_attn_inject_ratio: float = 0.6,
                 step_finished_inject_ratio: FLOAT(min=0, max=1, step=0.1, round=0.01) = 0.5,  # type: ignore
                 step_finished_stop_inject_timestep: INT(1, 1000, step=100) = 500,  # type: ignore
                 ) -> tuple[Corresponder, VAEDecodeCallback]:
        """The overlap corresponder: same-key latents are averaged and blended after every denoise step."""
        corresponder = _OverlapCorresponder(update_corrmap=update_corrmap,
                                            update_corrmap_mode=update_mode,
                                            pre_attn_inject_num_random_frames=pre_attn_inject_num_of_random_frames,
                                            post_attn_inject_ratio=post_attn_inject_ratio,
                                            step_finished_inject_ratio=step_finished_inject_ratio,
                                            step_finished_stop_inject_timestep=step_finished_stop_inject_timestep)
        return corresponder, _vae_callback(corresponder, engine_data)   # type: ignore


def step_callbacks(corresponder, engine_data) -> list:
    """The per-step callback list `CorrespondSampler` hands to the sampler (samplers.py:166-176)."""
    callbacks = []
    if hasattr(corresponder, "step_finished") and not is_empty_method(corresponder.step_finished):
        def on_1_step_finished(engine_data, context: SamplingCallbackContext):
            corresponder.step_finished(engine_data, context)
        callbacks.append(partial(on_1_step_finished, engine_data))
    return callbacks


class CorrespondSampler(StableRenderingNode):

    Category = "sampling"

    def __call__(self,
                 model: MODEL,
                 positive: "CONDITIONING",  # noqa: F821
                 negative: "CONDITIONING",  # noqa: F821
                 corresponder: Corresponder,
                 engine_data: EngineData,   # hidden value
                 latent: Optional[LATENT] = None,  # if none, data comes from `engine_data.noise_maps`
                 steps: INT(1, 10000) = 20,  # type: ignore
                 cfg: FLOAT(0.0, 100.0, 0.01, round=0.01) = 8.0,  # type: ignore
                 sampler_name: COMFY_SAMPLERS = _default_sampler,
                 scheduler: COMFY_SCHEDULERS = _default_scheduler,
                 denoise: FLOAT(0, 1) = 1.0  # type: ignore
                 ) -> LATENT:
        """Sampler of the baking process: the corresponder's hooks ride on the sampler's callbacks."""
        if isinstance(corresponder, _OverlapCorresponder) and sampler_name not in ['ddim', 'ddpm']:
            raise ValueError("OverlapCorresponder only works with ddim or ddpm sampler_name.")

        if hasattr(corresponder, 'prepare') and not is_empty_method(corresponder.prepare):
            corresponder.prepare(engine_data)

        callback = step_callbacks(corresponder, engine_data)
        if dev_looping():
            def print_progress(context: SamplingCallbackContext):
                print(f"Step {context.step_index + 1}/{context.total_steps} finished.")
            callback.append(print_progress)  # type: ignore

        if latent is None:
            if engine_data is not None:
                latent = engine_data.noise_maps
            else:
                raise ValueError("Input latent is None and engine_data is also None.")

        return get_ksampler()(model=model,
                              seed=None,
                              steps=steps,
                              cfg=cfg,
                              sampler_name=sampler_name,
                              scheduler=scheduler,
                              positive=positive,
                              negative=negative,
                              latent=latent,   # type: ignore
                              denoise=denoise,
                              noise_option='incoming',
                              engine_data=engine_data,     # kwargs for the attention layers
                              corresponder=corresponder,   # kwargs for the attention layers
                              callbacks=callback)[0]   # type: ignore


__all__ = ['DefaultCorresponder', 'OverlapCorresponder', 'CorrespondSampler', 'step_callbacks']
