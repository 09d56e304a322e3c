"""ComfyUI node surface of the hot path: the reference's node classes with their `__call__`
schemas, building this package's GPU corresponders / overlap objects."""
from ..loaders import CreateNoiseSequenceFromIdMap
from ..overlap.latent import CorrMapLatentNoiseInitializer
from ._base import IN_COMFY, StableRenderingNode, get_ksampler, is_empty_method, set_ksampler
from .legacy import OverlapScheduler, StableRenderSampler, estimated_denoising_timestep, make_overlap_callback
from .samplers import CorrespondSampler, DefaultCorresponder, OverlapCorresponder, step_callbacks

__all__ = ["CorrespondSampler", "DefaultCorresponder", "OverlapCorresponder", "StableRenderSampler", "OverlapScheduler",
           "CorrMapLatentNoiseInitializer", "CreateNoiseSequenceFromIdMap", "StableRenderingNode", "set_ksampler",
           "get_ksampler", "is_empty_method", "make_overlap_callback", "estimated_denoising_timestep", "step_callbacks",
           "IN_COMFY"]
