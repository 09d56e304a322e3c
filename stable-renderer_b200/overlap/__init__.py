"""Legacy-generation overlap API (reference: legacy_codes/stable_rendering_algo/{overlap,data_classes}):
`Overlap`, `ResizeOverlap`, the four `OverlapAlgorithm` strategies, `Scheduler`, `CorrespondenceMap` and the `CorrMapLatentNoiseInitializer` node (legacy_codes/nodes/latent.py)."""
from .algorithms import (AverageDistance, FrameDistance, OverlapAlgorithm, PerpendicularViewNormal, PixelDistance,
                         overlap_algorithm_factory)
from .correspondence import CorrespondenceMap
from .driver import Overlap, ResizeOverlap
from . import johnny as johnny_overlap
from .latent import CorrMapLatentNoiseInitializer
from .scheduler import Scheduler, value_interpolation
from .view_normal import build_view_normal_map

__all__ = ["OverlapAlgorithm", "AverageDistance", "FrameDistance", "PixelDistance", "PerpendicularViewNormal",
           "overlap_algorithm_factory", "CorrespondenceMap", "Overlap", "ResizeOverlap", "Scheduler",
           "value_interpolation", "build_view_normal_map", "CorrMapLatentNoiseInitializer", "johnny_overlap"]
