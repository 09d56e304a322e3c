"""The four weighting strategies (reference: legacy_codes/stable_rendering_algo/overlap/algorithms.py:6-133).

Inside `Overlap` / `ResizeOverlap` a strategy is only a tag: the CUDA kernel evaluates each row of the weight matrix
on the fly.  The per-trace `overlap()` method of the protocol is kept for callers that hold their own traces; it
builds the same dense [L, L] weights with device tensor ops."""
from __future__ import annotations

from typing import Literal, Protocol

import torch


class OverlapAlgorithm(Protocol):
    strategy: str

    def overlap(self, latent_seq: torch.Tensor, frame_index_trace: list, x_position_trace: list,
                y_position_trace: list, **kwargs) -> torch.Tensor: ...


def _mix(weights: torch.Tensor, latent_seq: torch.Tensor) -> torch.Tensor:
    flat = latent_seq.reshape(latent_seq.shape[0], -1)
    return (weights @ flat / weights.sum(dim=0).reshape(-1, 1)).reshape_as(latent_seq)


class AverageDistance:
    strategy = "average"

    def overlap(self, latent_seq, frame_index_trace, x_position_trace, y_position_trace, **kwargs):
        n = len(frame_index_trace)
        return _mix(torch.ones(n, n, dtype=latent_seq.dtype, device=latent_seq.device), latent_seq)


class FrameDistance:
    strategy = "frame_distance"

    def overlap(self, latent_seq, frame_index_trace, x_position_trace, y_position_trace, **kwargs):
        f = torch.tensor(frame_index_trace, dtype=latent_seq.dtype, device=latent_seq.device)
        return _mix(1 / ((f[:, None] - f[None, :]).abs() + 1), latent_seq)


class PixelDistance:
    strategy = "pixel_distance"

    def overlap(self, latent_seq, frame_index_trace, x_position_trace, y_position_trace, **kwargs):
        x = torch.tensor(x_position_trace, dtype=latent_seq.dtype, device=latent_seq.device)
        y = torch.tensor(y_position_trace, dtype=latent_seq.dtype, device=latent_seq.device)
        return _mix(1 / ((x[:, None] - x[None, :]).abs() + (y[:, None] - y[None, :]).abs() + 1), latent_seq)


class PerpendicularViewNormal:
    strategy = "perpendicular_view_normal"

    def overlap(self, latent_seq, frame_index_trace, x_position_trace, y_position_trace, view_normal_map=None, **kwargs):
        vn = view_normal_map[list(frame_index_trace), list(y_position_trace), list(x_position_trace)]
        vn = vn.reshape(-1).to(device=latent_seq.device, dtype=latent_seq.dtype)
        weights = (1 / ((1 - vn).abs() + 1))[None, :].expand(vn.numel(), -1)   # every row identical (algorithms.py:111-113)
        return _mix(weights, latent_seq)


def overlap_algorithm_factory(
        algorithm: Literal["average", "frame_distance", "pixel_distance", "perpendicular_view_normal"]) -> OverlapAlgorithm:
    table = {"average": AverageDistance, "frame_distance": FrameDistance, "pixel_distance": PixelDistance,
             "perpendicular_view_normal": PerpendicularViewNormal}
    if algorithm not in table:
        raise ValueError(f"Unknown algorithm {algorithm}")
    return table[algorithm]()
