"""`Scheduler` / `value_interpolation` — host-side scalars that drive alpha and the kernel radius
(reference: legacy_codes/stable_rendering_algo/overlap/overlap_scheduler.py:8-107, overlap/utils.py:24-53)."""
from __future__ import annotations

import math
from typing import Literal

__ALL_SCHEDULER_TYPE__ = ["constant", "linear", "cosine", "exponential"]


def value_interpolation(x, start: float, end: float, power: float = 1.0,
                        interpolate_function: Literal["constant", "linear", "cosine", "exponential"] = "constant"):
    """Interpolates from `start` to `end` as x goes 0 -> 1.  Accepts Python numbers and 0-d tensors (the reference's
    cosine branch only works for tensors, utils.py:49; here both do)."""
    assert x >= 0 and x <= 1
    assert power >= 0
    if interpolate_function == "constant":
        return start
    if interpolate_function == "linear":
        return start + (end - start) * x ** power
    if interpolate_function == "cosine":
        xp = x ** power
        c = xp.mul(math.pi).cos() if hasattr(xp, "cos") else math.cos(xp * math.pi)
        return start + (end - start) * (1 + c) / 2
    if interpolate_function == "exponential":
        return start * (end / start) ** (x ** power)
    raise NotImplementedError(interpolate_function)


class Scheduler:
    """Same constructor and call contract as the reference `Scheduler` (overlap_scheduler.py:18-107)."""

    def __init__(self, every_step: int = 1, start_step: int = 0, end_step: int = 1000, start_timestep: int = 0,
                 end_timestep: int = 1000, interpolate_begin: float = 0.0, interpolate_end: float = 1.0,
                 power: float = 1.0,
                 interpolate_type: Literal["constant", "linear", "cosine", "exponential", ""] = "constant",
                 no_interpolate_return: float = 0.0):
        self._every_step = every_step
        self._start_step = start_step
        self._end_step = end_step
        self._start_timestep = start_timestep
        self._end_timestep = end_timestep
        self._interpolate_start = interpolate_begin
        self._interpolate_end = interpolate_end
        self._power = power
        self._interpolate_type = interpolate_type
        self._no_interpolate_return = no_interpolate_return

    def __call__(self, step: int = None, timestep: int = None, **kwargs):
        inactive = (step < self._start_step or step > self._end_step or step % self._every_step != 0
                    or timestep < self._start_timestep or timestep > self._end_timestep)
        if inactive:
            return self._no_interpolate_return
        t = 1 - (timestep / 1000)
        return value_interpolation(t, start=self._interpolate_start, end=self._interpolate_end, power=self._power,
                                   interpolate_function=self._interpolate_type)
