"""Duck-typed stand-ins for the host application's EngineData / SamplingCallbackContext
(reference: source/comfyUI/types/hidden.py:249-331, types/runtime.py:543-592)."""
import numpy as np
import torch


class Ctx:
    def __init__(self, noise, timestep=900.0, step_index=0, total_steps=20, denoised=None):
        self.noise = noise
        self.denoised = noise if denoised is None else denoised
        self.timestep = timestep
        self.step_index = step_index
        self.total_steps = total_steps


class EngineData:
    def __init__(self, id_maps=None, correspond_maps=None, noise_maps=None, normal_maps=None):
        self.id_maps = id_maps
        self.correspond_maps = correspond_maps
        self.noise_maps = noise_maps
        self.normal_maps = normal_maps


def t2n(t: torch.Tensor) -> np.ndarray:
    return t.detach().float().cpu().numpy()


def assert_close(got, want, rtol, atol, what=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    err = np.abs(got - want)
    tol = atol + rtol * np.abs(want)
    bad = err > tol
    assert not bad.any(), (f"{what}: {bad.sum()} of {bad.size} elements differ; max abs err {err.max():.3e}, "
                           f"max rel err {(err / np.maximum(np.abs(want), 1e-30)).max():.3e}")
