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


class _StepContext:
    """SamplingCallbackContext stand-in (reference source/comfyUI/types/runtime.py:543-592)."""

    def __init__(self, noise, denoised, step_index, total_steps, timestep):
        self.noise, self.denoised = noise, denoised
        self.step_index, self.total_steps, self.timestep = step_index, total_steps, timestep


def scripted_ksampler(*args, callbacks=(), **kwargs):
    """A deterministic stand-in for the host's `custom_ksampler`: no model, a fixed arithmetic "denoiser", the callback
    list called after every step with the sampler's own tensors (callbacks mutate `noise` / `denoised` in place) — the
    contract the node callbacks rely on.  Accepts the legacy positional form
    (model, seed, steps, cfg, sampler_name, scheduler, positive, negative, latent, ...) and the keyword form."""
    steps = kwargs.get("steps", args[2] if len(args) > 2 else 8)
    latent = kwargs.get("latent", args[8] if len(args) > 8 else None)
    x = (latent["samples"] if isinstance(latent, dict) else latent).clone()
    for i in range(steps):
        denoised = x * 0.5 + torch.roll(x, 1, 0) * 0.25
        ctx = _StepContext(x, denoised, i, steps, 999.0 - i * (1000 // steps))
        for cb in callbacks:
            cb(ctx)
        x = ctx.noise * 0.8 + ctx.denoised * 0.2
    return ({"samples": x},)
