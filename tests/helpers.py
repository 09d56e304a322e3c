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


def torch_chain(ids: torch.Tensor, x: torch.Tensor, ratio: float) -> torch.Tensor:
    """The reference's op chain (corresponder.py:298-376) in torch ops with float64 sums and an explicit last writer per cell: the
    referee of the full-size GPU tests (any device; pinned to the reference fixtures on the CPU)."""
    F, H, W, _ = ids.shape
    B, C, h, w = x.shape
    x64 = x.double()
    K = int(ids[..., 3].max().item()) + 1
    sums = torch.zeros(K, C, dtype=torch.float64, device=x.device)
    cnts = torch.zeros(K, dtype=torch.float64, device=x.device)
    winner = torch.full((B * h * w,), -1, dtype=torch.int64, device=x.device)     # packed (pixel order, key) of the last valid pixel
    for f in range(F):
        idf = ids[f]
        keep = (idf[..., 2] != 2048) & (idf != 0).any(dim=-1)                     # corrmap.py:266-275
        yy, xx = keep.nonzero(as_tuple=True)
        key = idf[yy, xx, 3].to(torch.float32).long()                              # the key goes through float32 (corrmap.py:256-261)
        sx = ((xx.float() / H) * w).long()                                         # corresponder.py:312-313 with corrmap.py:239,249
        sy = ((yy.float() / W) * h).long()
        cell = (f * h + sy) * w + sx
        vals = x64[f][:, sy, sx].t()                                               # [n, C]
        sums.index_add_(0, key, vals)
        cnts.index_add_(0, key, torch.ones_like(key, dtype=torch.float64))
        order = yy * W + xx                                                        # entry order inside the frame
        winner.scatter_reduce_(0, cell, order * K + key, reduce="amax")
    mean = (sums / cnts.clamp(min=1).unsqueeze(1)).float()                         # reference sums and divides in float32
    has = winner >= 0
    wkey = (winner % K).clamp(min=0)
    flat = x.float().permute(0, 2, 3, 1).reshape(-1, C)                            # [B*h*w, C]
    blended = torch.where(has.unsqueeze(1), (1 - ratio) * flat + ratio * mean[wkey], flat)
    b = blended.reshape(B, h, w, C).permute(0, 3, 1, 2).reshape(B, C, -1).double()
    c = x64.reshape(B, C, -1)
    c_mean, c_std = c.mean(2, keepdim=True), (c.var(2, keepdim=True) + 1e-5).sqrt()
    s_mean, s_std = b.mean(2, keepdim=True), (b.var(2, keepdim=True) + 1e-5).sqrt()
    return ((c - c_mean) / c_std * s_std + s_mean).reshape(B, C, h, w)


