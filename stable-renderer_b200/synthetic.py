"""Deterministic synthetic correspondence buffers in the renderer's G-buffer format.

The reference's id attachment holds (spriteID, materialID, map_index, vertexID) per pixel
(`source/engine/shaders/default_Gbuffer.frag.glsl:27-37,125-166`): texel id ``Wt*v + u`` of the UV chart under the
pixel, ``map_index`` binned from the view-normal into k*k directions, 2048 for non-AI objects, all-zero where
nothing was drawn.  This generator emits the same format from integer-only arithmetic (bit-identical on CPU and
GPU): per frame one rotating / drifting disc per object carrying an affine UV chart, so that a texel recurs in
nearly every frame, ~40 % pixel coverage like the bundled `resources/example-sphere-and-object-views` dumps.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

NO_ID_MAP_INDEX = 2048


def make_ids(frames: int, height: int, width: int, *, tex_h: int = 512, tex_w: int = 512, k: int = 3,
             n_obj: int = 1, coverage: float = 0.40, frac_2048: float = 0.0, seed: int = 1234,
             device: str | torch.device = "cpu", dtype: torch.dtype = torch.int32,
             frame_offset: int = 0, legacy_layout: bool = False) -> torch.Tensor:
    """ids [frames, H, W, 4] (dtype int32 or int16).

    frame_offset: global index of the first frame (frame-sharded ranks generate their own block of one sequence).
    legacy_layout: emit the legacy dump layout (obj, mat, texX, texY) of
                   `legacy_codes/stable_rendering_algo/data_classes/correspondence_map.py:25-35` instead."""
    dev = torch.device(device)
    H, W = height, width
    ys = torch.arange(H, device=dev, dtype=torch.int64).view(1, H, 1)
    xs = torch.arange(W, device=dev, dtype=torch.int64).view(1, 1, W)
    fr = (torch.arange(frames, device=dev, dtype=torch.int64) + frame_offset).view(frames, 1, 1)

    # objects sit on a g x g grid of discs; total disc area = coverage * H * W
    g = int(math.ceil(math.sqrt(n_obj)))
    radius = int(math.sqrt(coverage * H * W / (math.pi * n_obj)))
    radius = max(2, min(radius, min(H, W) // (2 * g) - 1))
    out = torch.zeros(frames, H, W, 4, device=dev, dtype=torch.int64)
    FX = 1 << 12  # fixed-point one
    # texels per pixel (fixed point): the disc's bounding box maps onto ~90 % of the chart
    tex_scale = int(0.9 * min(tex_h, tex_w) * FX / (2 * radius + 1))
    for o in range(n_obj):
        gx, gy = o % g, o // g
        cx0 = (2 * gx + 1) * W // (2 * g)
        cy0 = (2 * gy + 1) * H // (2 * g)
        # per-frame drift (pixels) and rotation (fixed-point cos/sin) — host scalars, then broadcast
        drift = max(1, radius // 16)
        cxs, cys, cos_t, sin_t = [], [], [], []
        for f in range(frames):
            t = f + frame_offset
            ang = 0.013 * t + 0.5 * o + 0.001 * seed
            cxs.append(cx0 + int(round(drift * math.sin(0.21 * t + o))))
            cys.append(cy0 + int(round(drift * math.cos(0.17 * t + 2 * o))))
            cos_t.append(int(round(math.cos(ang) * FX)))
            sin_t.append(int(round(math.sin(ang) * FX)))
        cx = torch.tensor(cxs, device=dev, dtype=torch.int64).view(frames, 1, 1)
        cy = torch.tensor(cys, device=dev, dtype=torch.int64).view(frames, 1, 1)
        ca = torch.tensor(cos_t, device=dev, dtype=torch.int64).view(frames, 1, 1)
        sa = torch.tensor(sin_t, device=dev, dtype=torch.int64).view(frames, 1, 1)
        dx = xs - cx
        dy = ys - cy
        inside = (dx * dx + dy * dy) <= radius * radius
        # rotate then scale into texel space (all integer)
        ru = (dx * ca - dy * sa) // FX
        rv = (dx * sa + dy * ca) // FX
        u = (ru * tex_scale) // FX + tex_w // 2
        v = (rv * tex_scale) // FX + tex_h // 2
        u = u.clamp_(0, tex_w - 1)
        v = v.clamp_(0, tex_h - 1)
        # view-normal direction bin (default_Gbuffer.frag.glsl:155-163): x_idx + (k-1-y_idx)*k
        kx = ((dx + radius) * k // (2 * radius + 1)).clamp_(0, k - 1)
        ky = ((dy + radius) * k // (2 * radius + 1)).clamp_(0, k - 1)
        map_index = kx + (k - 1 - ky) * k
        if frac_2048 > 0:
            hsh = (xs * 73856093) ^ (ys * 19349663) ^ ((fr + seed) * 83492791)
            hsh = (hsh ^ (hsh >> 13)) * 1274126177
            sel = ((hsh >> 7) & 0xFFFF) < int(frac_2048 * 65536)
            map_index = torch.where(sel, torch.full_like(map_index, NO_ID_MAP_INDEX), map_index)
        if legacy_layout:
            comp = torch.stack([torch.full_like(u, o + 1), torch.zeros_like(u), u, v], dim=-1)
        else:
            comp = torch.stack([torch.full_like(u, o + 1), torch.zeros_like(u), map_index.expand_as(u),
                                v * tex_w + u], dim=-1)
        out = torch.where(inside.unsqueeze(-1), comp, out)
    if dtype == torch.int16:
        assert legacy_layout or tex_h * tex_w <= 32767, "vertex ids do not fit int16"
    return out.to(dtype)


def make_normal_depth(frames: int, height: int, width: int, *, seed: int = 1234,
                      device: str | torch.device = "cpu", frame_offset: int = 0) -> torch.Tensor:
    """normal+depth attachment [frames,H,W,4] float16: view-space unit normal * 0.5 + 0.5 in xyz and reversed
    depth ``1 - gl_FragCoord.z`` in w (`default_Gbuffer.frag.glsl:111,123`)."""
    dev = torch.device(device)
    H, W = height, width
    ys = torch.linspace(-1, 1, H, device=dev).view(1, H, 1)
    xs = torch.linspace(-1, 1, W, device=dev).view(1, 1, W)
    t = (torch.arange(frames, device=dev, dtype=torch.float32) + frame_offset).view(frames, 1, 1)
    nx = 0.8 * xs * torch.cos(0.05 * t) + 0.1 * torch.sin(0.11 * t + seed)
    ny = 0.8 * ys * torch.cos(0.03 * t)
    nz = torch.sqrt(torch.clamp(1 - nx * nx - ny * ny, min=0.01))
    nrm = torch.stack([nx.expand(frames, H, W), ny.expand(frames, H, W), nz], dim=-1)
    nrm = nrm / nrm.norm(dim=-1, keepdim=True)
    depth = (0.25 + 0.7 * nz).clamp(1e-3, 1.0)
    return torch.cat([nrm * 0.5 + 0.5, depth.unsqueeze(-1)], dim=-1).to(torch.float16)


def make_latents(batch: int, channels: int, h: int, w: int, *, seed: int = 0,
                 device: str | torch.device = "cpu", dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """`torch.randn` latents from a CPU generator (identical values on every device), cast to `dtype`."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(batch, channels, h, w, generator=gen, dtype=torch.float32).to(dtype).to(device)


def make_colors(frames: int, height: int, width: int, channels: int = 3, *, seed: int = 7,
                device: str | torch.device = "cpu", dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Decoded frames [frames,H,W,C] in [0,1]."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    return torch.rand(frames, height, width, channels, generator=gen, dtype=torch.float32).to(dtype).to(device)
