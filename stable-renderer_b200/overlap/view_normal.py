"""`build_view_normal_map` (reference: legacy_codes/stable_rendering_algo/overlap/utils.py:56-102)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def build_view_normal_map(normal_images, view_vector: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """|n . v/||v||| per pixel -> [T,H,W,1].  `normal_images` is a list of PIL images (as in the reference) or a float
    tensor [T,H,W,3] already scaled to [0,1]."""
    if isinstance(normal_images, torch.Tensor):
        normal_map = normal_images.to(dtype)
    else:
        if not isinstance(normal_images, list):
            raise TypeError("normal_images must be a list of PIL Image objects.")
        import numpy as np
        normal_map = torch.stack([torch.from_numpy(np.asarray(im, dtype=np.float32) / 255.0)[..., :3] for im in normal_images])
    # F.normalize(..., dim=0) on the vector AS GIVEN (utils.py:97): a [1,3] view vector is therefore normalised per
    # component (-> its sign pattern), a [3] vector to unit length — kept as the reference does it
    v = F.normalize(view_vector.to(dtype), p=2, dim=0).to(normal_map.device)
    v = v.reshape(-1, normal_map.shape[-1])[0] if v.dim() > 1 else v
    return (normal_map * v).sum(dim=-1, keepdim=True).abs()
