"""TEST / BASELINE INFRASTRUCTURE ONLY — multi-threaded CPU port of the reference's overlap step built from the same
class of torch CPU ops the reference spends its time in (`unique` sort, `scatter_add_`, advanced-index gather,
`index_put_`, `var`/`mean`; reference source/common_utils/stable_render_utils/corresponder.py:298-376 and
source/common_utils/math_utils.py:27-161).  It exists so that `bench.py` can time "the reference's CPU torch path"
on the GPU box, where /root/reference is not mounted (`cpu_baseline.kind = "port"`, and the `--impl reference` arm).

Only bench.py's CPU legs and tests/ may import this module.  Checked against the reference-generated golden
fixtures in tests/test_oracle_vs_golden.py.  Unlike oracle/srx_oracle.py (explicit winner per cell) the duplicate
write-back here is torch's `index_put_`, as in the reference — deterministic only with one thread."""
from __future__ import annotations

from typing import Optional, Sequence

import torch


class CpuOverlapPort:
    """Plan (cached like IDMap._vertex_screen_info_cache, corrmap.py:226,278) + per-step arithmetic."""

    def __init__(self, ids: torch.Tensor, frame_indices: Optional[Sequence[int]] = None):
        F, H, W, _ = ids.shape
        if frame_indices is None:
            frame_indices = list(range(F))
        keep = (ids[..., 2] != 2048) & (ids != 0).any(dim=-1)                 # corrmap.py:266-275
        f_idx, y_idx, x_idx = keep.nonzero(as_tuple=True)                     # row-major == entry order
        self.key = ids[f_idx, y_idx, x_idx, 3].to(torch.float32)              # key through float32 (corrmap.py:256-261)
        self.x_ratio = x_idx.to(torch.float32) / H                            # corrmap.py:239 (sic)
        self.y_ratio = y_idx.to(torch.float32) / W                            # corrmap.py:249 (sic)
        self.frame = torch.tensor(list(frame_indices), dtype=torch.int32)[f_idx]
        self.n_entries = int(self.key.numel())

    def step(self, x: torch.Tensor, ratio: float) -> torch.Tensor:
        """Returns the new latents for x [B,C,h,w] (float32)."""
        B, C, h, w = x.shape
        sx = (self.x_ratio * w).to(torch.int64)                               # corresponder.py:312-314
        sy = (self.y_ratio * h).to(torch.int64)
        fr = self.frame.to(torch.int64)
        work = x.clone().to(torch.float32)
        corr = work[fr, :, sy, sx]                                             # gather [N,C]
        uniq, inv = self.key.unique(return_inverse=True)                      # the sort the reference pays every step
        idx = inv.unsqueeze(1).expand(-1, C)
        sums = torch.zeros(uniq.numel(), C, dtype=torch.float32).scatter_add_(0, idx, corr)
        cnts = torch.zeros(uniq.numel(), C, dtype=torch.float32).scatter_add_(0, idx, torch.ones_like(corr))
        avg = (sums / cnts)[inv]
        mixed = (1 - ratio) * corr + ratio * avg
        work[fr, :, sy, sx] = mixed                                            # duplicate-index write-back
        flat_c, flat_s = x.reshape(B, C, -1).float(), work.reshape(B, C, -1)
        c_mean, c_std = flat_c.mean(2, keepdim=True), (flat_c.var(2, keepdim=True) + 1e-5).sqrt()
        s_mean, s_std = flat_s.mean(2, keepdim=True), (flat_s.var(2, keepdim=True) + 1e-5).sqrt()
        out = (flat_c - c_mean) / c_std * s_std + s_mean
        return out.reshape(B, C, h, w)


def cpu_bake_port(values: torch.Tensor, writtens: torch.Tensor, colors: torch.Tensor, ids: torch.Tensor,
                  masks: Optional[torch.Tensor], mode: str = "replace") -> None:
    """CorrespondMap.update with ignore_obj_mat_id=True, inverse_masks=True (corrmap.py:661-736) — the path
    DefaultCorresponder.finished takes in the reference's own `__main__` check (corresponder.py:421-422)."""
    C = values.shape[-1]
    for f in range(colors.shape[0]):
        col = colors[f]
        if C < col.shape[-1]:
            col = col[..., :C]
        elif C == 4 and col.shape[-1] == 3:
            col = torch.cat([col, torch.ones_like(col[..., :1])], dim=-1)
        col = col.reshape(-1, col.shape[-1])
        idf = ids[f].reshape(-1, 4).to(torch.int64)
        keep = torch.ones(idf.shape[0], dtype=torch.bool)
        if masks is not None:
            keep &= (1 - masks[f].reshape(-1)) > 0
        mi, vid, cc = idf[keep, 2], idf[keep, 3], col[keep]
        if mode in ("first", "first_avg"):
            fresh = ~writtens[mi, vid]
            mi, vid, cc = mi[fresh], vid[fresh], cc[fresh]
        values[mi, vid] = cc.to(values.dtype)
        writtens[mi, vid] = True
