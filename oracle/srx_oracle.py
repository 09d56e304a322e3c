"""TEST INFRASTRUCTURE ONLY — CPU (numpy) restatement of Stable-Renderer's overlap / bake hot path.

This module is the parity checker for the CUDA kernels.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it; the product package never does
(it fails loudly when the CUDA library is missing instead of falling back to this code).

Parity status: **pinned**.  Every function below is checked in ``tests/test_oracle_vs_golden.py`` against
fixtures under ``tests/golden/`` that were produced by running the reference's own, unmodified files through
``oracle/ref_shim.py`` (script: ``oracle/make_golden.py``), plus the two docstring known-answer examples of
``tensor_group_by_then_average`` (reference ``source/common_utils/math_utils.py:110-128``).  When
/root/reference is mounted, ``tests/test_oracle_vs_reference.py`` additionally runs live differential checks.

Conventions (all citations are relative to /root/reference):
  ids      int array [F, H, W, 4] = (spriteID, materialID, map_index, vertexID) per pixel
           (G-buffer format, ``source/engine/shaders/default_Gbuffer.frag.glsl:27-37``)
  x        float array [B, C, h, w] latents
  "entry"  = one valid pixel, in row-major (frame, y, x) order — the row order of
           ``IDMap.create_vertex_screen_info`` (``source/engine/static/corrmap.py:220-280``)

Where the reference leaves behaviour to chance (duplicate-index ``index_put_``: ``corresponder.py:354-359``,
``corrmap.py:735``) this restatement uses the order the reference exhibits with one CPU thread:
the LAST entry in entry order wins.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

NO_ID_MAP_INDEX = 2048  # default_Gbuffer.frag.glsl:8,150 ; corrmap.py:121,267


# =====================================================================================================
# K1-K3: keying
# =====================================================================================================
def idmap_masks(ids: np.ndarray) -> np.ndarray:
    """``IDMap.__attrs_post_init__`` mask rule (corrmap.py:119-124): 1.0 where the pixel has no id
    (map_index == 2048, or all four components zero)."""
    ids = np.asarray(ids)
    no_id = (ids[..., 2] == NO_ID_MAP_INDEX) | np.all(ids == 0, axis=-1)
    return no_id.astype(np.float32)


def valid_entries(ids: np.ndarray) -> np.ndarray:
    """Boolean [F,H,W]: rows that survive the two filters of ``create_vertex_screen_info``
    (corrmap.py:266-275): map_index != 2048, and not (all four ids == 0)."""
    ids = np.asarray(ids)
    return (ids[..., 2] != NO_ID_MAP_INDEX) & np.any(ids != 0, axis=-1)


def vertex_screen_info(ids: np.ndarray, frame_indices: Optional[Sequence[int]] = None) -> np.ndarray:
    """``IDMap.create_vertex_screen_info`` (corrmap.py:220-280): float32 [N,7] =
    (sprite, material, map_index, vertex_id, x/H, y/W, frame_index_value), rows in (frame, y, x) order.

    Faithful details: the ids are routed through float32 (torch.cat type promotion, :256-261); x is divided
    by the *height* and y by the *width* (:239,:249); column 6 carries the frame-index VALUE (:251-253)."""
    ids = np.asarray(ids)
    F, H, W, E = ids.shape
    if frame_indices is None:
        frame_indices = list(range(F))
    keep = valid_entries(ids)
    f_idx, y_idx, x_idx = np.nonzero(keep)  # row-major order == entry order
    out = np.empty((f_idx.size, E + 3), dtype=np.float32)
    out[:, :E] = ids[f_idx, y_idx, x_idx].astype(np.float32)
    out[:, E] = x_idx.astype(np.float32) / np.float32(H)
    out[:, E + 1] = y_idx.astype(np.float32) / np.float32(W)
    out[:, E + 2] = np.asarray(frame_indices, dtype=np.int32)[f_idx].astype(np.float32)
    return out


def entry_cells(vsi: np.ndarray, h: int, w: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """``step_finished`` coordinate step (corresponder.py:312-314): float32 multiply, truncate to int32."""
    sx = (vsi[:, 4] * np.float32(w)).astype(np.int32)
    sy = (vsi[:, 5] * np.float32(h)).astype(np.int32)
    fr = vsi[:, 6].astype(np.int32)
    return sx, sy, fr


def axis_cell_table(n_px: int, divisor: int, n_cells: int) -> np.ndarray:
    """Cell coordinate of every pixel coordinate along one axis: trunc(fl32(fl32(p)/fl32(divisor))*fl32(n_cells)).
    (divisor is H for the x axis and W for the y axis — corrmap.py:239,249.)"""
    p = np.arange(n_px, dtype=np.float32)
    return ((p / np.float32(divisor)) * np.float32(n_cells)).astype(np.int32)


# =====================================================================================================
# S4: group-by-key mean (math_utils.py:86-161)
# =====================================================================================================
def group_by_then_average(t: np.ndarray, index_column: int, value_columns: Sequence[int],
                          accumulate: str = "f32_serial") -> Tuple[np.ndarray, np.ndarray]:
    """``tensor_group_by_then_average``: per unique value of column `index_column` (sorted ascending, as
    ``Tensor.unique``) the float32 mean of `value_columns`, expanded back to every row.

    accumulate = "f32_serial": float32 running sums in row order (the 1-thread ``scatter_add_`` order,
                 math_utils.py:143-150);  "f64": float64 sums rounded once (tolerance referee)."""
    t = np.asarray(t)
    if index_column >= t.shape[-1] or index_column < -t.shape[-1]:
        raise ValueError(f"Index column {index_column} is out of range.")
    if any(c >= t.shape[-1] for c in value_columns):
        raise ValueError(f"Value columns {list(value_columns)} contain out of range values.")
    uniq, inv = np.unique(t[:, index_column], return_inverse=True)
    inv = inv.reshape(-1)
    vals = t[:, list(value_columns)].astype(np.float32)
    cnt = np.bincount(inv, minlength=uniq.size).astype(np.float32)
    if accumulate == "f32_serial":
        sums = np.zeros((uniq.size, vals.shape[1]), dtype=np.float32)
        np.add.at(sums, inv, vals)  # unbuffered, sequential in row order
    elif accumulate == "f64":
        sums = np.stack([np.bincount(inv, weights=vals[:, c].astype(np.float64), minlength=uniq.size)
                         for c in range(vals.shape[1])], axis=1)
    else:
        raise ValueError(accumulate)
    avg = (sums / cnt[:, None].astype(sums.dtype)).astype(np.float32)
    return avg[inv], uniq


# =====================================================================================================
# S6: AdaIN (math_utils.py:27-80)
# =====================================================================================================
def map_mean_std(feat: np.ndarray, eps: float = 1e-5) -> Tuple[np.ndarray, np.ndarray]:
    """``calc_map_mean_std``: per (n, c) mean and sqrt(unbiased var + eps) over h*w.  Statistics are evaluated
    in float64 and rounded once to float32 (torch's CPU var accumulates in double for float input)."""
    n, c = feat.shape[:2]
    flat = feat.reshape(n, c, -1).astype(np.float64)
    var = flat.var(axis=2, ddof=1).astype(np.float32) + np.float32(eps)
    std = np.sqrt(var.astype(np.float32)).astype(np.float32)
    mean = flat.mean(axis=2).astype(np.float32)
    return mean.reshape(n, c, 1, 1), std.reshape(n, c, 1, 1)


def adain(content: np.ndarray, style: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """``adaptive_instance_normalization`` (NCHW): float32 ops in the reference's order
    sub, div, mul, add (math_utils.py:78-80)."""
    assert content.shape[:2] == style.shape[:2]
    s_mean, s_std = map_mean_std(style, eps)
    c_mean, c_std = map_mean_std(content, eps)
    content = content.astype(np.float32)
    normalized = (content - c_mean) / c_std
    return (normalized * s_std + s_mean).astype(np.float32)


# =====================================================================================================
# S1-S7: the current-generation overlap step (corresponder.py:298-376)
# =====================================================================================================
def overlap_step(x: np.ndarray, ids: np.ndarray, frame_indices: Optional[Sequence[int]] = None,
                 ratio: float = 0.1, timestep: float = 1000, stop_timestep: float = 500,
                 accumulate: str = "f32_serial", return_parts: bool = False):
    """``OverlapCorresponder.step_finished``: returns the new latents (the reference mutates
    ``sampling_context.noise`` in place, corresponder.py:375-376).

    x may be float32 / float16 (numpy) — half inputs are up-cast to float32 like corresponder.py:317-318;
    the returned array is float32 (callers cast to the latent dtype).

    Steps: gate (:299-303); cells (:312-314); gather (:324-329); group mean keyed by column 3 only
    (:331-344); blend ``(1-r)*corr + r*avg`` in float32 (:351-352); duplicate-index write-back, last entry in
    entry order wins (:354-359); AdaIN of the ORIGINAL latents to the blended tensor's statistics (:361-364)."""
    x = np.asarray(x)
    if timestep < stop_timestep:
        return (x.astype(np.float32), None) if return_parts else x.astype(np.float32)
    B, C, h, w = x.shape
    vsi = vertex_screen_info(ids, frame_indices)
    sx, sy, fr = entry_cells(vsi, h, w)
    if fr.size and (fr.max() >= B or fr.min() < -B or sx.max() >= w or sy.max() >= h):
        raise IndexError("correspondence entry addresses a latent cell out of range")
    x32 = x.astype(np.float32)
    blended = x32.copy()
    corr = x32[fr, :, sy, sx]                                   # [N, C]
    table = np.concatenate([corr, vsi[:, 3:4]], axis=1)
    avg, uniq = group_by_then_average(table, -1, list(range(C)), accumulate=accumulate)
    r = np.float32(ratio)
    one_minus = np.float32(1 - ratio)
    mixed = one_minus * corr + r * avg                           # three float32 roundings, no fma
    # last writer wins: explicit winner per cell instead of relying on fancy-assignment order
    cell = (fr.astype(np.int64) % B) * (h * w) + sy.astype(np.int64) * w + sx
    winner = np.full(B * h * w, -1, dtype=np.int64)
    np.maximum.at(winner, cell, np.arange(cell.size, dtype=np.int64))
    touched = np.nonzero(winner >= 0)[0]
    bl = blended.transpose(0, 2, 3, 1).reshape(-1, C)            # view [B*h*w, C] of a transposed copy
    bl = np.ascontiguousarray(bl)
    bl[touched] = mixed[winner[touched]]
    blended = bl.reshape(B, h, w, C).transpose(0, 3, 1, 2)
    out = adain(x32, blended)
    if return_parts:
        return out, dict(vsi=vsi, sx=sx, sy=sy, fr=fr, unique_keys=uniq, avg=avg, blended=np.ascontiguousarray(blended),
                         winner=winner)
    return out


# =====================================================================================================
# B1-B4: the bake (corrmap.py:578-736)
# =====================================================================================================
BAKE_MODES = ("replace", "replace_avg", "first", "first_avg")


def corrmap_new(k: int = 3, height: int = 512, width: int = 512, channel_count: int = 4):
    """``CorrespondMap.__attrs_post_init__`` storage (corrmap.py:409-411)."""
    values = np.zeros((k * k, height * width, channel_count), dtype=np.float16)
    writtens = np.zeros((k * k, height * width), dtype=bool)
    return values, writtens


def corrmap_update(values: np.ndarray, writtens: np.ndarray, color_frames: np.ndarray, ids: np.ndarray,
                   spriteID: Optional[int] = None, materialID: Optional[int] = None, mode: str = "first_avg",
                   masks: Optional[np.ndarray] = None, inverse_masks: bool = False,
                   ignore_obj_mat_id: bool = False) -> None:
    """``CorrespondMap.update`` + ``_update`` for 4-D inputs, in place on (values, writtens).

    color_frames [F,H,W,Cin] float; ids [F,H,W,4] int; masks [F,H,W] or None (pixels with mask>0 are kept
    after the optional ``1 - mask`` inversion, :651-654,:704-708).

    Per frame, in order (:661-668): channel fix-up (:681-684); keep mask>0; unless ignore_obj_mat_id keep
    id[0]==spriteID / id[1]==materialID when given (:710-715); modes first/first_avg drop texels that were
    written BEFORE this frame (:720-725); the *_avg / replace de-duplication is a no-op in the reference
    (:727-732); then ``values[map_index, vertexID] = colour.half()`` with duplicates resolved last-pixel-wins
    (:735), ``writtens[...] = True`` (:736).

    Deviation (documented in DESIGN.md): when a mask removed pixels AND a sprite/material filter is active the
    reference re-indexes an already-compacted colour array with original pixel indices (:715) and raises /
    corrupts; here both filters are applied consistently."""
    if mode not in BAKE_MODES:
        raise ValueError(f"unknown update mode {mode}")
    color_frames = np.asarray(color_frames)
    ids = np.asarray(ids)
    if color_frames.ndim == 3:
        color_frames = color_frames[None]
    if ids.ndim == 3:
        ids = ids[None]
    F = color_frames.shape[0]
    if ids.shape[0] != F:
        raise ValueError(f"The length of color_frames and id_maps should be the same, but got: {F} and {ids.shape[0]}")
    if masks is not None:
        masks = np.asarray(masks)
        if masks.ndim == 4 and masks.shape[-1] == 1:
            masks = masks[..., 0]
        if masks.ndim == 2:
            masks = masks[None]
        if masks.shape[0] != F:
            raise ValueError(f"The length of masks should be the same as color_frames, but got: {masks.shape[0]} and {F}")
    n_maps, n_tex, C = values.shape
    for f in range(F):
        col = color_frames[f]
        if C < col.shape[-1]:
            col = col[..., :C]
        elif C == 4 and col.shape[-1] == 3:
            col = np.concatenate([col, np.ones_like(col[..., :1])], axis=-1)
        col = col.reshape(-1, col.shape[-1])
        idf = ids[f].reshape(-1, 4).astype(np.int64)
        keep = np.ones(idf.shape[0], dtype=bool)
        if masks is not None:
            m = masks[f].reshape(-1).astype(np.float32)
            if inverse_masks:
                m = np.float32(1) - m
            keep &= m > 0
        if not ignore_obj_mat_id:
            if spriteID is not None:
                keep &= idf[:, 0] == spriteID
            if materialID is not None:
                keep &= idf[:, 1] == materialID
        pix = np.nonzero(keep)[0]
        mi, vid = idf[pix, 2], idf[pix, 3]
        if pix.size and (mi.max() >= n_maps or mi.min() < -n_maps or vid.max() >= n_tex or vid.min() < -n_tex):
            raise IndexError("map_index / vertexID out of range for this CorrespondMap")
        mi = mi % n_maps
        vid = vid % n_tex
        if mode in ("first", "first_avg"):
            fresh = ~writtens[mi, vid]
            pix, mi, vid = pix[fresh], mi[fresh], vid[fresh]
        tex = mi * n_tex + vid
        # last pixel (row-major) wins among duplicates
        last = np.full(n_maps * n_tex, -1, dtype=np.int64)
        np.maximum.at(last, tex, np.arange(pix.size, dtype=np.int64))
        hit = np.nonzero(last >= 0)[0]
        values.reshape(-1, C)[hit] = col[pix[last[hit]]].astype(np.float16)
        writtens.reshape(-1)[hit] = True


def bake_weight(normal_depth_px: np.ndarray, weight_mode: str) -> np.ndarray:
    """Per-pixel bake weight from the normal+depth attachment (float32 [n,4] = view-space normal*0.5+0.5 in xyz,
    reversed depth in w — default_Gbuffer.frag.glsl:111,123).  All float32, op order as written."""
    nd = np.asarray(normal_depth_px, dtype=np.float32)
    if weight_mode == "uniform":
        return np.ones(nd.shape[0], dtype=np.float32)
    vn = np.abs(np.float32(2) * nd[:, 2] - np.float32(1))
    wgt = np.float32(1) / (np.abs(np.float32(1) - vn) + np.float32(1))
    if weight_mode == "view_normal":
        return wgt
    if weight_mode == "view_normal_depth":
        return wgt * nd[:, 3]
    raise ValueError(weight_mode)


def corrmap_update_weighted(acc: np.ndarray, wsum: np.ndarray, color_frames: np.ndarray, ids: np.ndarray,
                            normal_depth: Optional[np.ndarray], weight_mode: str = "view_normal_depth",
                            masks: Optional[np.ndarray] = None, spriteID: Optional[int] = None,
                            materialID: Optional[int] = None) -> None:
    """Depth/normal-weighted multi-view bake (SURVEY.md §8a row B6 — NOT in the reference: README.md:18-19 lists
    "baking algorithms" / "multi camera baking" as TODO; this is the definition the CUDA kernel is checked against).

    acc float64 [n_maps, n_tex, C] and wsum float64 [n_maps, n_tex] are running sums over all views: a kept pixel
    with weight w adds w*colour and w to its texel (map_index, vertexID).  Weights:
        "uniform"            w = 1
        "view_normal"        w = 1 / (|1 - vn| + 1),  vn = |2*n_z - 1|   (PerpendicularViewNormal's weight,
                             legacy overlap/algorithms.py:111-113, with the view vector (0,0,1))
        "view_normal_depth"  w = depth_reversed / (|1 - vn| + 1)         (closer pixels count more — the
                             'closer wins' test of renderManager.py:121-133 made soft)
    Keep rule: has an id (corrmap.py:119-124), mask>0 ('keep' masks), optional sprite/material filter.
    `corrmap_finalize_weighted` turns the sums into the fp16 atlas."""
    n_maps, n_tex, C = acc.shape
    F = color_frames.shape[0]
    for f in range(F):
        col = color_frames[f]
        if C < col.shape[-1]:
            col = col[..., :C]
        elif C == 4 and col.shape[-1] == 3:
            col = np.concatenate([col, np.ones_like(col[..., :1])], axis=-1)
        col = col.reshape(-1, col.shape[-1]).astype(np.float32)
        idf = ids[f].reshape(-1, 4).astype(np.int64)
        keep = (idf[:, 2] != NO_ID_MAP_INDEX) & np.any(idf != 0, axis=1)
        if masks is not None:
            keep &= masks[f].reshape(-1) > 0
        if spriteID is not None:
            keep &= idf[:, 0] == spriteID
        if materialID is not None:
            keep &= idf[:, 1] == materialID
        pix = np.nonzero(keep)[0]
        mi, vid = idf[pix, 2], idf[pix, 3]
        if pix.size and (mi.max() >= n_maps or mi.min() < 0 or vid.max() >= n_tex or vid.min() < 0):
            raise IndexError("map_index / vertexID out of range for this CorrespondMap")
        if weight_mode == "uniform" or normal_depth is None:
            wgt = np.ones(pix.size, dtype=np.float32)
        else:
            wgt = bake_weight(normal_depth[f].reshape(-1, 4)[pix], weight_mode)
        tex = mi * n_tex + vid
        np.add.at(acc.reshape(-1, C), tex, (wgt[:, None] * col[pix]).astype(np.float64))
        np.add.at(wsum.reshape(-1), tex, wgt.astype(np.float64))


def corrmap_finalize_weighted(values: np.ndarray, writtens: np.ndarray, acc: np.ndarray, wsum: np.ndarray,
                              mode: str = "replace") -> None:
    """Texels with wsum > 0 receive acc/wsum as fp16 and are marked written; mode 'first' keeps texels that were
    already written before this bake (the B3 'first' rule applied to the weighted result)."""
    hit = wsum > 0
    if mode in ("first", "first_avg"):
        hit &= ~writtens
    mean = (acc[hit] / wsum[hit][:, None]).astype(np.float32)
    values[hit] = mean.astype(np.float16)
    writtens[hit] = True


# =====================================================================================================
# L1: schedulers (legacy overlap/overlap_scheduler.py:89-107, overlap/utils.py:24-53)
# =====================================================================================================
def value_interpolation(x: float, start: float, end: float, power: float = 1.0,
                        interpolate_function: str = "constant") -> float:
    assert 0 <= x <= 1
    assert power >= 0
    if interpolate_function == "constant":
        return start
    if interpolate_function == "linear":
        return start + (end - start) * x ** power
    if interpolate_function == "cosine":
        return start + (end - start) * (1 + math.cos(x ** power * math.pi)) / 2
    if interpolate_function == "exponential":
        return start * (end / start) ** (x ** power)
    raise NotImplementedError(interpolate_function)


def scheduler_value(step: int, timestep: float, every_step: int = 1, start_step: int = 0, end_step: int = 1000,
                    start_timestep: int = 0, end_timestep: int = 1000, interpolate_begin: float = 0.0,
                    interpolate_end: float = 1.0, power: float = 1.0, interpolate_type: str = "constant",
                    no_interpolate_return: float = 0.0) -> float:
    """``Scheduler.__call__`` (overlap_scheduler.py:89-107)."""
    if (step < start_step or step > end_step or step % every_step != 0
            or timestep < start_timestep or timestep > end_timestep):
        return no_interpolate_return
    t = 1 - (timestep / 1000)
    return value_interpolation(t, interpolate_begin, interpolate_end, power, interpolate_type)


# =====================================================================================================
# K5: legacy correspondence map (legacy data_classes/correspondence_map.py:145-170, 276-286)
# =====================================================================================================
def correspondence_traces(ids: np.ndarray, merge_len: int = 0) -> Dict[tuple, List[Tuple[int, int, int]]]:
    """``CorrespondenceMap.FromExisting`` inner loops: key = the full id tuple, all-zero ids skipped (:153),
    traces hold (row, col, frame) in (frame, row, col) order, dict order = first appearance.
    merge_len > 0 applies ``merge_nearby`` (:276-286): key -> (obj, mat, texX//d, texY//d)."""
    ids = np.asarray(ids)
    F, H, W, _ = ids.shape
    nz = np.any(ids != 0, axis=-1)
    f_idx, r_idx, c_idx = np.nonzero(nz)
    keys = ids[f_idx, r_idx, c_idx].astype(np.int64)
    out: Dict[tuple, List[Tuple[int, int, int]]] = {}
    for k, f, r, c in zip(map(tuple, keys.tolist()), f_idx.tolist(), r_idx.tolist(), c_idx.tolist()):
        out.setdefault(k, []).append((r, c, f))
    if merge_len:
        merged: Dict[tuple, List[Tuple[int, int, int]]] = {}
        for (o, m, tx, ty), trace in out.items():
            merged.setdefault((o, m, tx // merge_len, ty // merge_len), []).extend(trace)
        out = merged
    return out


# =====================================================================================================
# L2-L4: legacy overlap (overlap/overlap.py:83-222, overlap/algorithms.py:34-118)
# =====================================================================================================
STRATEGIES = ("average", "frame_distance", "pixel_distance", "perpendicular_view_normal")


def strategy_weights(strategy: str, frames: np.ndarray, ys: np.ndarray, xs: np.ndarray,
                     view_normal: Optional[np.ndarray] = None, dtype=np.float64) -> np.ndarray:
    """The dense [L,L] weight matrix each ``OverlapAlgorithm.overlap`` builds (algorithms.py:42-45, 66-70,
    87-93, 109-113)."""
    L = len(frames)
    if strategy == "average":
        return np.ones((L, L), dtype=dtype)
    if strategy == "frame_distance":
        f = np.asarray(frames, dtype=dtype)
        return 1 / (np.abs(f[:, None] - f[None, :]) + 1)
    if strategy == "pixel_distance":
        x = np.asarray(xs, dtype=dtype)
        y = np.asarray(ys, dtype=dtype)
        return 1 / (np.abs(x[:, None] - x[None, :]) + np.abs(y[:, None] - y[None, :]) + 1)
    if strategy == "perpendicular_view_normal":
        vn = np.asarray(view_normal, dtype=dtype).reshape(-1)
        # ones_like(vn).unsqueeze(1) - vn  ->  [L,1] - [L] broadcasts to rows that are all identical
        return 1 / (np.abs(np.ones((L, 1), dtype=dtype) - vn[None, :]) + 1)
    raise ValueError(f"Unknown algorithm {strategy}")


def legacy_overlap(frames: np.ndarray, ids: np.ndarray, alpha: float, strategy: str = "average",
                   merge_len: int = 0, view_normal_map: Optional[np.ndarray] = None,
                   kernel_radius: int = 0, dtype=np.float64) -> np.ndarray:
    """``Overlap.__call__`` (overlap.py:83-152) on a stack [T,B,C,H,W] at correspondence-map resolution.

    For each trace of length >= 2: gather (:136), diagonal pooling with clamping (:61-80,:137-138),
    ``W @ X / W.sum(0)`` (algorithms.py), blend ``alpha*ov + (1-alpha)*latent`` written into the storage that is
    also being read (:97,:145) — i.e. in-place / Gauss–Seidel in dict order, which only matters when
    kernel_radius > 0."""
    X = np.array(frames, dtype=dtype, copy=True)
    T, B, C, H, W = X.shape
    traces = correspondence_traces(ids, merge_len)
    r = int(kernel_radius)
    for trace in traces.values():
        if len(trace) == 1:
            continue
        ys = np.array([t[0] for t in trace])
        xs = np.array([t[1] for t in trace])
        fs = np.array([t[2] for t in trace])
        latent = X[fs, :, :, ys, xs]                           # [L,B,C]
        if r == 0:
            pooled = latent
        else:
            acc = np.zeros_like(latent)
            for d in range(-r, r + 1):
                acc += X[fs, :, :, np.clip(ys + d, 0, H - 1), np.clip(xs + d, 0, W - 1)]
            pooled = acc / (2 * r + 1)
        vn = None
        if strategy == "perpendicular_view_normal":
            vn = np.asarray(view_normal_map)[fs, ys, xs].reshape(-1)
        Wm = strategy_weights(strategy, fs, ys, xs, vn, dtype=dtype)
        flat = pooled.reshape(len(trace), -1)
        ov = (Wm @ flat) / Wm.sum(axis=0).reshape(-1, 1)
        X[fs, :, :, ys, xs] = alpha * ov.reshape(latent.shape) + (1 - alpha) * latent
    return X


def nearest_resize_index(out_size: int, in_size: int) -> np.ndarray:
    """Source index of F.interpolate(mode='nearest') along one axis: min(floor(dst * fl32(in/out)), in-1)."""
    scale = np.float32(in_size) / np.float32(out_size)
    idx = np.floor(np.arange(out_size, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, in_size - 1)


def legacy_resize_overlap(frame_seq: np.ndarray, ids: np.ndarray, alpha: float, strategy: str = "average",
                          merge_len: int = 0, view_normal_map: Optional[np.ndarray] = None,
                          kernel_radius: int = 0, dtype=np.float64) -> np.ndarray:
    """``ResizeOverlap.__call__`` (overlap.py:180-222) for latents [T,B,C,h,w] and ids [T,H,W,4]:
    alpha == 0 returns the input (:200-201); nearest up-sample to the map size (:207); ``Overlap.__call__``;
    nearest down-sample (:216); ``where(ovlp != 0, ovlp, original)`` (:221)."""
    x = np.asarray(frame_seq, dtype=dtype)
    if alpha == 0:
        return x.copy()
    T, B, C, h, w = x.shape
    _, H, W, _ = ids.shape
    uy, ux = nearest_resize_index(H, h), nearest_resize_index(W, w)
    up = x[:, :, :, uy][:, :, :, :, ux]
    ov = legacy_overlap(up, ids, alpha, strategy, merge_len, view_normal_map, kernel_radius, dtype)
    dy, dx = nearest_resize_index(h, H), nearest_resize_index(w, W)
    down = ov[:, :, :, dy][:, :, :, :, dx]
    return np.where(down != 0, down, x)


def build_view_normal_map(normals: np.ndarray, view_vector: np.ndarray) -> np.ndarray:
    """``build_view_normal_map`` (overlap/utils.py:56-102) on float normals [T,H,W,3] in [0,1] (ToTensor scale):
    |n . normalize(view, dim=0)| per pixel -> [T,H,W,1].  ``F.normalize(v, p=2, dim=0)`` acts on the vector as given
    (utils.py:97): a [1,3] vector is divided component-wise by max(|v_c|, 1e-12), a [3] vector by its length."""
    v = np.asarray(view_vector, dtype=np.float32)
    v = v / np.maximum(np.sqrt((v * v).sum(axis=0, keepdims=True)), 1e-12)
    v = v.reshape(-1, 3)[0]
    return np.abs(np.einsum("thwc,c->thw", np.asarray(normals, dtype=np.float32), v))[..., None]


# =====================================================================================================
# L8: same-key broadcast initialiser (math_utils.py:164-229; _nodes/loaders.py:193-250) — grouping only
# =====================================================================================================
def group_slots(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(unique sorted keys, inverse) — the bucketing shared by average / randn_init."""
    uniq, inv = np.unique(np.asarray(keys), return_inverse=True)
    return uniq, inv.reshape(-1)


def randn_init_expand(keys: np.ndarray, random_values: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """``tensor_group_by_then_randn_init`` given the random rows (math_utils.py:216-224): row i receives
    ``random_values[inverse[i]]`` where inverse is the rank of ``keys[i]`` among the sorted unique keys."""
    uniq, inv = group_slots(keys)
    return np.asarray(random_values)[inv], uniq


def noise_sequence_from_ids(ids: np.ndarray, frame_indices: Optional[Sequence[int]], base_latent: np.ndarray,
                            base_noise: np.ndarray, key_latent: np.ndarray, key_noise: np.ndarray, size: int,
                            downsample_option: str = "nearest"):
    """``CreateNoiseSequenceFromIdMap.__call__`` given its four random draws (_nodes/loaders.py:193-271).

    base_* [1,4,size,size] (the CPU draws, :212-219), key_* [n_unique,4] (one row per sorted unique vertex id, the draws
    inside tensor_group_by_then_randn_init).  Steps: repeat the base per frame (:215,220); entries from
    ``create_vertex_screen_info``; pixel = (trunc(x_ratio*size), trunc(y_ratio*size)), frame = column 6 (:224-226); the
    entry's key row overwrites ``tensor[frame, :, y, x]`` — duplicate indices resolved last-entry-wins (:238-243,262-267);
    then 'nearest' (F.interpolate to size/8 keeps pixel (8i, 8j), :269-272) or the [-1,4,8,8] view reduction over dims
    (1,2), which runs over 256 CONSECUTIVE floats of a row and yields 2F frames (:274-285), samples = zeros there."""
    F = ids.shape[0]
    vsi = vertex_screen_info(ids, frame_indices)
    sx = (vsi[:, 4] * np.float32(size)).astype(np.int64)
    sy = (vsi[:, 5] * np.float32(size)).astype(np.int64)
    fr = vsi[:, 6].astype(np.int64)
    uniq, inv = group_slots(vsi[:, 3])
    outs = []
    for base, key in ((base_latent, key_latent), (base_noise, key_noise)):
        full = np.repeat(np.asarray(base, dtype=np.float32), F, axis=0)           # [F,4,size,size]
        flat = fr * (size * size) + sy * size + sx
        last = np.full(F * size * size, -1, dtype=np.int64)
        np.maximum.at(last, flat, np.arange(flat.size, dtype=np.int64))
        hit = np.nonzero(last >= 0)[0]
        vals = np.asarray(key, dtype=np.float32)[inv[last[hit]]]                  # [n_hit,4]
        f_h, rem = hit // (size * size), hit % (size * size)
        full[f_h, :, rem // size, rem % size] = vals
        outs.append(full)
    latent, noise = outs
    h = size // 8
    if downsample_option == "nearest":
        return latent[:, :, ::8, ::8].copy(), noise[:, :, ::8, ::8].copy()
    red = {"mean": lambda a: a.mean(axis=(1, 2), dtype=np.float32), "max": lambda a: a.max(axis=(1, 2)),
           "min": lambda a: a.min(axis=(1, 2))}[downsample_option]
    noise = red(noise.reshape(-1, 4, 8, 8)).reshape(-1, 4, h, h)
    return np.zeros_like(noise), noise


# =====================================================================================================
# On-disk atlas format (corrmap.py:738-872) — the arithmetic of dump / Load
# =====================================================================================================
def corrmap_dump_arrays(values: np.ndarray, writtens: np.ndarray, height: int, width: int):
    """What ``CorrespondMap.dump`` writes into the PNGs (corrmap.py:776-791): per map index
    ``np.clip(255. * map, 0, 255).astype(np.uint8)`` — the product is evaluated in float16 because ``map`` is a float16
    array and 255. a Python scalar — and the written flags as 0 / 255."""
    k2, _, C = values.shape
    v16 = np.asarray(values, dtype=np.float16).reshape(k2, height, width, C)
    img = np.clip(np.float16(255.0) * v16, 0, 255).astype(np.uint8)
    fl = np.clip(255.0 * np.asarray(writtens).reshape(k2, height, width).astype(np.float32), 0, 255).astype(np.uint8)
    return img, fl


def corrmap_load_arrays(img: np.ndarray, flags: np.ndarray):
    """What ``CorrespondMap.Load`` reads back (corrmap.py:846-858): ``float32(u8) / 255.`` stored into the float16
    ``_values``; ``(flag / 255.).bool()``."""
    k2 = img.shape[0]
    C = img.shape[-1] if img.ndim == 4 else 1
    vals = (img.astype(np.float32) / np.float32(255.0)).astype(np.float16).reshape(k2, -1, C)
    wr = (flags.astype(np.float32) / np.float32(255.0)) != 0
    return vals, wr.reshape(k2, -1)


# =====================================================================================================
# K5 maintenance: dropouts on the dict of traces (correspondence_map.py:207-274)
# =====================================================================================================
def traces_dropout_index(traces: Dict[tuple, list], probability: float, seed: int) -> Dict[tuple, list]:
    """``dropout_index`` (:207-223): one ``random.random()`` per key in dict order; keys with a draw < p are deleted."""
    import random
    random.seed(seed)
    out = dict(traces)
    for key in list(out.keys()):
        if random.random() < probability:
            del out[key]
    return out


def traces_dropout_in_rectangle(traces: Dict[tuple, list], rectangle, at_frame: int) -> Dict[tuple, list]:
    """``dropout_in_rectangle`` (:225-274): a key is deleted when one of its tracks at ``at_frame`` lies strictly inside
    ``((r0, c0), (r1, c1))`` (positions are [row, col], :156-160)."""
    (r0, c0), (r1, c1) = rectangle
    out = dict(traces)
    for key, tracks in traces.items():
        for (r, c, f) in tracks:
            if f == at_frame and r0 < r < r1 and c0 < c < c1:
                del out[key]
                break
    return out


# =====================================================================================================
# L8 (legacy): CorrMapLatentNoiseInitializer (legacy_codes/nodes/latent.py:10-40)
# =====================================================================================================
def corrmap_latent_noise_init(traces: Dict[tuple, list], map_width: int, map_height: int, width: int, height: int,
                              batch_size: int, base_latent: np.ndarray, base_noise: np.ndarray,
                              rows: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Arithmetic of the node given its random numbers: ``base_*`` ``[4,Hm,Wm]`` are the two base draws (:22,25, repeated
    over the batch), ``rows`` ``[n_traces,2,4]`` the per-trace (latent, noise) draws in dict order, singleton keys skipped
    (:28-35); then F.interpolate(mode='nearest') to (height // 8, width // 8) (:37-38).  Returns (samples, noise)."""
    latent = np.repeat(np.asarray(base_latent, dtype=np.float32)[None], batch_size, axis=0)
    noise = np.repeat(np.asarray(base_noise, dtype=np.float32)[None], batch_size, axis=0)
    assert latent.shape == (batch_size, 4, map_height, map_width)
    n = 0
    for trace in traces.values():
        if len(trace) == 1:
            continue
        for (r, c, f) in trace:
            latent[f, :, r, c] = rows[n, 0]
            noise[f, :, r, c] = rows[n, 1]
        n += 1
    assert n == len(rows)
    sy = nearest_resize_index(height // 8, map_height)
    sx = nearest_resize_index(width // 8, map_width)
    return latent[:, :, sy][:, :, :, sx], noise[:, :, sy][:, :, :, sx]


def corrmap_latent_noise_draws(seed: int, map_width: int, map_height: int, n_traces: int):
    """The node's random stream, call for call (:18-35): both manual_seed calls return the same default generator, so the
    effective seed is ``seed + 1``; then the two base draws and 2 x ``randn(4)`` per trace."""
    import torch
    torch.manual_seed(seed)
    torch.manual_seed(seed + 1)
    base_latent = torch.randn([1, 4, map_height, map_width])[0].numpy()
    base_noise = torch.randn([1, 4, map_height, map_width])[0].numpy()
    rows = np.zeros((n_traces, 2, 4), dtype=np.float32)
    for i in range(n_traces):
        rows[i, 0] = torch.randn(4).numpy()
        rows[i, 1] = torch.randn(4).numpy()
    return base_latent, base_noise, rows


# =====================================================================================================
# §8f-2: frame ingest (source/engine/managers/renderManager.py:877-948) and closer-pixel merge (:121-133)
# =====================================================================================================
def _h(a) -> np.ndarray:
    return np.asarray(a, dtype=np.float32).astype(np.float16)


def frame_ingest(color: np.ndarray, ids: np.ndarray, pos: np.ndarray, normal_depth: np.ndarray, noise: np.ndarray,
                 canny: np.ndarray, bg_noise: np.ndarray, flip: bool = True) -> Dict[str, np.ndarray]:
    """One frame of ``_save_frame_data``.  Inputs are the attachments ``[H,W,C]`` (fp16 colour / normal+depth / noise,
    int32 ids, f32 position, canny in its own dtype) in GL row order when ``flip``; ``bg_noise`` ``[H,W,4]`` f32.
    Every fp16 operation rounds where torch rounds it: mask (:883), ``1 - mask`` and ``noise * (1 - mask)`` (:929);
    the 64-consecutive-pixel mean of the NHWC view (:933); AdaIN with fp16 style statistics (math_utils.py:39-51)."""
    f = (lambda a: np.asarray(a)[::-1]) if flip else (lambda a: np.asarray(a))
    color, ids, pos, nd, noise, canny = f(color), f(ids), f(pos), f(normal_depth), f(noise), f(canny)
    H, W = color.shape[:2]
    mask = _h(np.float32(1.0) - color[..., 3].astype(np.float32))
    out = {"color_maps": color[..., :3].copy(), "masks": mask, "id_maps": ids.copy(), "pos_maps": pos.copy(),
           "normal_maps": nd[..., :3].copy(), "depth_maps": np.repeat(nd[..., 3:4], 3, axis=-1), "canny_maps": canny.copy()}
    m32 = mask.astype(np.float32)[..., None]
    one_minus = _h(np.float32(1.0) - m32)
    prod = _h(noise.astype(np.float32) * one_minus.astype(np.float32))
    mixed = prod.astype(np.float32) + np.asarray(bg_noise, dtype=np.float32).reshape(H, W, 4) * m32
    pooled = mixed.reshape(-1, 64, 4).astype(np.float64).mean(axis=1)                       # [H*W/64, 4]
    n = pooled.shape[0]
    c_mean = pooled.mean(axis=0).astype(np.float32)
    c_std = np.sqrt(pooled.var(axis=0, ddof=1).astype(np.float32) + np.float32(1e-5))
    s = noise.reshape(-1, 4).astype(np.float64)
    s_var = _h(s.var(axis=0, ddof=1))
    s_std = _h(np.sqrt(_h(s_var.astype(np.float32) + np.float32(1e-5)).astype(np.float32)))
    s_mean = _h(s.mean(axis=0))
    norm = (pooled.astype(np.float32) - c_mean) / c_std
    res = norm * s_std.astype(np.float32) + s_mean.astype(np.float32)
    out["noise_maps"] = np.ascontiguousarray(res.astype(np.float32).T).reshape(4, H // 8, W // 8)
    return out


def gbuffer_merge_closer(temp: Dict[str, np.ndarray], color, ids, pos, normal_depth, noise, canny, flip: bool = True) -> None:
    """One draw of the identical-G-buffer merge, in place on ``temp`` (keys color/ids/pos/normal/depth/noise/canny)."""
    f = (lambda a: np.asarray(a)[::-1]) if flip else (lambda a: np.asarray(a))
    nd = f(normal_depth)
    closer = nd[..., 3].astype(np.float32) > temp["depth"].astype(np.float32)
    temp["depth"][closer] = nd[..., 3][closer]
    temp["normal"][closer] = nd[..., :3][closer]
    temp["color"][closer] = f(color)[closer]
    temp["ids"][closer] = f(ids)[closer]
    temp["pos"][closer] = f(pos)[closer]
    temp["noise"][closer] = f(noise)[closer]
    temp["canny"][closer] = f(canny)[closer].astype(np.float16)


# =====================================================================================================
# §8f-4: wide-channel feature overlap — the body of OverlapCorresponder.post_atten_inject
# (corresponder.py:236-295; unreachable in the reference behind `return origin_values`, :228) — and the
# cell-similarity weighting of taichi_cells_overlap (corr_utils.py:110-134)
# =====================================================================================================
def feature_overlap(features: np.ndarray, ids: np.ndarray, ratio: float = 0.6, map_height: Optional[int] = None,
                    map_width: Optional[int] = None, frame_indices: Optional[Sequence[int]] = None,
                    accumulate: str = "f64", return_parts: bool = False):
    """features [B, h*w, c] (post-attention values, c = 320..1280) -> [B, h*w, c].

    Statement by statement (corresponder.py:240-295): h = w = sqrt(hw); features to [B,c,h,w]; entry coordinates
    ``(vsi[:,4] * id_map.width).to(int)``, ``(vsi[:,5] * id_map.height).to(int)``; nearest up-sampling to
    (id_map.height, id_map.width); gather; group mean keyed by the vertex id; blend with `ratio`; duplicate-index
    write-back (last entry wins); nearest down-sampling to (h, w); AdaIN(content = features, style = down-sampled).

    NB `IDMap.height` / `IDMap.width` are ``tensor.shape[-2]`` / ``tensor.shape[-1]`` (corrmap.py:85-93), i.e. (W, 4)
    for the [F,H,W,4] id tensor — the defaults here — not the image size; `map_height` / `map_width` override them."""
    features = np.asarray(features)
    B, hw, c = features.shape
    h = int(round(math.sqrt(hw)))
    if h * h != hw:
        raise ValueError(f"Dimension hw={hw} is not a perfect square.")
    w = h
    ids = np.asarray(ids)
    mh = int(ids.shape[-2] if map_height is None else map_height)
    mw = int(ids.shape[-1] if map_width is None else map_width)
    feat = features.astype(np.float32).reshape(B, h, w, c)
    vsi = vertex_screen_info(ids, frame_indices)
    sx = (vsi[:, 4] * np.float32(mw)).astype(np.int32)
    sy = (vsi[:, 5] * np.float32(mh)).astype(np.int32)
    fr = vsi[:, 6].astype(np.int32)
    if fr.size and (fr.max() >= B or sx.max() >= mw or sy.max() >= mh):
        raise IndexError("correspondence entry addresses an up-sampled cell out of range")
    up_y, up_x = nearest_resize_index(mh, h), nearest_resize_index(mw, w)     # up-sampled cell -> feature cell
    src = (up_y[sy] * w + up_x[sx]).astype(np.int64)                           # feature cell of every entry
    corr = feat.reshape(B, hw, c)[fr, src]                                     # [N, c]
    uniq, inv = np.unique(vsi[:, 3], return_inverse=True)
    inv = inv.reshape(-1)
    cnt = np.bincount(inv, minlength=uniq.size).astype(np.float32)
    acc_dt = np.float64 if accumulate == "f64" else np.float32
    sums = np.zeros((uniq.size, c), dtype=acc_dt)
    np.add.at(sums, inv, corr.astype(acc_dt))
    avg = (sums / cnt[:, None].astype(acc_dt)).astype(np.float32)
    r, one_minus = np.float32(ratio), np.float32(1 - ratio)
    mixed = one_minus * corr + r * avg[inv]
    # write-back into the up-sampled tensor: last entry per up-sampled cell wins; then nearest down-sampling
    ucell = (fr.astype(np.int64) * mh + sy) * mw + sx
    winner = np.full(B * mh * mw, -1, dtype=np.int64)
    np.maximum.at(winner, ucell, np.arange(ucell.size, dtype=np.int64))
    dn_y, dn_x = nearest_resize_index(h, mh), nearest_resize_index(w, mw)     # feature cell -> up-sampled cell it samples
    Y, X = np.meshgrid(dn_y, dn_x, indexing="ij")
    style = np.empty((B, h, w, c), dtype=np.float32)
    for b in range(B):
        wsel = winner[(b * mh + Y) * mw + X]                                   # [h, w]
        base = feat[b].reshape(hw, c)[(up_y[Y] * w + up_x[X]).reshape(-1)].reshape(h, w, c)   # the up-sampled value there
        style[b] = np.where((wsel >= 0)[..., None], mixed[np.maximum(wsel, 0)], base)
    out = adain(feat.transpose(0, 3, 1, 2), style.transpose(0, 3, 1, 2)).transpose(0, 2, 3, 1).reshape(B, hw, c)
    if return_parts:
        return out, dict(style=style.reshape(B, hw, c), winner=winner, src=src, avg=avg, unique_keys=uniq)
    return out


def cells_overlap(id_flatten_maps: np.ndarray, values: np.ndarray, contributions: np.ndarray) -> np.ndarray:
    """``taichi_cells_overlap`` (corr_utils.py:110-134): every cell becomes the similarity-weighted mean of all cells,
    ``new[A] = (val[A] + sum_{B != A} sim(A,B) val[B]) / (1 + sum_{B != A} sim(A,B))`` with
    ``sim(A,B) = sum_{x in A, i in B} contrib[x] contrib[i] [id[x] == id[i]]`` (all four id components, no validity filter:
    background pixels match each other, :28-44).  A cell is a run of ``pixels // cells`` CONSECUTIVE pixels of the flattened
    frame (:130).  Restated in the factorised form the GPU kernels use: with ``c_A[key]`` = the summed contribution of A's
    pixels carrying `key`, ``sim(A,B) = sum_key c_A[key] c_B[key]``."""
    ids = np.asarray(id_flatten_maps)
    vals = np.asarray(values, dtype=np.float64)
    con = np.asarray(contributions, dtype=np.float64)
    Bn, npx, _ = ids.shape
    cells = vals.shape[1]
    cpp = npx // cells
    used = cells * cpp                                              # trailing pixels belong to no cell
    _, key = np.unique(ids[:, :used].reshape(-1, ids.shape[-1]), axis=0, return_inverse=True)
    key = key.reshape(-1)
    cell = (np.arange(Bn)[:, None] * cells + np.arange(used)[None, :] // cpp).reshape(-1)
    nkeys, ncells = int(key.max()) + 1, Bn * cells
    pair = key.astype(np.int64) * ncells + cell
    up, pinv = np.unique(pair, return_inverse=True)
    cw = np.bincount(pinv.reshape(-1), weights=con[:, :used].reshape(-1), minlength=up.size)    # c_A[key] per (key, cell) pair
    pk, pc = up // ncells, up % ncells
    v = vals.reshape(ncells, -1)
    S = np.zeros((nkeys, v.shape[1]))
    np.add.at(S, pk, cw[:, None] * v[pc])
    T = np.bincount(pk, weights=cw, minlength=nkeys)
    num = v.copy()
    den = np.ones(ncells)
    np.add.at(num, pc, cw[:, None] * (S[pk] - cw[:, None] * v[pc]))
    np.add.at(den, pc, cw * (T[pk] - cw))
    return (num / den[:, None]).reshape(vals.shape)


# =====================================================================================================
# L7: johnny_overlap (legacy_diffuser/modules/diffuser_pipelines/overlap/johnny_overlap.py:15-141)
# =====================================================================================================
def johnny_overlap(frames: np.ndarray, ids: np.ndarray, alpha: float = 1.0, beta: float = 0.0,
                   base: Optional[np.ndarray] = None, merge_len: int = 0) -> np.ndarray:
    """frames [T,B,C,h,w] -> [T,B,C,h,w].  Nearest up-sampling to the map size (:51), per trace (>= 2 entries) every entry i in
    (frame,row,col) order: ``value = sum_j x_j / (|t_i-t_j|+1)``, ``count = sum_j 1/(|t_i-t_j|+1)`` over the CURRENT values
    of the trace (earlier entries already rewritten, :95-118), ``x_i = alpha*value/count + (1-alpha)*x_i``, then
    ``x_i = beta*base_first + (1-beta)*x_i`` with the base colour at the trace's first entry (:112-116); nearest
    down-sampling (:132).  `base` [T,B,C,h,w] = the noised original latents (:63-65)."""
    frames = np.asarray(frames, dtype=np.float64)
    T, B, C, h, w = frames.shape
    H, W = ids.shape[1], ids.shape[2]
    uy, ux = nearest_resize_index(H, h), nearest_resize_index(W, w)
    up = frames[:, :, :, uy][:, :, :, :, ux].copy()
    up_base = None if base is None or beta <= 0 else np.asarray(base, dtype=np.float64)[:, :, :, uy][:, :, :, :, ux]
    for trace in correspondence_traces(ids[:T], merge_len).values():
        L = len(trace)
        if L == 1:
            continue
        fs = np.array([t[2] for t in trace])
        for i, (y, x, f) in enumerate(trace):
            wgt = 1.0 / (np.abs(fs - f) + 1.0)
            cur = np.stack([up[tf, :, :, ty, tx] for (ty, tx, tf) in trace])        # [L,B,C], current values
            ov = alpha * (np.tensordot(wgt, cur, axes=(0, 0)) / wgt.sum()) + (1 - alpha) * up[f, :, :, y, x]
            if up_base is not None:
                y0, x0, f0 = trace[0]
                ov = beta * up_base[f0, :, :, y0, x0] + (1 - beta) * ov
            up[f, :, :, y, x] = ov
    dy, dx = nearest_resize_index(h, H), nearest_resize_index(w, W)
    return up[:, :, :, dy][:, :, :, :, dx]
