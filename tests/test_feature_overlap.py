"""§8f-4: wide-channel feature overlap (the body of OverlapCorresponder.post_atten_inject, corresponder.py:236-295, with its
early return bypassed — oracle/ref_shim.py::post_atten_inject_body) and the cell-similarity weighting of taichi_cells_overlap
(corr_utils.py:110-134, executed as plain Python through the shim).  Fixture: oracle/make_golden.py --only-features."""
import numpy as np
import pytest
import torch

import srx_oracle as O
from helpers import EngineData, assert_close, t2n


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_feature_overlap_matches_reference_body(golden, tag):
    g = golden("feature_overlap")
    out = O.feature_overlap(g[f"features_{tag}"], g[f"ids_{tag}"], ratio=float(g["ratio"]))
    assert_close(out, g[f"out_{tag}"], 1e-5, 3e-6, f"feature overlap {tag}")
    assert np.abs(g[f"out_{tag}"] - g[f"features_{tag}"]).max() > 0.1          # the fixture is not a no-op


def test_oracle_cells_overlap_matches_taichi_kernel(golden):
    g = golden("feature_overlap")
    new = O.cells_overlap(g["cells_ids"], g["cells_values"], g["cells_contrib"])
    assert_close(new, g["cells_new"], 1e-5, 3e-6, "cells overlap")


def test_oracle_nearest_index_is_torch_nearest():
    """The index rule both the oracle and the kernels use for F.interpolate(mode='nearest')."""
    import torch.nn.functional as F
    for n_in, n_out in ((8, 64), (64, 8), (12, 96), (96, 12), (8, 4), (4, 8), (7, 13), (13, 7), (64, 1024), (5, 3)):
        src = torch.arange(n_in, dtype=torch.float32).view(1, 1, n_in, 1).expand(1, 1, n_in, 2)
        up = F.interpolate(src, size=(n_out, 2), mode="nearest")[0, 0, :, 0].long().numpy()
        assert np.array_equal(up, O.nearest_resize_index(n_out, n_in)), (n_in, n_out)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_gpu_feature_overlap_vs_reference_body(golden, tag):
    from stable_renderer_b200.corrmap import IDMap
    from stable_renderer_b200.feature import feature_overlap
    g = golden("feature_overlap")
    ids = torch.from_numpy(g[f"ids_{tag}"]).cuda()
    feats = torch.from_numpy(g[f"features_{tag}"]).cuda()
    out = feature_overlap(feats, IDMap(tensor=ids), ratio=float(g["ratio"]))
    assert out.shape == feats.shape and out.dtype == feats.dtype and out.data_ptr() != feats.data_ptr()
    assert_close(t2n(out), g[f"out_{tag}"], 2e-5, 5e-6, f"feature overlap {tag}")


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, (2e-5, 5e-6)), (torch.float16, (1e-2, 1e-2)), (torch.bfloat16, (1e-2, 1e-2))])
@pytest.mark.parametrize("case", ["sane_320", "sane_1280", "quirk_640"])
def test_gpu_feature_overlap_vs_oracle_wide_channels(dtype, tol, case):
    """The production channel widths (320 / 640 / 1280), the intended up-sampling size (H, W) as well as the reference's
    (W, 4), frame indices that permute the batch, several pixels per feature cell."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.corrmap import IDMap
    from stable_renderer_b200.feature import feature_overlap
    F, H, hw, c, msize = {"sane_320": (4, 128, 256, 320, (128, 128)), "sane_1280": (3, 64, 64, 1280, (64, 64)),
                          "quirk_640": (3, 96, 144, 640, None)}[case]
    ids = synthetic.make_ids(F, H, H, tex_h=64, tex_w=64, frac_2048=0.05, seed=61)
    gen = torch.Generator().manual_seed(7)
    feats = torch.randn(F, hw, c, generator=gen).to(dtype)
    perm = list(range(F))[::-1]
    kw = {} if msize is None else {"map_height": msize[0], "map_width": msize[1]}
    want = O.feature_overlap(feats.float().numpy(), ids.numpy(), ratio=0.6, frame_indices=perm, **kw)
    out = feature_overlap(feats.cuda(), IDMap(tensor=ids.cuda(), frame_indices=perm), ratio=0.6, map_size=msize, key_capacity=64 * 64)
    assert_close(t2n(out), want, tol[0], tol[1], case)


@pytest.mark.gpu
def test_gpu_feature_overlap_reuses_buckets_across_layers_and_steps():
    """One id batch, three attention layers of two sizes and dtypes, two denoise steps: the bucketing pass runs once per
    (size, key capacity) and every later call reads the cached buckets; `IDMap.invalidate()` drops them."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.corrmap import IDMap
    from stable_renderer_b200.feature import feature_overlap
    F, H = 4, 128
    ids = synthetic.make_ids(F, H, H, tex_h=64, tex_w=64, frac_2048=0.05, seed=62)
    idm = IDMap(tensor=ids.cuda())
    gen = torch.Generator().manual_seed(11)
    seen = []
    for step in range(2):
        for hw, c, dtype, tol in ((256, 320, torch.float32, (2e-5, 5e-6)), (256, 640, torch.float16, (1e-2, 1e-2)),
                                  (64, 1280, torch.float32, (2e-5, 5e-6))):
            feats = torch.randn(F, hw, c, generator=gen).to(dtype)
            info = {}
            out = feature_overlap(feats.cuda(), idm, ratio=0.6, map_size=(H, H), key_capacity=64 * 64, info=info)
            want = O.feature_overlap(feats.float().numpy(), ids.numpy(), ratio=0.6, map_height=H, map_width=H)
            assert_close(t2n(out), want, tol[0], tol[1], f"step {step} hw {hw} c {c}")
            assert info["rows_gathered"] >= F * hw
            seen.append(info["buckets_reused"])
    assert seen == [False, True, False, True, True, True]
    assert len(idm._feature_buckets) == 2
    # new content in the same tensor: without invalidate() the cached buckets would be stale
    ids2 = synthetic.make_ids(F, H, H, tex_h=64, tex_w=64, frac_2048=0.05, seed=63)
    idm.tensor.copy_(ids2.cuda())
    idm.invalidate()
    assert not idm._feature_buckets
    feats = torch.randn(F, 256, 320, generator=gen)
    info = {}
    out = feature_overlap(feats.cuda(), idm, ratio=0.6, map_size=(H, H), key_capacity=64 * 64, info=info)
    assert info["buckets_reused"] is False
    assert_close(t2n(out), O.feature_overlap(feats.numpy(), ids2.numpy(), ratio=0.6, map_height=H, map_width=H), 2e-5, 5e-6, "after invalidate")
    # cache_buckets=False keeps nothing
    idm.invalidate()
    feature_overlap(feats.cuda(), idm, ratio=0.6, map_size=(H, H), key_capacity=64 * 64, cache_buckets=False)
    assert not idm._feature_buckets


@pytest.mark.gpu
def test_gpu_post_atten_inject_switch_and_errors():
    from stable_renderer_b200 import _lib, synthetic
    from stable_renderer_b200.corresponder import OverlapCorresponder
    from stable_renderer_b200.corrmap import IDMap
    from stable_renderer_b200.feature import feature_overlap
    ids = synthetic.make_ids(2, 64, 64, tex_h=32, tex_w=32, seed=5).cuda()
    feats = torch.randn(2, 64, 32).cuda()
    ed = EngineData(id_maps=IDMap(tensor=ids))
    assert OverlapCorresponder().post_atten_inject(None, ed, feats, 12) is feats              # the reference's behaviour
    oc = OverlapCorresponder(enable_post_attn_inject=True)
    assert oc.post_atten_inject(None, ed, feats, 3) is feats                                   # layers 0..10 are skipped
    got = oc.post_atten_inject(None, ed, feats, 12)
    want = O.feature_overlap(t2n(feats), t2n(ids).astype(np.int32), ratio=0.6)
    assert_close(t2n(got), want, 2e-5, 5e-6, "post_atten_inject")
    with pytest.raises(ValueError):
        feature_overlap(torch.randn(2, 60, 32).cuda(), IDMap(tensor=ids))                      # hw is not a square
    with pytest.raises(_lib.SrxError):
        feature_overlap(feats, IDMap(tensor=ids), key_capacity=8)                              # vertex ids beyond the capacity
    with pytest.raises(IndexError):
        feature_overlap(feats, IDMap(tensor=ids, frame_indices=[0, 5]))                        # frame index outside the batch


@pytest.mark.gpu
def test_gpu_cells_overlap_vs_taichi_kernel(golden):
    from stable_renderer_b200.feature import taichi_cells_overlap
    g = golden("feature_overlap")
    new = torch.zeros(g["cells_values"].shape, dtype=torch.float32, device="cuda")
    taichi_cells_overlap(torch.from_numpy(g["cells_ids"]).cuda(), torch.from_numpy(g["cells_values"]).cuda(), new,
                         torch.from_numpy(g["cells_contrib"]).cuda())
    assert_close(t2n(new), g["cells_new"], 2e-5, 5e-6, "cells overlap")


@pytest.mark.gpu
def test_gpu_cells_overlap_vs_oracle_wide():
    """2 frames of 64x64 pixels, 8x8 latent cells' worth of 64-pixel runs, 320 channels, background pixels included."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.feature import taichi_cells_overlap
    ids = synthetic.make_ids(2, 64, 64, tex_h=12, tex_w=12, seed=91).reshape(2, 64 * 64, 4)
    gen = torch.Generator().manual_seed(2)
    vals = torch.randn(2, 64, 320, generator=gen)
    contrib = torch.rand(2, 64 * 64, generator=gen) / 64
    want = O.cells_overlap(ids.numpy(), vals.numpy(), contrib.numpy())
    new = torch.zeros_like(vals).cuda()
    taichi_cells_overlap(ids.cuda(), vals.cuda(), new, contrib.cuda())
    assert_close(t2n(new), want, 5e-5, 1e-5, "cells overlap wide")
