"""Frame ingest (`RenderManager._save_frame_data`, renderManager.py:877-948) and the closer-pixel merge (:121-133) against a
fixture produced by the reference's own function bodies on stand-in textures (oracle/make_golden.py::ingest_cases,
oracle/ref_shim.py::render_manager_functions)."""
import os

import numpy as np
import pytest
import torch

from oracle import srx_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "frame_ingest.npz")
HALF_KEYS = ("color", "normal_depth", "noise", "canny")
ATT = ("color", "ids", "pos", "normal_depth", "noise", "canny")
EXACT = ("color_maps", "masks", "id_maps", "pos_maps", "normal_maps", "depth_maps", "canny_maps")
HALF_OUT = ("color_maps", "masks", "normal_maps", "depth_maps", "canny_maps")


def _att(g, prefix):
    return {k: (g[f"{prefix}_{k}"].view(np.float16) if k in HALF_KEYS else g[f"{prefix}_{k}"]) for k in ATT}


def test_oracle_frame_ingest_matches_reference_body():
    g = np.load(GOLD)
    for f in range(2):
        out = O.frame_ingest(bg_noise=g["bg_noise"][0], flip=True, **_att(g, f"src{f}"))
        for k in EXACT:
            ref = g[k][f].view(np.float16) if k in HALF_OUT else g[k][f]
            assert np.array_equal(out[k], ref), k
        # fp32 sums in another order; style statistics are fp16 on both sides
        assert np.allclose(out["noise_maps"], g["noise_maps"][f], rtol=2e-5, atol=2e-5)


def test_oracle_merge_closer_matches_reference_body():
    g = np.load(GOLD)
    H, W = g["temp_depth"].shape
    temp = dict(color=np.zeros((H, W, 4), np.float16), ids=np.zeros((H, W, 4), np.int32), pos=np.zeros((H, W, 3), np.float32),
                normal=np.zeros((H, W, 3), np.float16), depth=np.zeros((H, W), np.float16), noise=np.zeros((H, W, 4), np.float16),
                canny=np.zeros((H, W, 3), np.float16))
    for d in range(3):
        O.gbuffer_merge_closer(temp, flip=True, **_att(g, f"draw{d}"))
    for k, v in temp.items():
        ref = g[f"temp_{k}"]
        assert np.array_equal(v.view(np.uint16) if v.dtype == np.float16 else v, ref), k


@pytest.mark.gpu
def test_gpu_frame_ingest_matches_reference_replay():
    from stable_renderer_b200.corrmap import IDMap
    from stable_renderer_b200.ingest import FrameIngest
    g = np.load(GOLD)
    H, W = g["bg_noise"].shape[1:3]
    ing = FrameIngest(H, W, capacity=1, bg_noise=torch.from_numpy(g["bg_noise"]).cuda())     # capacity 1: the second frame grows it
    for f in range(2):
        a = {k: torch.from_numpy(v.copy()).cuda() for k, v in _att(g, f"src{f}").items()}
        ing.save_frame_data(100 + f, flip=True, **a)
    d = ing.data
    assert d["frame_indices"] == [100, 101] and isinstance(d["id_maps"], IDMap) and len(ing) == 2
    for k in EXACT:
        got = (d[k].tensor if k == "id_maps" else d[k]).cpu().numpy()
        ref = g[k].view(np.float16) if k in HALF_OUT else g[k]
        assert got.shape == ref.shape and np.array_equal(got, ref), k
    assert d["noise_maps"].shape == (2, 4, H // 8, W // 8)
    assert np.allclose(d["noise_maps"].cpu().numpy(), g["noise_maps"], rtol=2e-5, atol=2e-5)
    ing.clear()
    assert len(ing) == 0 and ing.data["frame_indices"] == []


@pytest.mark.gpu
def test_gpu_frame_ingest_vs_oracle_unflipped_f32_canny_and_partial_attachments():
    from stable_renderer_b200.ingest import FrameIngest
    gen = torch.Generator().manual_seed(5)
    H, W = 64, 40                                                       # W % 64 != 0: pooling groups straddle rows
    color = torch.rand(H, W, 4, generator=gen).half()
    ids = torch.randint(0, 1 << 20, (H, W, 4), generator=gen, dtype=torch.int32)
    pos = torch.randn(H, W, 3, generator=gen)
    nd = torch.rand(H, W, 4, generator=gen).half()
    noise = torch.randn(H, W, 4, generator=gen).half()
    canny = torch.rand(H, W, 3, generator=gen)                          # f32 attachment
    ing = FrameIngest(H, W)
    bg = ing.GlobalBGNoise
    ing.save_frame_data(0, color.cuda(), ids.cuda(), pos.cuda(), nd.cuda(), noise.cuda(), canny.cuda(), flip=False)
    ing.save_frame_data(1, color.cuda(), ids.cuda(), flip=True)         # only the required attachments
    ref = O.frame_ingest(color.numpy(), ids.numpy(), pos.numpy(), nd.numpy(), noise.numpy(), canny.numpy(), bg[0].cpu().numpy(), flip=False)
    d = ing.data
    for k in EXACT:
        got = (d[k].tensor if k == "id_maps" else d[k])[0].cpu().numpy()
        assert np.array_equal(got, ref[k]), k
    assert d["canny_maps"].dtype == torch.float32
    assert np.allclose(d["noise_maps"][0].cpu().numpy(), ref["noise_maps"], rtol=2e-5, atol=2e-5)
    assert torch.equal(d["color_maps"][1].cpu(), color.flip(0)[..., :3]) and torch.equal(d["id_maps"].tensor[1].cpu(), ids.flip(0))
    with pytest.raises(ValueError):
        ing.save_frame_data(2, color.cuda()[:8], ids.cuda())


@pytest.mark.gpu
def test_gpu_merge_closer_matches_reference_replay():
    from stable_renderer_b200.ingest import GBufferTemp
    g = np.load(GOLD)
    H, W = g["temp_depth"].shape
    t = GBufferTemp(H, W)
    for d in range(3):
        a = {k: torch.from_numpy(v.copy()).cuda() for k, v in _att(g, f"draw{d}").items()}
        t.merge_closer(flip=True, **a)
    for k in ("color", "ids", "pos", "normal", "depth", "noise", "canny"):
        got = getattr(t, k).cpu().numpy()
        ref = g[f"temp_{k}"]
        assert np.array_equal(got.view(np.uint16) if got.dtype == np.float16 else got, ref), k
    t.clear()
    assert float(t.depth.float().abs().sum()) == 0.0
