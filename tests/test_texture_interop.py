"""-m gpu: the zero-copy interop path on plain CUDA arrays — what registered GL textures look
like to CUDA; no GL context exists on the GPU box.  `Texture.tensor()` / `set_data()` (texture.py:221-254, 326-408), the ingest
and closer-pixel merge reading the mapped arrays directly (renderManager.py:877-948, 121-133), and the atlas upload into the
layers of a 2D-array texture (corrmap.py:443-489)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "frame_ingest.npz")
HALF_KEYS = ("color", "normal_depth", "noise", "canny")
ATT = ("color", "ids", "pos", "normal_depth", "noise", "canny")


def test_texture_tensor_and_set_data_round_trips():
    from stable_renderer_b200.texture import Texture
    gen = torch.Generator().manual_seed(1)
    H, W = 24, 40
    tex = Texture.offscreen(W, H, 4, torch.float16)
    data = torch.randn(H, W, 4, generator=gen).half().cuda()
    tex.set_data(data)                                              # rows as given = GL order
    assert torch.equal(tex.tensor(flip=False), data)
    assert torch.equal(tex.tensor(flip=True), data.flip(0))         # flip fused into the copy (texture.py:236,253)
    assert tex.tensor(update=False, flip=True).data_ptr() == tex.tensor(update=False, flip=True).data_ptr()
    assert torch.equal(tex.tensor(update=False, flip=False), data)  # cached tensor, other orientation
    # region write with offsets (glTexSubImage2D semantics), the rest untouched
    patch = torch.randn(5, 7, 4, generator=gen).half().cuda()
    tex.set_data(patch, xOffset=3, yOffset=2, width=7, height=5)
    want = data.clone()
    want[2:7, 3:10] = patch
    assert torch.equal(tex.tensor(flip=False), want)
    # RGB data for an RGBA texture: alpha = 1 (texture.py:379-380); float32 data is cast; numpy accepted
    rgb = torch.rand(H, W, 3, generator=gen)
    tex.set_data(rgb.numpy())
    got = tex.tensor(flip=False)
    assert torch.equal(got[..., :3], rgb.half().cuda()) and bool((got[..., 3] == 1).all())
    # single channel is repeated (texture.py:377-378), a transposed [W,H,C] input is transposed back (:370-371)
    mono = torch.rand(H, W, 1, generator=gen).half()
    tex.set_data(mono.cuda())
    assert torch.equal(tex.tensor(flip=False), mono.expand(-1, -1, 4).cuda())
    tex.set_data(data.transpose(0, 1).contiguous())
    assert torch.equal(tex.tensor(flip=False), data)
    with pytest.raises(Exception):
        tex.set_data(torch.zeros(3, 3, 4))
    tex.clear()


def test_three_channel_and_integer_textures():
    """RGB32F (position / canny, renderManager.py:268,352) is a four-channel CUDA array; `tensor()` still hands out [H,W,3].
    The RGBA_32I id attachment comes back as int32."""
    from stable_renderer_b200.texture import Texture
    gen = torch.Generator().manual_seed(2)
    H, W = 16, 24
    pos = Texture.offscreen(W, H, 3, torch.float32)
    p = torch.randn(H, W, 3, generator=gen).cuda()
    pos.set_data(p)
    assert pos.tensor(flip=False).shape == (H, W, 3) and torch.equal(pos.tensor(flip=False), p)
    ids = Texture.offscreen(W, H, 4, torch.int32)
    i = torch.randint(-5, 1 << 30, (H, W, 4), generator=gen, dtype=torch.int32).cuda()
    ids.set_data(i)
    assert ids.tensor().dtype == torch.int32 and torch.equal(ids.tensor(flip=True), i.flip(0))


def _textures_from(att, H, W):
    from stable_renderer_b200.texture import Texture
    spec = {"color": (4, torch.float16), "ids": (4, torch.int32), "pos": (3, torch.float32), "normal_depth": (4, torch.float16),
            "noise": (4, torch.float16), "canny": (3, att["canny"].dtype)}
    out = {}
    for k, (ch, dt) in spec.items():
        t = Texture.offscreen(W, H, ch, dt)
        t.set_data(att[k])
        out[k] = t
    return out


@pytest.mark.parametrize("canny_dtype", [torch.float16, torch.float32])
def test_ingest_from_arrays_equals_ingest_from_tensors(canny_dtype):
    """The zero-copy ingest reads the attachment arrays in place; its batches must be bit-identical to those of the
    linear-buffer ingest (itself pinned to the reference replay in test_frame_ingest.py)."""
    from stable_renderer_b200.ingest import FrameIngest
    g = np.load(GOLD)
    H, W = g["bg_noise"].shape[1:3]
    bg = torch.from_numpy(g["bg_noise"]).cuda()
    a_lin, a_arr = FrameIngest(H, W, capacity=1, bg_noise=bg), FrameIngest(H, W, capacity=1, bg_noise=bg)
    for f in range(2):
        att = {k: torch.from_numpy((g[f"src{f}_{k}"].view(np.float16) if k in HALF_KEYS else g[f"src{f}_{k}"]).copy()).cuda() for k in ATT}
        att["canny"] = att["canny"].to(canny_dtype)
        a_lin.save_frame_data(7 + f, flip=True, **att)
        tex = _textures_from(att, H, W)
        a_arr.save_frame_arrays(7 + f, flip=True, canny_dtype=canny_dtype, **tex)
    d0, d1 = a_lin.data, a_arr.data
    assert d1["frame_indices"] == [7, 8]
    for k in ("color_maps", "masks", "pos_maps", "normal_maps", "depth_maps", "canny_maps", "noise_maps"):
        assert torch.equal(d0[k], d1[k]), k
    assert torch.equal(d0["id_maps"].tensor, d1["id_maps"].tensor)
    # raw cudaArray handles instead of Texture objects; only the required attachments
    a_arr.save_frame_arrays(9, tex["color"].map_array(), tex["ids"].map_array(), flip=False)
    assert torch.equal(a_arr.data["id_maps"].tensor[2], tex["ids"].tensor(flip=False))
    assert float(a_arr.data["pos_maps"][2].abs().sum()) == 0.0      # absent attachment: zeros, not stale memory


def test_merge_closer_from_arrays_equals_tensors():
    from stable_renderer_b200.ingest import GBufferTemp
    g = np.load(GOLD)
    H, W = g["temp_depth"].shape
    t_lin, t_arr = GBufferTemp(H, W), GBufferTemp(H, W)
    for d in range(3):
        att = {k: torch.from_numpy((g[f"draw{d}_{k}"].view(np.float16) if k in HALF_KEYS else g[f"draw{d}_{k}"]).copy()).cuda() for k in ATT}
        t_lin.merge_closer(flip=True, **att)
        tex = _textures_from(att, H, W)
        t_arr.merge_closer_arrays(flip=True, canny_dtype=att["canny"].dtype, **tex)
    for k in ("color", "ids", "pos", "normal", "depth", "noise", "canny"):
        assert torch.equal(getattr(t_lin, k), getattr(t_arr, k)), k


@pytest.mark.parametrize("channels,shape", [(4, (64, 64)), (1, (48, 80)), (2, (96, 96))])
def test_atlas_upload_into_array_layers(channels, shape):
    """CorrespondMap.load on the device: layer i of the width x height 2D-array texture receives the bytes of
    get_map(i, order='whc') (corrmap.py:470-480) — the transpose for square atlases."""
    from stable_renderer_b200 import _lib
    from stable_renderer_b200.corrmap import CorrespondMap
    lib = _lib.load()
    k, (Ht, Wt) = 2, shape
    cm = CorrespondMap(name="t", k=k, height=Ht, width=Wt, channel_count=channels)
    gen = torch.Generator().manual_seed(3)
    cm._values.copy_(torch.randn(k * k, Ht * Wt, channels, generator=gen).half())
    arrays = []
    for _ in range(k * k):
        a = C.c_void_p()
        _lib.check(lib.srx_array_alloc(C.byref(a), Wt, Ht, channels, 16, 2))
        arrays.append(a.value)
    cm.load(arrays=arrays)
    for i in range(k * k):
        back = torch.empty(Ht, Wt, channels, dtype=torch.float16, device="cuda")
        _lib.check(lib.srx_array_to_tensor(arrays[i], back.data_ptr(), Wt, Ht, channels * 2, 0, None))
        want = cm.get_map(i, order="whc").contiguous().view(Ht, Wt, channels)   # the bytes glTexSubImage3D would read
        assert torch.equal(back, want), i
        if Ht == Wt:
            assert torch.equal(back, cm.get_map(i).transpose(0, 1))
    cm.load(arrays=arrays, transpose=False)
    back = torch.empty(Ht, Wt, channels, dtype=torch.float16, device="cuda")
    _lib.check(lib.srx_array_to_tensor(arrays[1], back.data_ptr(), Wt, Ht, channels * 2, 0, None))
    assert torch.equal(back, cm.get_map(1))
    with pytest.raises(ValueError):
        cm.load(arrays=arrays[:1])
    for a in arrays:
        lib.srx_array_free(C.c_void_p(a))
