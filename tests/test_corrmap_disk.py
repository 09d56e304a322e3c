"""On-disk atlas format: `CorrespondMap.dump` / `Load` (reference corrmap.py:738-872) against fixtures
written by the reference's own dump / Load."""
import json
import os

import numpy as np
import pytest
import torch

import srx_oracle as O


def test_oracle_dump_load_arrays_match_reference(golden):
    g = golden("corrmap_dump")
    img, fl = O.corrmap_dump_arrays(g["values"].view(np.float16), g["writtens"], 8, 16)
    assert np.array_equal(img, g["png"]) and np.array_equal(fl, g["png_written"])
    v, w = O.corrmap_load_arrays(g["png"], g["png_written"])
    assert np.array_equal(v.view(np.uint16), g["loaded_values"]) and np.array_equal(w, g["loaded_writtens"])


@pytest.mark.gpu
@pytest.mark.parametrize("zipped", [False, True])
def test_gpu_dump_load_match_reference(golden, tmp_path, zipped):
    from PIL import Image
    from stable_renderer_b200.corrmap import CorrespondMap
    g = golden("corrmap_dump")
    m = CorrespondMap(name="t", k=2, height=8, width=16, channel_count=4)
    m._values.copy_(torch.from_numpy(g["values"].view(np.float16)))
    m._writtens.copy_(torch.from_numpy(g["writtens"]))
    p = m.dump(tmp_path, name="abc", zip=zipped)
    assert os.path.basename(p) == ("abc.zip" if zipped else "abc")
    p2 = m.dump(tmp_path, name="abc", zip=zipped)                      # no force: a free name is chosen
    assert os.path.basename(p2) == ("abc_1.zip" if zipped else "abc_1")
    assert m.dump(tmp_path, name="abc", zip=zipped, force=True) == p  # force: replaced in place
    if not zipped:
        imgs = np.stack([np.array(Image.open(os.path.join(p, f"{i}.png"))) for i in range(4)])
        flags = np.stack([np.array(Image.open(os.path.join(p, f"{i}_written.png"))) for i in range(4)])
        assert np.array_equal(imgs, g["png"]) and np.array_equal(flags, g["png_written"])     # byte for byte
        assert json.load(open(os.path.join(p, "meta.json"))) == json.loads(str(g["meta"]))
    m2 = CorrespondMap.Load(p)
    assert (m2.k, m2.height, m2.width, m2.channel_count, m2.name) == (2, 8, 16, 4, "abc")
    assert np.array_equal(m2._values.cpu().numpy().view(np.uint16), g["loaded_values"])
    assert np.array_equal(m2._writtens.cpu().numpy(), g["loaded_writtens"])
    assert CorrespondMap.Load(p, name="other").name == "other"


@pytest.mark.gpu
def test_gpu_dump_quantisation_vs_oracle_all_halves():
    """every finite float16 in [-2, 2] through the quantiser: bit-exact against numpy's float16 arithmetic"""
    from stable_renderer_b200 import _lib
    bits = np.arange(0, 1 << 16, dtype=np.uint16)
    h = bits.view(np.float16)
    h = h[np.isfinite(h) & (np.abs(h.astype(np.float32)) <= 2)]
    want = np.clip(np.float16(255.0) * h, 0, 255).astype(np.uint8)
    v = torch.from_numpy(h.copy()).cuda()
    fl = torch.zeros(8, dtype=torch.bool, device="cuda")
    out = torch.empty(v.numel(), dtype=torch.uint8, device="cuda")
    of = torch.empty(8, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load().srx_atlas_quantize(v.data_ptr(), fl.data_ptr(), out.data_ptr(), of.data_ptr(), v.numel(), 8,
                                              _lib.current_stream_ptr(v.device)))
    assert np.array_equal(out.cpu().numpy(), want)
