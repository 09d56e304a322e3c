"""Legacy `CorrMapLatentNoiseInitializer` (reference: legacy_codes/nodes/latent.py:10-40) against fixtures produced by the
reference node itself (oracle/make_golden.py::latent_init_cases)."""
import os

import numpy as np
import pytest
import torch

from oracle import srx_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "legacy_latent_init.npz")


def _cases():
    g = np.load(GOLD)
    for c in g["cases"].tolist():
        tag, w, h, b, s = c.split(":")
        yield g, tag, int(w), int(h), int(b), int(s)


def test_oracle_matches_reference_node():
    for g, tag, w, h, b, s in _cases():
        ids = g["ids"]
        traces = O.correspondence_traces(ids, 4 if tag == "merge4" else 0)
        n = sum(1 for t in traces.values() if len(t) > 1)
        assert n == int(g[tag + "_n_traces"])
        bl, bn, rows = O.corrmap_latent_noise_draws(s, ids.shape[2], ids.shape[1], n)
        lat, noi = O.corrmap_latent_noise_init(traces, ids.shape[2], ids.shape[1], w, h, b, bl, bn, rows)
        name = f"{tag}_{w}x{h}_b{b}_s{s}"
        assert np.array_equal(lat, g[name + "_samples"]) and np.array_equal(noi, g[name + "_noise"]), name


def test_trace_rows_equal_the_loop_of_randn4():
    from stable_renderer_b200.overlap.latent import trace_rows
    torch.manual_seed(99)
    a = trace_rows(777)
    torch.manual_seed(99)
    b = torch.stack([torch.stack([torch.randn(4), torch.randn(4)]) for _ in range(777)])
    assert torch.equal(a, b)
    assert trace_rows(0).shape == (0, 2, 4)


@pytest.mark.gpu
def test_gpu_node_matches_reference_fixtures():
    from stable_renderer_b200.overlap import CorrespondenceMap, CorrMapLatentNoiseInitializer
    for g, tag, w, h, b, s in _cases():
        cm = CorrespondenceMap(torch.from_numpy(g["ids"]).cuda())
        if tag == "merge4":
            cm.merge_nearby(4)
        (d,) = CorrMapLatentNoiseInitializer()(w, h, b, s, cm)
        name = f"{tag}_{w}x{h}_b{b}_s{s}"
        assert d["samples"].shape == g[name + "_samples"].shape
        assert np.array_equal(d["samples"].cpu().numpy(), g[name + "_samples"]), name
        assert np.array_equal(d["noise"].cpu().numpy(), g[name + "_noise"]), name


@pytest.mark.gpu
def test_gpu_node_matches_oracle_on_int32_ids_and_rejects_small_batch():
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.overlap import CorrespondenceMap, CorrMapLatentNoiseInitializer
    T, H, W = 5, 96, 64
    ids = synthetic.make_ids(T, H, W, tex_h=40, tex_w=40, seed=3, legacy_layout=True, dtype=torch.int32)
    ids[2, 10:20, 10:20] = torch.tensor([7, 3, 4000, 90000], dtype=torch.int32)      # a large one-frame trace
    ids[4, 0, 0] = torch.tensor([9, 9, 9, 9], dtype=torch.int32)                        # a singleton key: skipped
    for merge in (0, 3):
        cm = CorrespondenceMap(ids.cuda(), merge_len=merge)
        (d,) = CorrMapLatentNoiseInitializer()(W * 2, H * 2, T + 1, 31, cm)
        traces = O.correspondence_traces(ids.numpy(), merge)
        n = sum(1 for t in traces.values() if len(t) > 1)
        bl, bn, rows = O.corrmap_latent_noise_draws(31, W, H, n)
        lat, noi = O.corrmap_latent_noise_init(traces, W, H, W * 2, H * 2, T + 1, bl, bn, rows)
        assert np.array_equal(d["samples"].cpu().numpy(), lat) and np.array_equal(d["noise"].cpu().numpy(), noi)
    with pytest.raises(IndexError):
        CorrMapLatentNoiseInitializer()(W, H, T - 1, 0, CorrespondenceMap(ids.cuda()))
