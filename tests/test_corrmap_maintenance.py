"""Legacy `CorrespondenceMap` maintenance (reference data_classes/correspondence_map.py:177-286): dropouts, merge, cache —
against fixtures produced by the reference's own dict-based implementation."""
import numpy as np
import pytest
import torch

import srx_oracle as O


def _keys(d):
    return np.array(list(d.keys()), dtype=np.int64)


@pytest.mark.parametrize("tag,merge", [("full", 0), ("merge4", 4)])
def test_oracle_dropouts_match_reference(golden, tag, merge):
    g, c = golden("legacy_corrmap_dropout"), golden("legacy_corrmap")
    tr = O.correspondence_traces(c["ids"], merge)
    a = O.traces_dropout_index(tr, float(g["probability"]), int(g["seed"]))
    b = O.traces_dropout_in_rectangle(tr, tuple(map(tuple, g["rect"].tolist())), int(g["at_frame"]))
    assert np.array_equal(_keys(a), g[f"{tag}_index_keys"])
    assert np.array_equal(_keys(b), g[f"{tag}_rect_keys"])


def test_cache_round_trip_cpu(golden, tmp_path):
    from stable_renderer_b200.overlap import CorrespondenceMap
    c = golden("legacy_corrmap")
    m = CorrespondenceMap(torch.from_numpy(c["ids"]))
    m.merge_nearby(4)
    m.save_cache(str(tmp_path))                       # a directory -> corr_map.pkl inside it
    assert (tmp_path / "corr_map.pkl").exists()
    m2 = CorrespondenceMap.LoadFromCache(str(tmp_path))
    assert torch.equal(m2.ids, m.ids) and m2.merge_len == 4 and m2.size == m.size
    with pytest.raises(FileNotFoundError):
        CorrespondenceMap.LoadFromCache(str(tmp_path / "nope.pkl"))


@pytest.mark.gpu
@pytest.mark.parametrize("tag,merge", [("full", 0), ("merge4", 4)])
@pytest.mark.parametrize("id_dtype", [torch.int16, torch.int32])
def test_gpu_dropouts_match_reference(golden, tag, merge, id_dtype):
    from stable_renderer_b200.overlap import CorrespondenceMap
    g, c = golden("legacy_corrmap_dropout"), golden("legacy_corrmap")
    ids = torch.from_numpy(c["ids"]).to(id_dtype)

    def fresh():
        m = CorrespondenceMap(ids.clone().cuda())
        if merge:
            m.merge_nearby(merge)
        return m
    a = fresh()
    a.dropout_index(float(g["probability"]), int(g["seed"]))
    assert np.array_equal(_keys(a.Map), g[f"{tag}_index_keys"])          # same keys, same insertion order
    b = fresh()
    b.dropout_in_rectangle(tuple(map(tuple, g["rect"].tolist())), int(g["at_frame"]))
    assert np.array_equal(_keys(b.Map), g[f"{tag}_rect_keys"])
    assert len(b) == len(g[f"{tag}_rect_keys"])
    # surviving traces are untouched
    want = O.traces_dropout_in_rectangle(O.correspondence_traces(c["ids"], merge), tuple(map(tuple, g["rect"].tolist())),
                                         int(g["at_frame"]))
    # (within a merged key the reference lists the tracks grouped by their original key; `Map` lists them in pixel order)
    got = {k: sorted((p[0], p[1], f) for p, f in v) for k, v in b.Map.items()}
    assert got == {k: sorted(v) for k, v in want.items()}
    # nothing to drop / everything dropped
    z = fresh()
    z.dropout_index(0.0, 1)
    assert len(z) == len(_keys(O.correspondence_traces(c["ids"], merge)))
    z.dropout_index(1.0, 1)
    assert len(z) == 0 and not z.ids.any()


@pytest.mark.parametrize("tag", ["13", "3"])
def test_build_view_normal_map_matches_reference(golden, tag):
    """`build_view_normal_map` (overlap/utils.py:56-102) for a [1,3] view vector (normalised per component by the reference's
    `F.normalize(dim=0)`) and a [3] vector (normalised to unit length): oracle and host helper against the reference's output."""
    from PIL import Image
    from stable_renderer_b200.overlap import build_view_normal_map
    g = golden("legacy_view_normal")
    o = O.build_view_normal_map(g["images"].astype(np.float32) / 255.0, g["v" + tag])
    assert np.array_equal(o, g["out" + tag])
    pil = [Image.fromarray(a, mode="RGB") for a in g["images"]]
    got = build_view_normal_map(pil, torch.from_numpy(g["v" + tag])).numpy()
    assert got.shape == g["out" + tag].shape and np.array_equal(got, g["out" + tag])
    with pytest.raises(TypeError):
        build_view_normal_map(tuple(pil), torch.from_numpy(g["v" + tag]))
