"""CPU-only checks: the C-ABI library loads and exports every symbol include/srx.h declares; host-side logic
(schedulers, correspondence-map inspection, argument validation) matches the reference-generated fixtures.
No kernel is launched here."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from stable_renderer_b200 import build, _lib
    build.build_lib()
    return _lib.load()


def _header_functions():
    src = open(os.path.join(ROOT, "include", "srx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(srx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from stable_renderer_b200 import _lib
    declared = _header_functions()
    assert len(declared) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (srx_[a-z0-9_]+)", out))
    missing = [f for f in declared if f not in exported]
    assert not missing, f"declared in include/srx.h but not exported: {missing}"
    assert sorted(_lib.exported_symbols()) == declared          # the ctypes table covers the whole header
    assert lib.srx_version() == 100


def test_library_is_sm100a_only():
    from stable_renderer_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback():
    from stable_renderer_b200 import _lib
    from stable_renderer_b200.corrmap import IDMap
    from stable_renderer_b200.corresponder import OverlapCorresponder
    from stable_renderer_b200.plan import OverlapPlan
    from helpers import Ctx, EngineData
    ids = torch.zeros(1, 8, 8, 4, dtype=torch.int32)
    with pytest.raises(_lib.SrxUnavailable):
        OverlapPlan(ids, (1, 4, 1, 1))
    with pytest.raises(_lib.SrxUnavailable):
        OverlapCorresponder().step_finished(EngineData(IDMap(tensor=ids)), Ctx(torch.zeros(1, 4, 1, 1), 900))
    # the round-2 entry points refuse CPU tensors the same way
    from stable_renderer_b200.feature import feature_overlap, taichi_cells_overlap
    from stable_renderer_b200.loaders import CreateNoiseSequenceFromIdMap
    from stable_renderer_b200.overlap.johnny import overlap as johnny_overlap  # noqa: F401  (imports without a GPU)
    with pytest.raises(_lib.SrxUnavailable):
        feature_overlap(torch.zeros(1, 4, 16), IDMap(tensor=ids))
    with pytest.raises(_lib.SrxUnavailable):
        taichi_cells_overlap(torch.zeros(1, 64, 4), torch.zeros(1, 4, 8), torch.zeros(1, 4, 8), torch.zeros(1, 64))
    with pytest.raises(_lib.SrxUnavailable):
        CreateNoiseSequenceFromIdMap()(IDMap(tensor=torch.ones(1, 64, 64, 4, dtype=torch.int32)), 1)


def test_argument_validation_without_gpu(lib):
    import ctypes as C
    from stable_renderer_b200 import _lib
    d = _lib.srx_plan_desc()
    h = C.c_void_p()
    assert lib.srx_plan_create(C.byref(h), C.byref(d), None, None) == _lib.SRX_ERR_INVALID
    assert b"dtype" in lib.srx_last_error()
    with pytest.raises(ValueError):
        _lib.check(lib.srx_bake_update(None, None))
    ld = _lib.srx_legacy_desc()
    assert lib.srx_legacy_workspace_bytes(C.byref(ld)) == -1
    ld.id_dtype, ld.frames, ld.height, ld.width, ld.channels, ld.lat_h, ld.lat_w = _lib.SRX_I16, 4, 32, 32, 4, 4, 4
    assert lib.srx_legacy_workspace_bytes(C.byref(ld)) > 0
    assert lib.srx_bake_workspace_bytes(9, 512 * 512, 4, 0) >= 9 * 512 * 512 * 4


def test_scheduler_matches_reference_table(golden):
    from stable_renderer_b200.overlap import Scheduler
    names = ("constant", "linear", "exponential", "cosine")
    for itype, power, step, ts, want in golden("legacy_scheduler")["table"]:
        s = Scheduler(every_step=2, start_step=2, end_step=40, start_timestep=100, end_timestep=900,
                      interpolate_begin=0.9, interpolate_end=0.2, power=power, interpolate_type=names[int(itype)],
                      no_interpolate_return=0.05)
        assert float(s(int(step), ts)) == pytest.approx(want, rel=1e-12, abs=1e-15)
        assert float(s(int(step), torch.tensor(float(ts), dtype=torch.float64))) == pytest.approx(want, rel=1e-12, abs=1e-15)


def test_correspondence_map_inspection_matches_reference(golden):
    from stable_renderer_b200.overlap import CorrespondenceMap
    g = golden("legacy_corrmap")
    cm = CorrespondenceMap.from_ids(torch.from_numpy(g["ids"]))
    assert cm.size == tuple(g["size"]) and cm.num_frames == g["ids"].shape[0]
    m = cm.Map
    assert np.array_equal(np.array(list(m.keys())), g["keys"])
    assert np.array_equal(np.array([len(v) for v in m.values()]), g["lens"])
    flat = np.array([(p[0], p[1], f) for v in m.values() for (p, f) in v])
    assert np.array_equal(flat, g["traces"])
    assert len(cm) == len(g["keys"])
    cm.merge_nearby(4)
    mg = golden("legacy_corrmap_merge4")
    m = cm.Map
    assert np.array_equal(np.array(list(m.keys())), mg["keys"])
    assert np.array_equal(np.array([len(v) for v in m.values()]), mg["lens"])
    assert len(cm) == len(mg["keys"])


def test_correspondence_map_from_directory(tmp_path, golden):
    from stable_renderer_b200.overlap import CorrespondenceMap
    from stable_renderer_b200.corrmap import IDMap
    g = golden("legacy_corrmap")
    d = tmp_path / "id"
    d.mkdir()
    for f in (2, 0, 3, 1):
        np.save(d / f"id_{f * 5}.npy", g["ids"][f])
    cm = CorrespondenceMap.FromExisting(str(tmp_path), num_frames=3)
    assert cm.num_frames == 3 and np.array_equal(cm.ids.numpy(), g["ids"][:3])
    idm = IDMap.from_directory(str(d))
    assert idm.frame_indices == [0, 5, 10, 15] and np.array_equal(idm.tensor.numpy(), g["ids"])
    idm2 = IDMap.from_directory(str(d), 1, 2, use_frame_indices_from_filename=False)
    assert idm2.frame_indices == [1, 2] and np.array_equal(idm2.tensor.numpy(), g["ids"][1:3])


def test_reference_corr_map_pkl_loads_into_id_buffers(golden, tmp_path):
    """`corr_map.pkl` as written by the reference's own `save_cache` (fixture: oracle/make_golden.py::legacy_cases) -> id buffers,
    same dict in the same order; the unpickler runs no reference code and refuses anything but builtins / numpy."""
    import pickle
    import shutil
    from stable_renderer_b200.overlap import CorrespondenceMap
    g = golden("legacy_corrmap")
    src = os.path.join(os.path.dirname(__file__), "golden", "legacy_corr_map.pkl")
    cm = CorrespondenceMap.LoadFromCache(src)
    assert np.array_equal(cm.ids.numpy(), g["ids"].astype(np.int32))
    assert cm.num_frames == g["ids"].shape[0] and (cm.width, cm.height) == tuple(g["size"])
    m = cm.Map
    assert np.array_equal(np.array(list(m.keys())), g["keys"])
    assert np.array_equal(np.array([(p[0], p[1], f) for v in m.values() for (p, f) in v]), g["traces"])
    shutil.copy(src, tmp_path / "corr_map.pkl")                         # a directory holding the cache
    assert len(CorrespondenceMap.LoadFromCache(str(tmp_path))) == len(g["keys"])
    (tmp_path / "id").mkdir()                                           # FromExisting finds it like the reference does
    for where in (str(tmp_path), str(tmp_path / "id"), str(tmp_path / "corr_map.pkl")):
        assert np.array_equal(CorrespondenceMap.FromExisting(where).ids.numpy(), cm.ids.numpy())
    with pytest.raises(FileNotFoundError):
        CorrespondenceMap.FromExisting(str(tmp_path / "id"), enable_cache=False)     # no id dumps there
    cm.save_cache(str(tmp_path / "own.pkl"))                            # this package's own cache format
    assert np.array_equal(CorrespondenceMap.LoadFromCache(str(tmp_path / "own.pkl")).ids.numpy(), cm.ids.numpy())

    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))
    with open(tmp_path / "evil.pkl", "wb") as f:
        pickle.dump(Evil(), f)
    with pytest.raises(pickle.UnpicklingError):
        CorrespondenceMap.LoadFromCache(str(tmp_path / "evil.pkl"))
    with open(tmp_path / "other.pkl", "wb") as f:
        pickle.dump([1, 2, 3], f)
    with pytest.raises(ValueError):
        CorrespondenceMap.LoadFromCache(str(tmp_path / "other.pkl"))
    with pytest.raises(FileNotFoundError):
        CorrespondenceMap.LoadFromCache(str(tmp_path / "missing.pkl"))


def test_idmap_masks_and_shapes(golden):
    from stable_renderer_b200.corrmap import IDMap
    g = golden("step_sq64_r8")
    idm = IDMap(tensor=torch.from_numpy(g["ids"]))
    assert np.array_equal(idm.masks.numpy(), g["masks"])
    assert len(idm) == 4 and idm.frame_indices == [0, 1, 2, 3]
    one = IDMap(tensor=torch.from_numpy(g["ids"][0]), frame_indices=7)
    assert one.frame_indices == [7] and tuple(one.tensor.shape) == (1, 64, 64, 4)
    with pytest.raises(ValueError):
        IDMap(tensor=torch.zeros(4, 4))
    with pytest.raises(ValueError):
        IDMap.from_tensor([0, 1], torch.zeros(3, 4, 4, 4))


def test_algorithm_objects_match_reference_weights(golden):
    """The per-trace OverlapAlgorithm.overlap protocol method (torch, any device) against the oracle's dense weights."""
    import srx_oracle as O
    from stable_renderer_b200.overlap import overlap_algorithm_factory
    rng = np.random.default_rng(0)
    L = 7
    lat = torch.from_numpy(rng.standard_normal((L, 1, 4)))
    fs, xs, ys = list(rng.integers(0, 5, L)), list(rng.integers(0, 9, L)), list(rng.integers(0, 9, L))
    vn_map = torch.from_numpy(rng.random((5, 9, 9, 1)))
    for strat in O.STRATEGIES:
        got = overlap_algorithm_factory(strat).overlap(lat, fs, xs, ys, view_normal_map=vn_map)
        vn = vn_map.numpy()[fs, ys, xs].reshape(-1)
        Wm = O.strategy_weights(strat, np.array(fs), np.array(ys), np.array(xs), vn)
        want = (Wm @ lat.numpy().reshape(L, -1)) / Wm.sum(axis=0).reshape(-1, 1)
        np.testing.assert_allclose(got.numpy().reshape(L, -1), want, rtol=1e-12)
