"""Pins oracle/srx_oracle.py against fixtures produced by the reference's own code (oracle/make_golden.py)."""
import numpy as np
import pytest

import srx_oracle as O

STEP_CASES = ["step_sq64_r8", "step_sq96_r8_perm", "step_sq100_nonint", "step_sq60_to_16", "step_dupframe"]


def test_docstring_kats(golden):
    # reference source/common_utils/math_utils.py:110-128
    t = np.array([[2, 1, 4], [2, 9, 12], [6, 4, 4], [7, 3, 99], [8, 1, 3]])
    a0, _ = O.group_by_then_average(t, 0, [1, 2])
    assert np.array_equal(a0, np.array([[5, 8], [5, 8], [4, 4], [3, 99], [1, 3]], dtype=np.float32))
    a1, u1 = O.group_by_then_average(t, 1, [0])
    assert np.array_equal(a1, np.array([[5], [2], [6], [7], [5]], dtype=np.float32))
    assert np.array_equal(u1, [1, 3, 4, 9])
    g = golden("group_by_average")
    assert np.array_equal(a0, g["a0"]) and np.array_equal(a1, g["a1"]) and np.array_equal(u1, g["u1"])


def test_group_by_average_big_and_adain(golden):
    g = golden("group_by_average")
    b, ub = O.group_by_then_average(g["big"], -1, [0, 1, 2, 3])
    assert np.array_equal(ub, g["ub"])
    assert np.array_equal(b, g["b"])          # serial float32 accumulation reproduces the 1-thread reference bit for bit
    b64, _ = O.group_by_then_average(g["big"], -1, [0, 1, 2, 3], accumulate="f64")
    np.testing.assert_allclose(b64, g["b"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(O.adain(g["content"], g["style"]), g["adain"], rtol=1e-5, atol=1e-6)


def test_group_by_errors():
    t = np.zeros((3, 3))
    with pytest.raises(ValueError):
        O.group_by_then_average(t, 3, [0])
    with pytest.raises(ValueError):
        O.group_by_then_average(t, 0, [5])


@pytest.mark.parametrize("name", STEP_CASES + ["step_sphere_crop_int16"])
def test_keying_bit_exact(golden, name):
    g = golden(name)
    fi = g["frame_indices"] if "frame_indices" in g.files else None
    vsi = O.vertex_screen_info(g["ids"], fi)
    if "vsi" in g.files:
        assert vsi.shape == g["vsi"].shape
        assert np.array_equal(vsi.view(np.uint32), g["vsi"].view(np.uint32))   # bit-exact, all 7 columns
        assert np.array_equal(O.idmap_masks(g["ids"]), g["masks"])
    else:
        assert vsi.shape[0] == int(g["n_entries"])


@pytest.mark.parametrize("name", STEP_CASES + ["step_sphere_crop_int16"])
def test_overlap_step_matches_reference(golden, name):
    g = golden(name)
    fi = g["frame_indices"] if "frame_indices" in g.files else None
    out = O.overlap_step(g["x"], g["ids"], fi, ratio=float(g["ratio"]))
    np.testing.assert_allclose(out, g["out"], rtol=1e-5, atol=2e-6)
    out64 = O.overlap_step(g["x"], g["ids"], fi, ratio=float(g["ratio"]), accumulate="f64")
    np.testing.assert_allclose(out64, g["out"], rtol=1e-5, atol=2e-6)


def test_overlap_step_gate(golden):
    g = golden("step_gate_off")
    out = O.overlap_step(g["x"], g["ids"], None, ratio=0.5, timestep=float(g["timestep"]), stop_timestep=float(g["stop"]))
    assert np.array_equal(out, g["out"]) and np.array_equal(out, g["x"])


@pytest.mark.parametrize("tag", ["f16", "bf16"])
def test_overlap_step_half_tolerance(golden, tag):
    # the reference rounds through half precision (corresponder.py:317-318,351-352; math_utils.py:39-51);
    # north_star tolerance for fp16/bf16 is 1e-2
    g = golden(f"step_half_{tag}")
    out = O.overlap_step(g["x"], g["ids"], None, ratio=float(g["ratio"]))
    # the reference's bf16 arithmetic is itself 1.15e-2 * max(1, |x|) away from the exact result: 1e-2 relative + one bf16 ulp
    np.testing.assert_allclose(out, g["out"], rtol=1e-2, atol=1e-2 if tag == "f16" else 2.0 ** -6)


def test_overlap_step_out_of_range_raises():
    ids = np.zeros((1, 8, 16, 4), dtype=np.int32)   # W > H: x/H*w leaves the latent (corrmap.py:239)
    ids[0, :, :, 3] = 5
    with pytest.raises(IndexError):
        O.overlap_step(np.zeros((1, 4, 2, 4), dtype=np.float32), ids)


@pytest.mark.parametrize("mode", ["first", "replace", "first_avg", "replace_avg"])
def test_bake_masked(golden, mode):
    g = golden(f"bake_{mode}_masked")
    k, tex = int(g["k"]), int(g["tex"])
    values, writtens = O.corrmap_new(k, tex, tex, 4)
    O.corrmap_update(values, writtens, g["colors"], g["ids"], spriteID=1, materialID=0, mode=mode, masks=g["masks"],
                     inverse_masks=True, ignore_obj_mat_id=True)
    assert np.array_equal(writtens, g["writtens"])
    assert np.array_equal(values.view(np.uint16), g["values"].view(np.uint16))


def test_bake_first_two_calls_sprite_filter(golden):
    g = golden("bake_first_sprite2_two_calls")
    k, tex = int(g["k"]), int(g["tex"])
    values, writtens = O.corrmap_new(k, tex, tex, 4)
    O.corrmap_update(values, writtens, g["colors"][:2], g["ids"][:2], spriteID=2, materialID=0, mode="first")
    O.corrmap_update(values, writtens, g["colors"][2:], g["ids"][2:], spriteID=2, materialID=0, mode="first")
    assert np.array_equal(writtens, g["writtens"])
    assert np.array_equal(values.view(np.uint16), g["values"].view(np.uint16))


def test_bake_via_finished_c3(golden):
    g = golden("bake_finished_replace_c3")
    k, tex = int(g["k"]), int(g["tex"])
    values, writtens = O.corrmap_new(k, tex, tex, 3)
    masks = O.idmap_masks(g["ids"])
    O.corrmap_update(values, writtens, g["colors"], g["ids"], spriteID=1, materialID=0, mode="replace", masks=masks,
                     inverse_masks=True, ignore_obj_mat_id=True)
    assert np.array_equal(writtens, g["writtens"])
    assert np.array_equal(values.view(np.uint16), g["values"].view(np.uint16))


def test_bake_bad_index_raises():
    values, writtens = O.corrmap_new(1, 4, 4, 4)
    ids = np.zeros((1, 2, 2, 4), dtype=np.int32)
    ids[..., 2] = 2048
    with pytest.raises(IndexError):   # verified reference behaviour without masks
        O.corrmap_update(values, writtens, np.zeros((1, 2, 2, 3), np.float32), ids, mode="replace")


def _flatten(traces):
    keys = np.array(list(traces.keys()), dtype=np.int64)
    lens = np.array([len(v) for v in traces.values()], dtype=np.int64)
    flat = np.array([e for v in traces.values() for e in v], dtype=np.int64)
    return keys, lens, flat


def test_legacy_correspondence_map(golden):
    g = golden("legacy_corrmap")
    keys, lens, flat = _flatten(O.correspondence_traces(g["ids"]))
    assert np.array_equal(keys, g["keys"]) and np.array_equal(lens, g["lens"]) and np.array_equal(flat, g["traces"])
    m = golden("legacy_corrmap_merge4")
    keys, lens, flat = _flatten(O.correspondence_traces(g["ids"], merge_len=4))
    assert np.array_equal(keys, m["keys"]) and np.array_equal(lens, m["lens"]) and np.array_equal(flat, m["traces"])


@pytest.mark.parametrize("strategy", O.STRATEGIES)
@pytest.mark.parametrize("radius", [0, 1])
@pytest.mark.parametrize("cm", ["full", "merge4"])
def test_legacy_resize_overlap(golden, strategy, radius, cm):
    g = golden("legacy_resize_overlap")
    out = O.legacy_resize_overlap(g["frames"], g["ids"], float(g["alpha"]), strategy, merge_len=4 if cm == "merge4" else 0,
                                  view_normal_map=g["view_normal"], kernel_radius=radius)
    np.testing.assert_allclose(out, g[f"out_{strategy}_r{radius}_{cm}"], rtol=1e-12, atol=1e-13)


def test_legacy_scheduler_table(golden):
    table = golden("legacy_scheduler")["table"]
    names = ("constant", "linear", "exponential", "cosine")
    for itype, power, step, ts, want in table:
        got = O.scheduler_value(int(step), ts, every_step=2, start_step=2, end_step=40, start_timestep=100,
                                end_timestep=900, interpolate_begin=0.9, interpolate_end=0.2, power=power,
                                interpolate_type=names[int(itype)], no_interpolate_return=0.05)
        assert got == pytest.approx(want, rel=1e-12, abs=1e-15)


@pytest.mark.parametrize("name", STEP_CASES)
def test_torch_port_matches_reference(golden, name):
    """oracle/torch_port.py (bench.py's CPU baseline) against the same reference-generated fixtures."""
    import torch
    from torch_port import CpuOverlapPort
    g = golden(name)
    torch.set_num_threads(1)     # index_put_ with duplicate indices is only ordered with one thread
    port = CpuOverlapPort(torch.from_numpy(g["ids"]), [int(v) for v in g["frame_indices"]])
    assert port.n_entries == g["vsi"].shape[0]
    out = port.step(torch.from_numpy(g["x"]), float(g["ratio"]))
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=1e-5, atol=2e-6)


def test_torch_bake_port_matches_reference(golden):
    import torch
    from torch_port import cpu_bake_port
    torch.set_num_threads(1)
    for mode in ("first", "replace"):
        g = golden(f"bake_{mode}_masked")
        k, tex = int(g["k"]), int(g["tex"])
        values = torch.zeros(k * k, tex * tex, 4, dtype=torch.float16)
        writtens = torch.zeros(k * k, tex * tex, dtype=torch.bool)
        cpu_bake_port(values, writtens, torch.from_numpy(g["colors"]), torch.from_numpy(g["ids"]),
                      torch.from_numpy(g["masks"]), mode)
        assert np.array_equal(writtens.numpy(), g["writtens"])
        assert np.array_equal(values.numpy().view(np.uint16), g["values"].view(np.uint16))


@pytest.mark.parametrize("name", ["step_sq64_r8", "step_sq100_nonint", "step_sq60_to_16"])
def test_torch_chain_referee_matches_reference(golden, name):
    """tests/helpers.py::torch_chain — the referee of the full-size GPU tests — against reference-generated fixtures."""
    import torch
    from helpers import torch_chain
    try:
        g = golden(name)
    except FileNotFoundError:
        pytest.skip(f"no fixture {name}")
    if "frame_indices" in g.files and not np.array_equal(g["frame_indices"], np.arange(len(g["frame_indices"]))):
        pytest.skip("the referee assumes identity frame indices")
    out = torch_chain(torch.from_numpy(g["ids"].astype(np.int32)), torch.from_numpy(g["x"].astype(np.float32)), float(g["ratio"]))
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=1e-5, atol=3e-6)
