"""-m gpu: legacy overlap (ResizeOverlap / Overlap, four strategies) and the texture<->tensor copy kernels."""
import ctypes as C

import numpy as np
import pytest
import torch

import srx_oracle as O
from helpers import assert_close, t2n

pytestmark = pytest.mark.gpu


def _schedulers(alpha, radius=0.0):
    from stable_renderer_b200.overlap import Scheduler
    return (Scheduler(interpolate_begin=alpha, interpolate_end=alpha, interpolate_type="constant"),
            Scheduler(interpolate_begin=radius, interpolate_end=radius, interpolate_type="constant"))


@pytest.mark.parametrize("strategy", O.STRATEGIES)
@pytest.mark.parametrize("cm_name", ["full", "merge4"])
def test_resize_overlap_vs_reference_golden(golden, strategy, cm_name):
    from stable_renderer_b200.overlap import CorrespondenceMap, ResizeOverlap, overlap_algorithm_factory
    g = golden("legacy_resize_overlap")
    cmap = CorrespondenceMap.from_ids(torch.from_numpy(g["ids"]).cuda())
    if cm_name == "merge4":
        cmap.merge_nearby(4)
    a_s, r_s = _schedulers(float(g["alpha"]))
    ov = ResizeOverlap(a_s, r_s, overlap_algorithm_factory(strategy), verbose=False)
    frames = [torch.from_numpy(f).float().cuda() for f in g["frames"]]
    outs = ov(frames, cmap, step=0, timestep=500, view_normal_map=torch.from_numpy(g["view_normal"]).float().cuda())
    assert isinstance(outs, list) and len(outs) == len(frames) and outs[0].shape == frames[0].shape
    got = torch.stack(outs)
    assert_close(t2n(got), g[f"out_{strategy}_r0_{cm_name}"], 1e-5, 3e-6, f"{strategy}/{cm_name}")


def test_resize_overlap_alpha_zero_and_unknown_algorithm():
    from stable_renderer_b200.overlap import CorrespondenceMap, ResizeOverlap, overlap_algorithm_factory
    ids = torch.ones(2, 16, 16, 4, dtype=torch.int32).cuda()
    cmap = CorrespondenceMap.from_ids(ids)
    frames = [torch.randn(1, 4, 2, 2).cuda() for _ in range(2)]
    a_s, r_s = _schedulers(0.0)
    out = ResizeOverlap(a_s, r_s, overlap_algorithm_factory("average"), verbose=False)(frames, cmap, step=0, timestep=500)
    assert out is frames                                     # alpha == 0 returns the input list (overlap.py:200-201)
    with pytest.raises(ValueError):
        overlap_algorithm_factory("bogus")


@pytest.mark.parametrize("strategy", O.STRATEGIES)
@pytest.mark.parametrize("cm_name", ["full", "merge4"])
def test_resize_overlap_radius1_vs_reference_golden(golden, strategy, cm_name):
    """kernel_radius = 1: the reference's in-place sweep in dict order (overlap.py:97,136-145), reproduced trace after trace."""
    from stable_renderer_b200.overlap import CorrespondenceMap, ResizeOverlap, overlap_algorithm_factory
    g = golden("legacy_resize_overlap")
    cmap = CorrespondenceMap.from_ids(torch.from_numpy(g["ids"]).cuda())
    if cm_name == "merge4":
        cmap.merge_nearby(4)
    a_s, r_s = _schedulers(float(g["alpha"]), 1.0)
    ov = ResizeOverlap(a_s, r_s, overlap_algorithm_factory(strategy), verbose=False)
    frames = [torch.from_numpy(f).float().cuda() for f in g["frames"]]
    outs = ov(frames, cmap, step=0, timestep=500, view_normal_map=torch.from_numpy(g["view_normal"]).float().cuda())
    assert_close(t2n(torch.stack(outs)), g[f"out_{strategy}_r1_{cm_name}"], 2e-5, 5e-6, f"{strategy}/{cm_name} radius 1")


@pytest.mark.parametrize("strategy", ["average", "pixel_distance"])
@pytest.mark.parametrize("radius", [0, 2])
def test_overlap_full_resolution_with_radius_vs_oracle(strategy, radius):
    """Overlap.__call__ at map resolution through the ordered sweep: radius 2, and radius 0 (must equal the unordered kernels)."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.overlap import CorrespondenceMap, Overlap, overlap_algorithm_factory
    T, H = 4, 40
    ids = synthetic.make_ids(T, H, H, tex_h=20, tex_w=20, n_obj=2, seed=19)
    gen = torch.Generator().manual_seed(2)
    frames = [torch.randn(2, 4, H, H, generator=gen) for _ in range(T)]
    want = O.legacy_overlap(torch.stack(frames).numpy(), ids.numpy(), 0.6, strategy, kernel_radius=radius)
    a_s, r_s = _schedulers(0.6, float(radius))
    ov = Overlap(a_s, r_s, overlap_algorithm_factory(strategy), verbose=False)
    cmap = CorrespondenceMap.from_ids(ids.cuda())
    if radius == 0:                                   # force the ordered entry for radius 0 as well
        stack = torch.stack([f.cuda() for f in frames]).contiguous()
        ov._run(stack.view(T, 8, H, H), cmap, 0.6, None, radius=0)
        got0 = stack.clone()
        from stable_renderer_b200 import _lib
        import ctypes as C
        lib = _lib.load()
        d = _lib.srx_legacy_desc()
        d.id_dtype, d.frames, d.height, d.width = _lib.torch_dtype_code(torch.int32), T, H, H
        d.channels, d.lat_h, d.lat_w, d.merge_len, d.strategy = 8, H, H, 0, _lib.SRX_STRATEGY[strategy]
        ws = torch.empty(int(lib.srx_legacy_ordered_workspace_bytes(C.byref(d))), dtype=torch.uint8, device="cuda")
        stack2 = torch.stack([f.cuda() for f in frames]).contiguous()
        a = _lib.srx_legacy_args()
        a.x_dev, a.x_dtype, a.ids_dev, a.alpha = stack2.data_ptr(), _lib.SRX_F32, cmap.device_ids(stack2.device).data_ptr(), 0.6
        a.workspace_dev, a.workspace_bytes = ws.data_ptr(), ws.numel()
        _lib.check(lib.srx_legacy_overlap_ordered(C.byref(d), C.byref(a), 0, _lib.current_stream_ptr(stack2.device)))
        torch.cuda.synchronize()
        assert_close(t2n(stack2), t2n(got0), 1e-5, 3e-6, "ordered radius 0 vs unordered kernels")
        got = stack2
    else:
        got = ov([f.cuda() for f in frames], cmap, step=0, timestep=500)
    assert_close(t2n(got), want, 2e-5, 5e-6, f"{strategy} radius {radius}")


@pytest.mark.parametrize("strategy", O.STRATEGIES)
def test_overlap_full_resolution_vs_oracle(strategy):
    """Overlap.__call__ at correspondence-map resolution (no resize), int32 current-generation ids."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.overlap import CorrespondenceMap, Overlap, overlap_algorithm_factory
    T, H = 5, 48
    ids = synthetic.make_ids(T, H, H, tex_h=24, tex_w=24, n_obj=2, seed=9)
    gen = torch.Generator().manual_seed(1)
    frames = [torch.randn(2, 4, H, H, generator=gen) for _ in range(T)]
    vn = torch.rand(T, H, H, 1, generator=gen)
    want = O.legacy_overlap(torch.stack(frames).numpy(), ids.numpy(), 0.6, strategy, view_normal_map=vn.numpy())
    a_s, r_s = _schedulers(0.6)
    ov = Overlap(a_s, r_s, overlap_algorithm_factory(strategy), verbose=False)
    got = ov([f.cuda() for f in frames], CorrespondenceMap.from_ids(ids.cuda()), step=0, timestep=500,
             view_normal_map=vn.cuda())
    assert tuple(got.shape) == (T, 2, 4, H, H)
    assert_close(t2n(got), want, 1e-5, 3e-6, strategy)


def test_resize_overlap_cfg1_like_vs_oracle_average():
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.overlap import CorrespondenceMap, ResizeOverlap, overlap_algorithm_factory
    T, H, h = 6, 256, 32
    ids = synthetic.make_ids(T, H, H, tex_h=128, tex_w=128, seed=10, legacy_layout=True, dtype=torch.int16)
    frames = [synthetic.make_latents(1, 4, h, h, seed=20 + i) for i in range(T)]
    want = O.legacy_resize_overlap(torch.stack(frames).numpy(), ids.numpy(), 0.9, "average", merge_len=4)
    cmap = CorrespondenceMap.from_ids(ids.cuda())
    cmap.merge_nearby(4)
    a_s, r_s = _schedulers(0.9)
    outs = ResizeOverlap(a_s, r_s, overlap_algorithm_factory("average"), verbose=False)(
        [f.cuda() for f in frames], cmap, step=0, timestep=500)
    assert_close(t2n(torch.stack(outs)), want, 1e-5, 3e-6)


@pytest.mark.parametrize("channels,bits,kind,dtype", [(4, 32, 0, torch.int32), (4, 16, 2, torch.float16),
                                                       (1, 32, 2, torch.float32), (2, 16, 1, torch.int16), (4, 8, 1, torch.uint8)])
def test_array_tensor_round_trip_with_flip(channels, bits, kind, dtype):
    """The copy kernels behind Texture.tensor()/set_data() (texture.py:221-254, 326-408) on a plain cudaArray:
    tensor -> array -> tensor is the identity, and flip=1 reverses the rows (GL bottom-left origin)."""
    from stable_renderer_b200 import _lib
    lib = _lib.load()
    W, H = 70, 37
    arr = C.c_void_p()
    _lib.check(lib.srx_array_alloc(C.byref(arr), W, H, channels, bits, kind))
    try:
        if dtype.is_floating_point:
            src = torch.randn(H, W, channels, device="cuda").to(dtype)
        else:
            src = torch.randint(0, 100, (H, W, channels), device="cuda").to(dtype)
        texel = channels * bits // 8
        stream = _lib.current_stream_ptr()
        _lib.check(lib.srx_tensor_to_array(arr, src.data_ptr(), W, H, texel, 1, 0, 0, stream))      # set_data(flip)
        same = torch.empty_like(src)
        _lib.check(lib.srx_array_to_tensor(arr, same.data_ptr(), W, H, texel, 1, stream))           # tensor(flip=True)
        raw = torch.empty_like(src)
        _lib.check(lib.srx_array_to_tensor(arr, raw.data_ptr(), W, H, texel, 0, stream))
        torch.cuda.synchronize()
        assert torch.equal(same.view(torch.uint8), src.view(torch.uint8))
        assert torch.equal(raw.view(torch.uint8), src.flip(0).contiguous().view(torch.uint8))
        with pytest.raises(ValueError):
            _lib.check(lib.srx_array_to_tensor(arr, raw.data_ptr(), W, H, texel * 2 if texel < 16 else 8, 0, stream))
    finally:
        lib.srx_array_free(arr)


@pytest.mark.parametrize("strategy", ["average", "frame_distance"])
def test_level_scheduled_sweep_equals_the_sequential_sweep(strategy, monkeypatch):
    """kernel_radius = 2 on a 16-frame map: the level schedule (traces of one conflict level in parallel) must reproduce the
    one-after-another sweep bit for bit — same arithmetic per trace, same values read."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.overlap import CorrespondenceMap, ResizeOverlap, overlap_algorithm_factory
    T, H, h = 16, 128, 16
    ids = synthetic.make_ids(T, H, H, tex_h=48, tex_w=48, seed=83, legacy_layout=True).cuda()
    cmap = CorrespondenceMap.from_ids(ids)
    gen = torch.Generator().manual_seed(4)
    frames = [torch.randn(1, 4, h, h, generator=gen).cuda() for _ in range(T)]
    a_s, r_s = _schedulers(0.8, 2.0)
    ov = ResizeOverlap(a_s, r_s, overlap_algorithm_factory(strategy), verbose=False)
    monkeypatch.delenv("SRX_ORD_SEQUENTIAL", raising=False)
    par = torch.stack(ov([f.clone() for f in frames], cmap, step=0, timestep=500))
    monkeypatch.setenv("SRX_ORD_SEQUENTIAL", "1")
    seq = torch.stack(ov([f.clone() for f in frames], cmap, step=0, timestep=500))
    assert torch.equal(par, seq)
    assert not torch.equal(par, torch.stack(frames))


@pytest.mark.parametrize("mode", ["bilinear", "bicubic", "area"])
@pytest.mark.parametrize("strategy", ["average", "frame_distance"])
def test_resize_overlap_smooth_interpolation_vs_reference_golden(golden, mode, strategy):
    """interpolate_mode != 'nearest' (overlap.py:205-221): up-sample, overlap at map resolution, down-sample, where()."""
    from stable_renderer_b200.overlap import CorrespondenceMap, ResizeOverlap, overlap_algorithm_factory
    g = golden("legacy_resize_overlap_interp")
    cmap = CorrespondenceMap.from_ids(torch.from_numpy(g["ids"]).cuda())
    a_s, r_s = _schedulers(float(g["alpha"]))
    ov = ResizeOverlap(a_s, r_s, overlap_algorithm_factory(strategy), verbose=False, interpolate_mode=mode)
    frames = [torch.from_numpy(f).cuda() for f in g["frames"]]
    outs = ov(frames, cmap, step=0, timestep=500)
    assert isinstance(outs, list) and outs[0].shape == frames[0].shape
    assert_close(t2n(torch.stack(outs)), g[f"out_{mode}_{strategy}"], 2e-5, 5e-6, f"{mode}/{strategy}")
