"""L7: johnny_overlap.overlap (legacy_diffuser/modules/diffuser_pipelines/overlap/johnny_overlap.py:15-141).  Fixture produced
by the reference function itself with its one broken statement replaced (oracle/ref_shim.py::johnny_overlap_function)."""
import numpy as np
import pytest
import torch

import srx_oracle as O
from helpers import assert_close, t2n


class _Sched:
    @staticmethod
    def add_noise(lat, nz, t):
        return lat * 0.8 + nz * 0.6


class _Pipe:
    scheduler = _Sched()


def test_oracle_johnny_overlap_matches_reference(golden):
    g = golden("johnny_overlap")
    assert bool(g["gated_is_input"])
    assert_close(O.johnny_overlap(g["frames"], g["ids"], alpha=1.0), g["out_beta0"], 1e-12, 1e-12, "beta 0")
    base = g["orig"] * 0.8 + g["noise"] * 0.6
    assert_close(O.johnny_overlap(g["frames"], g["ids"], alpha=1.0, beta=0.3, base=base), g["out_beta03"], 1e-12, 1e-12, "beta 0.3")


def test_johnny_schedule_is_the_reference_table():
    from stable_renderer_b200.overlap.johnny import schedule
    assert schedule(3, 500, "constant", 1) == 1 and schedule(3, 1500, "constant", 1) == 0 and schedule(3, -1, "constant", 1) == 0
    assert schedule(3, 250, "linear", 0.2, 1.0) == pytest.approx(0.2 + 0.8 * 0.75)
    assert schedule(4, 500, "constant", 0.7, every_step=3) == 0
    with pytest.raises(TypeError):
        schedule(3, 500, "bogus", 1)


@pytest.mark.gpu
def test_gpu_johnny_overlap_vs_reference(golden):
    from stable_renderer_b200.overlap import CorrespondenceMap
    from stable_renderer_b200.overlap.johnny import overlap
    g = golden("johnny_overlap")
    cmap = CorrespondenceMap.from_ids(torch.from_numpy(g["ids"]).cuda())
    frames = [torch.from_numpy(f).float().cuda() for f in g["frames"]]
    out = overlap([f.clone() for f in frames], cmap, _Pipe(), step=3, timestep=500)
    assert isinstance(out, list) and len(out) == len(frames) and out[0].shape == frames[0].shape
    assert_close(t2n(torch.stack(out)), g["out_beta0"], 2e-5, 5e-6, "beta 0")
    orig = [torch.from_numpy(f).float().cuda() for f in g["orig"]]
    noise = [torch.from_numpy(f).float().cuda() for f in g["noise"]]
    out = overlap([f.clone() for f in frames], cmap, _Pipe(), step=3, timestep=500, init_latents_orig_seq=orig, noise_seq=noise, beta=0.3)
    assert_close(t2n(torch.stack(out)), g["out_beta03"], 2e-5, 5e-6, "beta 0.3")
    gated = overlap(frames, cmap, _Pipe(), step=3, timestep=1500)          # alpha schedule returns 0: the input list comes back
    assert gated is frames
    with pytest.raises(NotImplementedError):
        overlap(frames, cmap, _Pipe(), interpolate_mode="bilinear", step=3, timestep=500)


@pytest.mark.gpu
def test_gpu_johnny_overlap_vs_oracle_longer_traces():
    """16 frames at 64x64 on a 24x24 texture: traces of ~100 entries (several per frame), half-precision latents."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.overlap import CorrespondenceMap
    from stable_renderer_b200.overlap.johnny import overlap
    T, H, h = 16, 64, 8
    ids = synthetic.make_ids(T, H, H, tex_h=24, tex_w=24, seed=73, legacy_layout=True)
    gen = torch.Generator().manual_seed(3)
    frames = [torch.randn(1, 4, h, h, generator=gen) for _ in range(T)]
    want = O.johnny_overlap(torch.stack(frames).numpy(), ids.numpy(), alpha=1.0)
    cmap = CorrespondenceMap.from_ids(ids.cuda())
    out = overlap([f.cuda() for f in frames], cmap, None, step=0, timestep=900)
    assert_close(t2n(torch.stack(out)), want, 5e-5, 1e-5, "f32")
    out16 = overlap([f.half().cuda() for f in frames], cmap, None, step=0, timestep=900)
    assert out16[0].dtype == torch.float16
    assert_close(t2n(torch.stack(out16)), O.johnny_overlap(torch.stack(frames).half().float().numpy(), ids.numpy(), alpha=1.0),
                 1e-2, 1e-2, "f16")


@pytest.mark.gpu
def test_gpu_johnny_overlap_rejects_merged_maps():
    """Inside a trace the update is sequential, so the entry order matters; a merged map concatenates its sub-traces in dict
    order (correspondence_map.py:276-286), which the (frame,row,col) sort of the GPU path does not reproduce."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.overlap import CorrespondenceMap
    from stable_renderer_b200.overlap.johnny import overlap
    ids = synthetic.make_ids(2, 32, 32, tex_h=16, tex_w=16, seed=1, legacy_layout=True).cuda()
    cmap = CorrespondenceMap.from_ids(ids)
    cmap.merge_nearby(2)
    with pytest.raises(NotImplementedError):
        overlap([torch.randn(1, 4, 4, 4).cuda() for _ in range(2)], cmap, None, step=0, timestep=900)
