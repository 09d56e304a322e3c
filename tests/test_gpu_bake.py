"""-m gpu: CorrespondMap.update on the GPU vs reference-generated fixtures (bit-exact fp16 atlas + written flags)."""
import numpy as np
import pytest
import torch

import srx_oracle as O
from helpers import EngineData, assert_close, t2n

pytestmark = pytest.mark.gpu


def _atlas(cm):
    return cm._values.cpu().numpy().view(np.uint16), cm._writtens.cpu().numpy()


@pytest.mark.parametrize("mode", ["first", "replace", "first_avg", "replace_avg"])
def test_bake_masked_bit_exact(golden, mode):
    from stable_renderer_b200.corrmap import CorrespondMap
    g = golden(f"bake_{mode}_masked")
    k, tex = int(g["k"]), int(g["tex"])
    cm = CorrespondMap(name="t", k=k, height=tex, width=tex, channel_count=4)
    cm.update(torch.from_numpy(g["colors"]), torch.from_numpy(g["ids"]), spriteID=1, materialID=0, mode=mode,
              masks=torch.from_numpy(g["masks"]), inverse_masks=True, ignore_obj_mat_id=True)
    v, w = _atlas(cm)
    assert np.array_equal(w, g["writtens"])
    assert np.array_equal(v, g["values"].view(np.uint16))


def test_bake_first_two_calls_sprite_filter(golden):
    from stable_renderer_b200.corrmap import CorrespondMap
    g = golden("bake_first_sprite2_two_calls")
    k, tex = int(g["k"]), int(g["tex"])
    cm = CorrespondMap(name="t", k=k, height=tex, width=tex, channel_count=4)
    colors, ids = torch.from_numpy(g["colors"]).cuda(), torch.from_numpy(g["ids"]).cuda()
    cm.update(colors[:2], ids[:2], spriteID=2, materialID=0, mode="first")
    cm.update([c for c in colors[2:]], [i for i in ids[2:]], spriteID=2, materialID=0, mode="first")   # list-of-frames form
    v, w = _atlas(cm)
    assert np.array_equal(w, g["writtens"])
    assert np.array_equal(v, g["values"].view(np.uint16))


def test_bake_through_default_corresponder_finished(golden):
    from stable_renderer_b200.corresponder import DefaultCorresponder
    from stable_renderer_b200.corrmap import CorrespondMap, IDMap
    g = golden("bake_finished_replace_c3")
    k, tex = int(g["k"]), int(g["tex"])
    cm = CorrespondMap(name="t", k=k, height=tex, width=tex, channel_count=3)
    dc = DefaultCorresponder(update_corrmap_mode="replace", ignore_obj_mat_id_when_update=True)
    ed = EngineData(IDMap(tensor=torch.from_numpy(g["ids"]).cuda()), {(1, 0): cm})
    dc.finished(ed, torch.from_numpy(g["colors"]).cuda())
    v, w = _atlas(cm)
    assert np.array_equal(w, g["writtens"])
    assert np.array_equal(v, g["values"].view(np.uint16))
    dc2 = DefaultCorresponder(update_corrmap=False)
    cm2 = CorrespondMap(name="t2", k=k, height=tex, width=tex, channel_count=3)
    dc2.finished(EngineData(ed.id_maps, {(1, 0): cm2}), torch.from_numpy(g["colors"]).cuda())
    assert not cm2._writtens.any()


@pytest.mark.parametrize("mode", ["first", "replace"])
def test_bake_vs_oracle_bigger(mode):
    """256^2 views into a k=3 128^2 atlas, fp16 and bf16 colours, sprite filter + masks together (the consistent
    definition)."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.corrmap import CorrespondMap
    F, H, tex, k = 6, 256, 128, 3
    ids = synthetic.make_ids(F, H, H, tex_h=tex, tex_w=tex, k=k, n_obj=2, frac_2048=0.05, seed=3)
    colors = synthetic.make_colors(F, H, H, 3, seed=4)
    masks = torch.from_numpy(O.idmap_masks(ids.numpy()))
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        col = colors.to(dt)
        values, writtens = O.corrmap_new(k, tex, tex, 4)
        O.corrmap_update(values, writtens, col.float().numpy(), ids.numpy(), spriteID=2, materialID=0, mode=mode,
                         masks=masks.numpy(), inverse_masks=True)
        cm = CorrespondMap(name="t", k=k, height=tex, width=tex, channel_count=4)
        cm.update(col, ids, spriteID=2, materialID=0, mode=mode, masks=masks, inverse_masks=True)
        v, w = _atlas(cm)
        assert np.array_equal(w, writtens)
        assert np.array_equal(v, values.view(np.uint16))


def test_bake_index_error_and_bad_args():
    from stable_renderer_b200.corrmap import CorrespondMap
    cm = CorrespondMap(name="t", k=1, height=4, width=4, channel_count=4)
    ids = torch.zeros(1, 2, 2, 4, dtype=torch.int32)
    ids[..., 2] = 2048
    with pytest.raises(IndexError):      # verified reference behaviour without masks
        cm.update(torch.zeros(1, 2, 2, 3), ids, mode="replace")
    with pytest.raises(ValueError):
        cm.update(torch.zeros(2, 2, 2, 3), torch.zeros(1, 2, 2, 4, dtype=torch.int32))
    with pytest.raises(ValueError):
        cm.update(torch.zeros(1, 2, 2, 3), torch.zeros(1, 2, 2, 4, dtype=torch.int32), mode="bogus")


@pytest.mark.parametrize("weight_mode", ["uniform", "view_normal", "view_normal_depth"])
def test_weighted_multi_view_bake_vs_oracle(weight_mode):
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.corrmap import CorrespondMap
    F, H, tex, k = 8, 128, 64, 2
    ids = synthetic.make_ids(F, H, H, tex_h=tex, tex_w=tex, k=k, frac_2048=0.05, seed=8)
    colors = synthetic.make_colors(F, H, H, 3, seed=5)
    nd = synthetic.make_normal_depth(F, H, H, seed=2)
    acc = np.zeros((k * k, tex * tex, 4)); wsum = np.zeros((k * k, tex * tex))
    O.corrmap_update_weighted(acc, wsum, colors.numpy(), ids.numpy(), nd.float().numpy(), weight_mode, spriteID=1)
    values, writtens = O.corrmap_new(k, tex, tex, 4)
    O.corrmap_finalize_weighted(values, writtens, acc, wsum)
    cm = CorrespondMap(name="t", k=k, height=tex, width=tex, channel_count=4)
    cm.update(colors, ids, spriteID=1, mode="replace", weight_mode=weight_mode, normal_depth=nd)
    assert np.array_equal(cm._writtens.cpu().numpy(), writtens)
    # float32 atomics vs float64 sums, then one fp16 rounding: allow one fp16 ulp (2^-11 relative)
    assert_close(t2n(cm._values), values.astype(np.float32), 1e-3, 1e-6, weight_mode)


def test_weighted_bake_sharded_phases_equal_single_bake():
    """View-sharded multi-GPU bake emulated on one GPU: two maps accumulate half of the views each (phase 1), the weighted
    sums are added (what the all-reduce does), both finalise (phase 2) -> the atlas of one bake over all views."""
    import torch
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.corrmap import CorrespondMap
    F, H, tex = 6, 128, 64
    ids = synthetic.make_ids(F, H, H, tex_h=tex, tex_w=tex, k=1, seed=3).cuda()
    colors = synthetic.make_colors(F, H, H, 3, seed=4).cuda()
    nd = synthetic.make_normal_depth(F, H, H).cuda()
    kw = dict(mode="replace", weight_mode="view_normal_depth")
    ref = CorrespondMap(name="ref", k=1, height=tex, width=tex, channel_count=4)
    ref.update(colors, ids, normal_depth=nd, **kw)
    parts = []
    for r in range(2):
        sl = slice(r * F // 2, (r + 1) * F // 2)
        cm = CorrespondMap(name=f"p{r}", k=1, height=tex, width=tex, channel_count=4)
        cm.update(colors[sl], ids[sl], normal_depth=nd[sl], phase=1, **kw)
        parts.append(cm)
    n = parts[0]._workspace.numel() - 256
    total = parts[0]._workspace[:n].view(torch.float32) + parts[1]._workspace[:n].view(torch.float32)
    for r, cm in enumerate(parts):
        sl = slice(r * F // 2, (r + 1) * F // 2)
        cm._workspace[:n].view(torch.float32).copy_(total)
        cm.update(colors[sl], ids[sl], normal_depth=nd[sl], phase=2, **kw)
        assert torch.equal(cm._writtens, ref._writtens)
        assert torch.allclose(cm._values.float(), ref._values.float(), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("mode", ["first", "replace"])
def test_bake_small_update_into_large_atlas(mode):
    """Few pixels into a large atlas: the pixel-major write kernel (the texel-major one serves views with at least half as
    many pixels as the atlas has texels); two calls, so `first` must keep what the first call wrote."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.corrmap import CorrespondMap
    H, tex, k = 64, 128, 3
    ids = synthetic.make_ids(2, H, H, tex_h=tex, tex_w=tex, k=k, frac_2048=0.0, seed=8)
    colors = synthetic.make_colors(2, H, H, 3, seed=9)
    masks = torch.from_numpy(O.idmap_masks(ids.numpy()))
    values, writtens = O.corrmap_new(k, tex, tex, 4)
    cm = CorrespondMap(name="t", k=k, height=tex, width=tex, channel_count=4)
    for f in range(2):
        O.corrmap_update(values, writtens, colors[f:f + 1].numpy(), ids[f:f + 1].numpy(), mode=mode, masks=masks[f:f + 1].numpy(),
                         inverse_masks=True, ignore_obj_mat_id=True)
        cm.update(colors[f:f + 1], ids[f:f + 1], mode=mode, masks=masks[f:f + 1], inverse_masks=True, ignore_obj_mat_id=True)
    v, w = _atlas(cm)
    assert np.array_equal(w, writtens)
    assert np.array_equal(v, values.view(np.uint16))


@pytest.mark.parametrize("shape", [(3, 9, 7), (2, 16, 24)])
def test_bake_odd_and_even_pixel_counts_all_paths(shape):
    """Odd pixel counts take the one-pixel-per-thread kernels, even ones the pair kernels; reference modes bit exact, weighted
    bake within one fp16 ulp, with RGB colours (weight sums in the alpha channel), RGBA colours and fp16 colours (scalar)."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.corrmap import CorrespondMap
    F, H, W = shape
    tex, k = 16, 2
    ids = synthetic.make_ids(F, H, W, tex_h=tex, tex_w=tex, k=k, frac_2048=0.05, seed=21)
    masks = torch.from_numpy(O.idmap_masks(ids.numpy()))
    nd = synthetic.make_normal_depth(F, H, W, seed=6)
    for cin, cdt in ((3, torch.float32), (4, torch.float32), (3, torch.float16)):
        colors = synthetic.make_colors(F, H, W, cin, seed=5).to(cdt)
        for mode in ("replace", "first"):
            values, writtens = O.corrmap_new(k, tex, tex, 4)
            O.corrmap_update(values, writtens, colors.float().numpy(), ids.numpy(), mode=mode, masks=masks.numpy(),
                             inverse_masks=True, ignore_obj_mat_id=True)
            cm = CorrespondMap(name="t", k=k, height=tex, width=tex, channel_count=4)
            cm.update(colors, ids, mode=mode, masks=masks, inverse_masks=True, ignore_obj_mat_id=True)
            v, w = _atlas(cm)
            assert np.array_equal(w, writtens), (cin, cdt, mode)
            assert np.array_equal(v, values.view(np.uint16)), (cin, cdt, mode)
        acc = np.zeros((k * k, tex * tex, 4)); wsum = np.zeros((k * k, tex * tex))
        O.corrmap_update_weighted(acc, wsum, colors.float().numpy(), ids.numpy(), nd.float().numpy(), "view_normal_depth")
        values, writtens = O.corrmap_new(k, tex, tex, 4)
        O.corrmap_finalize_weighted(values, writtens, acc, wsum)
        cm = CorrespondMap(name="t", k=k, height=tex, width=tex, channel_count=4)
        cm.update(colors, ids, mode="replace", weight_mode="view_normal_depth", normal_depth=nd)
        assert np.array_equal(cm._writtens.cpu().numpy(), writtens), (cin, cdt)
        assert_close(t2n(cm._values), values.astype(np.float32), 1e-3, 1e-6, f"weighted {cin} {cdt}")


@pytest.mark.parametrize("shape", [(3, 9, 7), (4, 64, 64)])
@pytest.mark.parametrize("mode", ["replace", "first"])
def test_bake_without_masks_background_goes_to_texel_zero(shape, mode):
    """No masks, no sprite filter: pixels without an id carry (0,0,0,0) and all write texel 0 of map 0, like in the reference
    (corrmap.py:735 with nothing filtered).  Those claims are aggregated per block in the kernel; the winner must still be the
    reference's (last pixel of the last / earliest frame)."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.corrmap import CorrespondMap
    F, H, W = shape
    tex, k = 16, 2
    ids = synthetic.make_ids(F, H, W, tex_h=tex, tex_w=tex, k=k, frac_2048=0.0, seed=33)
    assert (ids == 0).all(dim=-1).any()
    colors = synthetic.make_colors(F, H, W, 3, seed=12)
    values, writtens = O.corrmap_new(k, tex, tex, 4)
    cm = CorrespondMap(name="t", k=k, height=tex, width=tex, channel_count=4)
    for sl in (slice(0, F - 1), slice(F - 1, F)):                      # two calls: `first` keeps texel 0 from the first call
        O.corrmap_update(values, writtens, colors[sl].numpy(), ids[sl].numpy(), mode=mode, ignore_obj_mat_id=True)
        cm.update(colors[sl], ids[sl], mode=mode, ignore_obj_mat_id=True)
    v, w = _atlas(cm)
    assert w[0, 0] and np.array_equal(w, writtens)
    assert np.array_equal(v, values.view(np.uint16))


@pytest.mark.parametrize("mode", ["replace", "first"])
def test_reference_mode_bake_sharded_phases_equal_single_bake(mode):
    """View-sharded bake of the reference modes emulated on one GPU: three 'ranks' hold 2 + 3 + 1 views; claim
    with order keys over all views (phase 1), MAX of the owner words (the all-reduce), each rank writes the texels its views won
    (phase 2), SUM of the partial atlases as int32 words, merge (phase 3) -> bit for bit the atlas of one bake over all views.
    Run twice, so that `first` has to respect texels written by the first round."""
    import torch
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.corrmap import CorrespondMap
    F, H, tex, k = 6, 64, 48, 2
    cuts = [(0, 2), (2, 5), (5, 6)]
    ntex = k * k * tex * tex
    ref = CorrespondMap(name="ref", k=k, height=tex, width=tex, channel_count=4)
    parts = [CorrespondMap(name=f"p{r}", k=k, height=tex, width=tex, channel_count=4) for r in range(len(cuts))]
    for rnd in range(2):
        ids = synthetic.make_ids(F, H, H, tex_h=tex, tex_w=tex, k=k, frac_2048=0.0, seed=40 + rnd, frame_offset=3 * rnd).cuda()
        colors = synthetic.make_colors(F, H, H, 3, seed=41 + rnd).cuda()
        kw = dict(mode=mode, ignore_obj_mat_id=True)
        ref.update(colors, ids, **kw)
        sh = dict(frames_global=F, **kw)
        for (lo, hi), cm in zip(cuts, parts):
            cm.update(colors[lo:hi], ids[lo:hi], phase=1, frame_offset=lo, **sh)
        owner = torch.stack([cm._workspace[:ntex * 4].view(torch.int32) for cm in parts]).amax(dim=0)
        for (lo, hi), cm in zip(cuts, parts):
            cm._workspace[:ntex * 4].view(torch.int32).copy_(owner)
            cm.update(colors[lo:hi], ids[lo:hi], phase=2, frame_offset=lo, **sh)
        a0 = (ntex * 4 + 255) // 256 * 256 + 256
        n = parts[0]._workspace.numel()
        delta = torch.stack([cm._workspace[a0:n].view(torch.int32) for cm in parts]).sum(dim=0, dtype=torch.int32)
        for (lo, hi), cm in zip(cuts, parts):
            cm._workspace[a0:n].view(torch.int32).copy_(delta)
            cm.update(colors[lo:hi], ids[lo:hi], phase=3, frame_offset=lo, **sh)
            assert torch.equal(cm._writtens, ref._writtens), (mode, rnd)
            assert torch.equal(cm._values.view(torch.int16), ref._values.view(torch.int16)), (mode, rnd)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_view_sharded_bake_two_gpus():
    """Real view-sharded bakes over NCCL: reference modes bit exact, weighted bake within an fp16 ulp (tests/mp_bake_worker.py)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(root, "tests", "mp_bake_worker.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    assert "BAKE_SHARD_OK" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-2000:]


def test_bake_deferred_status_check():
    """`update(..., defer_check=True)` never syncs the host; an out-of-atlas texel is reported by `check()`, once."""
    from stable_renderer_b200.corrmap import CorrespondMap
    H = 16
    ids = torch.zeros(1, H, H, 4, dtype=torch.int32)
    ids[..., 0], ids[..., 3] = 1, torch.arange(H * H, dtype=torch.int32).view(H, H)
    colors = torch.rand(1, H, H, 3)
    cm = CorrespondMap(name="d", k=1, height=H, width=H, channel_count=4)
    cm.update(colors.cuda(), ids.cuda(), mode="replace", ignore_obj_mat_id=True, defer_check=True)
    cm.check()                                                     # all texels inside: nothing to report
    ref = CorrespondMap(name="r", k=1, height=H, width=H, channel_count=4)
    ref.update(colors.cuda(), ids.cuda(), mode="replace", ignore_obj_mat_id=True)
    assert torch.equal(cm._values, ref._values) and torch.equal(cm._writtens, ref._writtens)
    bad = ids.clone()
    bad[0, 3, 3, 3] = H * H + 5                                    # vertex id outside the atlas
    cm.update(colors.cuda(), bad.cuda(), mode="replace", ignore_obj_mat_id=True, defer_check=True)
    cm.update(colors.cuda(), ids.cuda(), mode="replace", ignore_obj_mat_id=True, defer_check=True)    # the flag survives later calls
    with pytest.raises(IndexError):
        cm.check()
    cm.check()                                                     # reported once, then cleared
    with pytest.raises(IndexError):                                # the default stays the reference's behaviour
        cm.update(colors.cuda(), bad.cuda(), mode="replace", ignore_obj_mat_id=True)
