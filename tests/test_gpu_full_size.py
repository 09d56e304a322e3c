"""-m gpu parity at BASELINE.json's full sizes (configs 3 and 5: 96 x 1024^2 ids / 128x128x4 latents, 768 x 512^2 ids / 64x64x4
latents), where the numpy oracle would take minutes: size-independent properties of the overlap step, and an independent
restatement of the reference's op chain (corresponder.py:298-376) in torch CUDA ops with float64 accumulation and an explicit
last-writer per cell (the order the reference's one-thread `index_put_` has) — `helpers.torch_chain`, itself pinned to the
reference-generated fixtures at small size by tests/test_oracle_vs_golden.py::test_torch_chain_referee_matches_reference."""
import pytest
import torch

from helpers import assert_close, t2n, torch_chain as _torch_chain

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 3e-6
CONFIGS = {"cfg3": (96, 1024, 128, 1024), "cfg5": (768, 512, 64, 512)}     # frames, id size, latent size, texture size


def _make(cfg):
    from stable_renderer_b200 import synthetic
    F, H, h, tex = CONFIGS[cfg]
    ids = synthetic.make_ids(F, H, H, tex_h=tex, tex_w=tex, frac_2048=0.05, seed=1234 + F, device="cuda")
    x0 = synthetic.make_latents(F, 4, h, h, seed=0, device="cuda")
    return ids, x0, tex * tex


@pytest.mark.parametrize("cfg", ["cfg3", "cfg5"])
def test_full_size_step_properties(cfg):
    from stable_renderer_b200.plan import OverlapPlan
    ids, x0, cap = _make(cfg)
    F = ids.shape[0]
    plan = OverlapPlan(ids, x0.shape, key_capacity=cap)
    assert plan.fused

    # (1) against the torch restatement of the reference's op chain at full size
    x = x0.clone()
    plan.step(x, 0.5)
    plan.check()
    want = _torch_chain(ids, x0, 0.5)
    assert_close(t2n(x), t2n(want), RTOL, ATOL, cfg + " vs torch chain")

    # (2) inject ratio 0: the blended tensor is x itself, so AdaIN returns x
    y = x0.clone()
    plan.step(y, 0.0)
    assert_close(t2n(y), t2n(x0), RTOL, ATOL, cfg + " ratio 0")

    # (3) the split kernels (another implementation of the same step) agree at full size
    z = x0.clone()
    split = OverlapPlan(ids, x0.shape, key_capacity=cap, split_kernels=True)
    assert not split.fused
    split.step(z, 0.5)
    split.check()
    assert_close(t2n(z), t2n(x), RTOL, ATOL, cfg + " split vs fused")

    # (4) cached-plan regime from the bucketing pass
    c = x0.clone()
    plan.build_cache()
    plan.step(c, 0.5, cached=True)
    plan.check()
    assert_close(t2n(c), t2n(x), RTOL, ATOL, cfg + " cached vs streaming")

    # (5) frame-permutation equivariance: frames are independent except through the key means, which are sums over all frames
    perm = torch.randperm(F, generator=torch.Generator().manual_seed(3)).cuda()
    p = x0[perm].contiguous()
    OverlapPlan(ids[perm].contiguous(), x0.shape, key_capacity=cap).step(p, 0.5)
    assert_close(t2n(p), t2n(x[perm]), RTOL, ATOL, cfg + " frame permutation")

    # (6) scaling the latents by 2 scales the result by 2 (exact in floating point except for the +1e-5 under the root)
    s = (x0 * 2).contiguous()
    plan.step(s, 0.5)
    assert_close(t2n(s), t2n(x) * 2, 2e-5, 1e-5, cfg + " scaling")

    # (7) without AdaIN the step is linear in the latents: (1 - r) x + r * mean_key(x)
    x1 = x0.clone()
    x2 = torch.roll(x0, 7, 0).contiguous() * 0.5
    x12 = (x1 + x2).contiguous()
    plan.step(x1, 0.5, adain=False)
    plan.step(x2, 0.5, adain=False)
    plan.step(x12, 0.5, adain=False)
    plan.check()
    assert_close(t2n(x12), t2n(x1 + x2), 2e-5, 1e-5, cfg + " linearity")
    assert (x1 - x0).abs().max() > 0.1                                   # the step did something


def test_full_size_step_bf16_cfg3():
    """config 3's dtype: bf16 latents, tolerance 1e-2 (north_star)."""
    from stable_renderer_b200.plan import OverlapPlan
    ids, x0, cap = _make("cfg3")
    xb = x0.to(torch.bfloat16)
    want = _torch_chain(ids, xb.float(), 0.5)
    plan = OverlapPlan(ids, xb.shape, key_capacity=cap)
    plan.step(xb, 0.5)
    plan.check()
    assert_close(t2n(xb), t2n(want), 1e-2, 1e-2, "cfg3 bf16")


def _torch_bake(colors, ids, mode, k2, texels, prior_written=None):
    """CorrespondMap.update semantics (corrmap.py:672-736, ignore_obj_mat_id / no masks) with the last writer made explicit:
    `replace` = last pixel of the last frame that shows the texel; `first` = last pixel of the EARLIEST frame that shows it,
    among texels not written before the call.  Returns (values fp16 [k2*texels,4], written bool)."""
    F, H, W, _ = ids.shape
    n = k2 * texels
    dev = ids.device
    texel = (ids[..., 2].long() * texels + ids[..., 3].long()).reshape(-1)
    pix = torch.arange(H * W, device=dev).repeat(F)
    frame = torch.arange(F, device=dev).repeat_interleave(H * W)
    rank = (frame if mode == "replace" else (F - 1 - frame)) * (H * W) + pix
    owner = torch.full((n,), -1, dtype=torch.int64, device=dev)
    owner.scatter_reduce_(0, texel, rank, reduce="amax")
    written = owner >= 0
    if prior_written is not None and mode == "first":
        written &= ~prior_written
    own = owner.clamp(min=0)
    src_frame = own // (H * W)
    if mode == "first":
        src_frame = F - 1 - src_frame
    src = src_frame * (H * W) + own % (H * W)
    rgb = colors.reshape(-1, colors.shape[-1])[src]
    rgba = torch.cat([rgb, torch.ones_like(rgb[:, :1])], dim=1).half()                 # alpha appended (corrmap.py:681-684)
    values = torch.where(written.unsqueeze(1), rgba, torch.zeros_like(rgba))
    return values, written


@pytest.mark.parametrize("mode", ["replace", "first"])
def test_full_size_bake_cfg4_bit_exact(mode):
    """config 4: 64 views 1024^2 (RGB f32) into a 4096^2 fp16 RGBA atlas, bit for bit; then a second `first` call keeps it."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.corrmap import CorrespondMap
    V, H, tex = 64, 1024, 4096
    ids = synthetic.make_ids(V, H, H, tex_h=tex, tex_w=tex, k=1, frac_2048=0.0, seed=99, device="cuda")
    keep = (ids != 0).any(dim=-1)
    colors = torch.rand(V, H, H, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    masks = (~keep).float()                                                               # IDMap.masks: 1 = no id
    cm = CorrespondMap(name="t", k=1, height=tex, width=tex, channel_count=4, device="cuda")
    cm.update(colors, ids, mode=mode, masks=masks, inverse_masks=True, ignore_obj_mat_id=True)
    # pixels without an id are masked out: give them a texel outside the atlas for the restatement by dropping them
    ids_kept = ids.clone()
    ids_kept[..., 2][~keep] = 0
    flat_tex = torch.where(keep, ids[..., 3], torch.full_like(ids[..., 3], tex * tex))   # dump slot for masked pixels
    ids_kept[..., 3] = flat_tex
    values, written = _torch_bake(colors, ids_kept, mode, 1, tex * tex + 1)
    values, written = values[:-1], written[:-1]
    assert torch.equal(cm._writtens.reshape(-1).bool(), written)
    assert torch.equal(cm._values.reshape(-1, 4).view(torch.int16), values.view(torch.int16))
    if mode == "first":
        before = cm._values.clone()
        cm.update(colors.flip(0), ids.flip(0), mode="first", masks=masks.flip(0), inverse_masks=True, ignore_obj_mat_id=True)
        assert torch.equal(cm._values, before)                                            # every texel of these views is already written
