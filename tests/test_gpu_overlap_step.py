"""-m gpu parity tests of the current-generation overlap step: CUDA path (through the C ABI) vs the reference-generated
golden fixtures and vs the numpy oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

import srx_oracle as O
from helpers import Ctx, EngineData, assert_close, t2n

pytestmark = pytest.mark.gpu

STEP_CASES = ["step_sq64_r8", "step_sq96_r8_perm", "step_sq100_nonint", "step_sq60_to_16", "step_dupframe",
              "step_sphere_crop_int16"]
# float32 tolerance of north_star: 1e-5 relative (plus a few float32 ulps of the O(1) latents as absolute floor)
RTOL, ATOL = 1e-5, 3e-6


def _run_step(ids_np, x_np, ratio, frame_indices=None, dtype=torch.float32, timestep=900.0, stop=500, **kw):
    from stable_renderer_b200.corresponder import OverlapCorresponder
    from stable_renderer_b200.corrmap import IDMap
    ids = torch.from_numpy(ids_np).cuda()
    x = torch.from_numpy(x_np).to(dtype).cuda()
    idm = IDMap(tensor=ids, frame_indices=None if frame_indices is None else [int(v) for v in frame_indices])
    oc = OverlapCorresponder(step_finished_inject_ratio=ratio, step_finished_stop_inject_timestep=stop, **kw)
    ctx = Ctx(x, timestep)
    ret = oc.step_finished(EngineData(idm), ctx)
    assert ret is None
    assert ctx.noise is x            # mutated in place, like the reference (corresponder.py:375-376)
    torch.cuda.synchronize()
    return x, idm


@pytest.mark.parametrize("name", STEP_CASES)
def test_step_matches_reference_golden(golden, name):
    g = golden(name)
    fi = g["frame_indices"] if "frame_indices" in g.files else None
    x, idm = _run_step(g["ids"], g["x"], float(g["ratio"]), fi)
    assert_close(t2n(x), g["out"], RTOL, ATOL, name)
    expect_fast = name in ("step_sq64_r8", "step_sq96_r8_perm", "step_sphere_crop_int16")
    plan = next(iter(idm._plans.values()))
    assert plan.fast_path == expect_fast


@pytest.mark.parametrize("name", STEP_CASES)
def test_step_matches_oracle(golden, name):
    g = golden(name)
    fi = g["frame_indices"] if "frame_indices" in g.files else None
    want = O.overlap_step(g["x"], g["ids"], fi, ratio=float(g["ratio"]), accumulate="f64")
    x, _ = _run_step(g["ids"], g["x"], float(g["ratio"]), fi)
    assert_close(t2n(x), want, RTOL, ATOL, name)


@pytest.mark.parametrize("name", ["step_sq64_r8", "step_sq100_nonint"])
def test_step_deterministic_mode(golden, name):
    g = golden(name)
    want = O.overlap_step(g["x"], g["ids"], None, ratio=float(g["ratio"]), accumulate="f64")
    a, _ = _run_step(g["ids"], g["x"], float(g["ratio"]), deterministic=True)
    b, _ = _run_step(g["ids"], g["x"], float(g["ratio"]), deterministic=True)
    assert torch.equal(a, b)                         # order-independent accumulation: bit reproducible
    assert_close(t2n(a), want, RTOL, ATOL, name)


def test_step_gate(golden):
    g = golden("step_gate_off")
    x, idm = _run_step(g["ids"], g["x"], 0.5, timestep=float(g["timestep"]), stop=float(g["stop"]))
    assert np.array_equal(t2n(x), g["x"])
    assert not idm._plans                             # nothing was even planned


@pytest.mark.parametrize("tag,dtype", [("f16", torch.float16), ("bf16", torch.bfloat16)])
def test_step_half_precision(golden, tag, dtype):
    g = golden(f"step_half_{tag}")
    x, _ = _run_step(g["ids"], g["x"], float(g["ratio"]), dtype=dtype)
    assert x.dtype == dtype
    # (1) against the exact result on the same half-precision inputs: fp32 arithmetic, ONE rounding to the latent dtype, so
    #     half an ulp of that dtype (fp16: 2^-11, bf16: 2^-8 relative) — well inside north_star's 1e-2
    exact = O.overlap_step(g["x"].astype(np.float32), g["ids"], None, ratio=float(g["ratio"]), accumulate="f64")
    half_ulp = 2.0 ** -11 if dtype == torch.float16 else 2.0 ** -8
    assert_close(t2n(x), exact, half_ulp * 1.01, half_ulp * 1.01, tag + " vs exact")
    # (2) against the reference's own half-precision output.  The reference blends and normalises IN the half type (several
    #     roundings, corresponder.py:351-352, math_utils.py:39-80): its bf16 fixture is itself up to 1.15e-2 * max(1, |x|) away
    #     from the exact result (measured), i.e. more than 1e-2 — hence 1e-2 relative plus one bf16 ulp at |x| < 4 (2^-6)
    assert_close(t2n(x), g["out"], 1e-2, 1e-2 if dtype == torch.float16 else 2.0 ** -6, tag)


def test_step_key_capacity_hint_and_overflow(golden):
    g = golden("step_sq64_r8")
    kmax = int(g["ids"][..., 3].max())
    x, _ = _run_step(g["ids"], g["x"], float(g["ratio"]), key_capacity=kmax + 1)
    assert_close(t2n(x), g["out"], RTOL, ATOL)
    from stable_renderer_b200 import _lib
    with pytest.raises(_lib.SrxError):
        _run_step(g["ids"], g["x"], float(g["ratio"]), key_capacity=max(kmax // 2, 1))


def test_step_out_of_range_cell_raises():
    ids = np.zeros((1, 8, 16, 4), dtype=np.int32)    # W > H: x / H * w leaves the latent (corrmap.py:239)
    ids[..., 3] = 5
    with pytest.raises(IndexError):
        _run_step(ids, np.zeros((1, 4, 2, 4), dtype=np.float32), 0.5)


def test_step_bad_frame_index_raises(golden):
    g = golden("step_sq64_r8")
    with pytest.raises(IndexError):
        _run_step(g["ids"], g["x"], 0.5, frame_indices=[0, 1, 2, 7])


def test_step_no_valid_pixels_is_adain_identity():
    ids = np.zeros((2, 64, 64, 4), dtype=np.int32)
    x0 = np.random.default_rng(0).standard_normal((2, 4, 8, 8)).astype(np.float32)
    x, _ = _run_step(ids, x0, 0.5)
    assert_close(t2n(x), O.overlap_step(x0, ids, None, 0.5), RTOL, ATOL)


def test_step_non_contiguous_noise(golden):
    g = golden("step_sq64_r8")
    from stable_renderer_b200.corresponder import OverlapCorresponder
    from stable_renderer_b200.corrmap import IDMap
    base = torch.from_numpy(g["x"]).cuda()
    noise = base.permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2)   # same values, non-contiguous strides
    assert not noise.is_contiguous()
    oc = OverlapCorresponder(step_finished_inject_ratio=float(g["ratio"]))
    oc.step_finished(EngineData(IDMap(tensor=torch.from_numpy(g["ids"]).cuda())), Ctx(noise))
    assert_close(t2n(noise), g["out"], RTOL, ATOL)


@pytest.mark.parametrize("cfg", ["cfg1_like", "sdxl_slice"])
def test_step_full_size_vs_oracle(cfg):
    """cfg1 shape (16 x 512^2 ids, 64x64x4 latents) and an SDXL slice (4 x 1024^2, 128x128x4 bf16-sized cells in fp32)."""
    from stable_renderer_b200 import synthetic
    if cfg == "cfg1_like":
        F, H, h, tex = 16, 512, 64, 512
    else:
        F, H, h, tex = 4, 1024, 128, 512
    ids = synthetic.make_ids(F, H, H, tex_h=tex, tex_w=tex, frac_2048=0.05, seed=1235)
    x0 = synthetic.make_latents(F, 4, h, h, seed=0)
    want = O.overlap_step(x0.numpy(), ids.numpy(), None, ratio=0.5, accumulate="f64")
    x, idm = _run_step(ids.numpy(), x0.numpy(), 0.5)
    assert next(iter(idm._plans.values())).fast_path
    assert_close(t2n(x), want, RTOL, ATOL, cfg)
    # the generic kernel must agree with the warp-per-cell kernel on the same input (perm of frames forces nothing;
    # duplicate frame indices force the ordered path) — compare against the oracle again
    fi = list(range(F))
    fi[-1] = fi[-2]
    want2 = O.overlap_step(x0.numpy(), ids.numpy(), fi, ratio=0.5, accumulate="f64")
    x2, idm2 = _run_step(ids.numpy(), x0.numpy(), 0.5, frame_indices=fi)
    assert not next(iter(idm2._plans.values())).fast_path
    assert_close(t2n(x2), want2, RTOL, ATOL, cfg + " generic")


def test_split_reduce_allreduce_gather_equals_single_plan():
    """Frame-sharded form emulated on one GPU: two 'ranks' reduce their own frames, the accumulators
    are summed (what the NCCL all-reduce does), each rank gathers its frames -> same result as one plan over all frames."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.plan import OverlapPlan
    F, H, h = 8, 256, 32
    ids = synthetic.make_ids(F, H, H, tex_h=256, tex_w=256, seed=77).cuda()
    x0 = synthetic.make_latents(F, 4, h, h, seed=1).cuda()
    cap = 256 * 256
    ref = x0.clone()
    OverlapPlan(ids, ref.shape, key_capacity=cap).step(ref, 0.5)
    halves = []
    plans = []
    for r in range(2):
        sl = slice(r * F // 2, (r + 1) * F // 2)
        xr = x0[sl].clone().contiguous()
        p = OverlapPlan(ids[sl].contiguous(), xr.shape, key_capacity=cap)
        p.reduce(xr)
        halves.append(xr)
        plans.append(p)
    total = plans[0].accumulator + plans[1].accumulator
    for p in plans:
        p.accumulator.copy_(total)
    for p, xr in zip(plans, halves):
        p.gather(xr, 0.5)
    torch.cuda.synchronize()
    got = torch.cat(halves, dim=0)
    assert_close(t2n(got), t2n(ref), 1e-5, 3e-6)
