"""-m gpu tests of the persistent single-kernel overlap step (csrc/srx_fused.cu): against the numpy oracle, against the
split reduce + gather kernels, over consecutive steps (double-buffered accumulators / device step counter), on ragged
rows, replayed from a CUDA graph, and its in-kernel peer exchange emulated with two ranks on one GPU (plus a real
2-GPU run when two devices are visible)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import srx_oracle as O
from helpers import assert_close, t2n

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 3e-6
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inputs(F, H, W, tex, seed, id_dtype=torch.int32, frac_2048=0.05, n_obj=1):
    from stable_renderer_b200 import synthetic
    ids = synthetic.make_ids(F, H, W, tex_h=tex, tex_w=tex, frac_2048=frac_2048, seed=seed, dtype=id_dtype, n_obj=n_obj)
    x = synthetic.make_latents(F, 4, H // 8, W // 8, seed=seed + 1)
    return ids, x


@pytest.mark.parametrize("shape", [(4, 256, 256), (3, 320, 320), (1, 576, 576)])
@pytest.mark.parametrize("adain", [True, False])
def test_fused_step_matches_oracle(shape, adain):
    """w = 32 (one full chunk per row), w = 40 (a 32-cell and an 8-cell chunk) and w = 72 (two full chunks + 8 cells).
    Only square id buffers are valid: the reference divides x by the HEIGHT (corrmap.py:239)."""
    from stable_renderer_b200.plan import OverlapPlan
    F, H, W = shape
    ids, x0 = _inputs(F, H, W, 128, seed=11)
    want, parts = O.overlap_step(x0.numpy(), ids.numpy(), None, ratio=0.5, accumulate="f64", return_parts=True)
    if not adain:
        want = parts["blended"]          # adain=False writes the blended latents (before AdaIN)
    x = x0.cuda()
    plan = OverlapPlan(ids.cuda(), x.shape, key_capacity=128 * 128)
    assert plan.fused and plan.fast_path
    plan.step(x, 0.5, adain=adain)
    plan.check()
    assert_close(t2n(x), want, RTOL, ATOL, f"fused {shape} adain={adain}")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, (1e-5, 3e-6)), (torch.float16, (1e-2, 1e-2)), (torch.bfloat16, (1e-2, 1e-2))])
@pytest.mark.parametrize("id_dtype", [torch.int32, torch.int16])
def test_fused_equals_split_kernels(dtype, tol, id_dtype):
    from stable_renderer_b200.plan import OverlapPlan
    tex = 128 if id_dtype == torch.int16 else 256     # int16 vertex ids must stay below 32768
    ids, x0 = _inputs(6, 256, 256, tex, seed=5, id_dtype=id_dtype)
    ids = ids.cuda()
    a = x0.to(dtype).cuda()
    b = a.clone()
    pf = OverlapPlan(ids, a.shape, key_capacity=tex * tex)
    ps = OverlapPlan(ids, b.shape, key_capacity=tex * tex, split_kernels=True)
    assert pf.fused and not ps.fused
    pf.step(a, 0.3)
    ps.step(b, 0.3)
    torch.cuda.synchronize()
    assert_close(t2n(a), t2n(b), tol[0], tol[1], f"fused vs split {dtype} {id_dtype}")


def test_fused_consecutive_steps_and_graph_replay():
    """Five steps on one plan (accumulator parity, statistics double buffer, barrier targets from the device step
    counter) eagerly and replayed from a CUDA graph; both must track the oracle applied five times."""
    from stable_renderer_b200.plan import OverlapPlan
    ids, x0 = _inputs(4, 256, 256, 128, seed=21)
    want = x0.numpy()
    for _ in range(5):
        want = O.overlap_step(want, ids.numpy(), None, ratio=0.5, accumulate="f64")
    ids_d = ids.cuda()
    x = x0.cuda()
    plan = OverlapPlan(ids_d, x.shape, key_capacity=128 * 128)
    for _ in range(5):
        plan.step(x, 0.5)
    plan.check()
    assert_close(t2n(x), want, 5e-5, 1e-5, "5 eager steps")

    xg = x0.cuda()
    plan2 = OverlapPlan(ids_d, xg.shape, key_capacity=128 * 128)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        plan2.step(xg, 0.5)                       # step 1 eagerly (warm-up outside capture)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        plan2.step(xg, 0.5)
    for _ in range(4):                            # capture does not execute: 4 replays = steps 2..5
        graph.replay()
    torch.cuda.synchronize()
    plan2.check()
    assert_close(t2n(xg), want, 5e-5, 1e-5, "1 eager + 4 replayed steps")


def test_fused_steps_with_changing_ids():
    """Consecutive steps on DIFFERENT id batches (the streaming regime): a step that skipped or repeated work items would
    leave winners of the previous batch behind or weigh frames twice."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.plan import OverlapPlan
    F, H, tex = 6, 512, 256
    ids0 = synthetic.make_ids(F, H, H, tex_h=tex, tex_w=tex, frac_2048=0.05, seed=8)
    batches = [ids0, torch.roll(ids0, 1, 0).contiguous(), synthetic.make_ids(F, H, H, tex_h=tex, tex_w=tex, frac_2048=0.2, seed=9)]
    x0 = synthetic.make_latents(F, 4, H // 8, H // 8, seed=3)
    want = x0.numpy()
    x = x0.cuda()
    plan = OverlapPlan(None, x.shape, id_shape=ids0.shape, id_dtype=ids0.dtype, key_capacity=tex * tex, device=x.device)
    dev_batches = [b.cuda() for b in batches]
    for i in range(5):
        want = O.overlap_step(want, batches[i % 3].numpy(), None, ratio=0.5, accumulate="f64")
        plan.step(x, 0.5, ids=dev_batches[i % 3])
        assert_close(t2n(x), want, 5e-5, 1e-5, f"step {i} on id batch {i % 3}")
    plan.check()


def test_fused_key_out_of_range_is_reported():
    from stable_renderer_b200 import _lib
    from stable_renderer_b200.plan import OverlapPlan
    ids, x0 = _inputs(2, 128, 128, 128, seed=3)
    x = x0.cuda()
    plan = OverlapPlan(ids.cuda(), x.shape, key_capacity=1000)     # vertex ids reach 16383
    assert plan.fused
    plan.step(x, 0.5)
    with pytest.raises(_lib.SrxError):
        plan.check()


def test_fused_peer_exchange_two_ranks_on_one_gpu():
    """The in-kernel accumulator exchange (phase X) with two 'ranks' sharing this GPU: each plan runs on half of the SMs
    on its own stream, the peers' workspaces are each other's.  Must equal one plan over all frames — for several
    consecutive steps, since each step's barriers build on the previous step's counters."""
    from stable_renderer_b200 import _lib
    from stable_renderer_b200.plan import OverlapPlan
    sms = _lib.load().srx_device_sm_count()
    F, H = 8, 256
    ids, x0 = _inputs(F, H, H, 256, seed=77)
    ids = ids.cuda()
    ref = x0.cuda()
    cap = 256 * 256
    pref = OverlapPlan(ids, ref.shape, key_capacity=cap)
    xs = [x0[r * F // 2:(r + 1) * F // 2].contiguous().cuda() for r in range(2)]
    plans = [OverlapPlan(ids[r * F // 2:(r + 1) * F // 2].contiguous(), xs[r].shape, key_capacity=cap) for r in range(2)]
    ptrs = [p.workspace.data_ptr() for p in plans]
    for r, p in enumerate(plans):
        p.set_grid(sms // 2)
        p.bind_peers(r, ptrs)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for step in range(3):
        pref.step(ref, 0.5)
        torch.cuda.synchronize()
        for r in range(2):
            with torch.cuda.stream(streams[r]):
                plans[r].step(xs[r], 0.5)
        torch.cuda.synchronize()
        got = torch.cat(xs, dim=0)
        assert_close(t2n(got), t2n(ref), 2e-5, 5e-6, f"peer exchange, step {step}")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("nvls", [True, False])
def test_fused_peer_exchange_two_gpus(nvls):
    """Real frame-sharded run: 2 processes, symmetric-memory workspaces, exchange inside the step kernel over NVLink — the
    NVLS form (in-switch reduction + multicast of the totals) and the pull form."""
    script = os.path.join(ROOT, "tests", "mp_peer_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533" if nvls else "29534", script]
    env = dict(os.environ)
    env.pop("SRX_NVLS", None)
    if nvls:
        env["SRX_NVLS"] = "1"
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    assert "PEER_OK" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert f"nvls={nvls}" in proc.stdout


def test_fused_large_vertex_ids_round_like_float32():
    """Vertex ids in [2^24, 2^25) are not exact in the reference's float32 rows (corrmap.py:256-261): neighbours collapse
    onto one key.  The persistent kernel rounds in integer arithmetic; the split kernels convert through float."""
    from stable_renderer_b200.plan import OverlapPlan
    H = 64
    gen = torch.Generator().manual_seed(5)
    ids = torch.zeros(2, H, H, 4, dtype=torch.int32)
    ids[..., 0] = 1
    ids[..., 3] = (1 << 24) - 40 + torch.randint(0, 200, (2, H, H), generator=gen, dtype=torch.int32)
    x0 = torch.randn(2, 4, 8, 8, generator=gen)
    want = O.overlap_step(x0.numpy(), ids.numpy(), None, ratio=0.5, accumulate="f64")
    cap = (1 << 24) + 256
    a, b = x0.cuda(), x0.cuda()
    pf = OverlapPlan(ids.cuda(), a.shape, key_capacity=cap)
    ps = OverlapPlan(ids.cuda(), b.shape, key_capacity=cap, split_kernels=True)
    assert pf.fused and not ps.fused
    pf.step(a, 0.5)
    ps.step(b, 0.5)
    pf.check()
    assert_close(t2n(a), want, RTOL, ATOL, "fused, ids around 2^24")
    assert_close(t2n(b), want, RTOL, ATOL, "split, ids around 2^24")


@pytest.mark.parametrize("id_dtype", [torch.int32, torch.int16])
def test_cached_plan_steps_match_oracle(id_dtype):
    """Cached-plan regime: one bucketing pass over the ids, then several steps that never touch the ids again."""
    from stable_renderer_b200.plan import OverlapPlan
    ids, x0 = _inputs(5, 320, 320, 128, seed=41, id_dtype=id_dtype)
    want = x0.numpy()
    for _ in range(3):
        want = O.overlap_step(want, ids.numpy(), None, ratio=0.5, accumulate="f64")
    x = x0.cuda()
    plan = OverlapPlan(ids.cuda(), x.shape, key_capacity=128 * 128)
    plan.build_cache()
    kept, cap = plan.cache_entries()
    assert 0 < kept <= cap                      # only pairs whose key wins a cell are kept
    n_valid = int(((ids[..., 2] != 2048) & (ids.abs().sum(-1) != 0)).sum())
    assert cap <= n_valid                       # reduce-by-key inside the cell never produces more pairs than entries
    for _ in range(3):
        plan.step(x, 0.5, cached=True)
    plan.check()
    assert_close(t2n(x), want, 3e-5, 1e-5, f"3 cached steps {id_dtype}")
    # a streaming step after cached steps keeps working on the same plan (shared accumulators / step counter)
    want4 = O.overlap_step(want, ids.numpy(), None, ratio=0.5, accumulate="f64")
    plan.step(x, 0.5)
    assert_close(t2n(x), want4, 4e-5, 1e-5, "streaming step after cached steps")


def test_cached_plan_requires_build():
    from stable_renderer_b200 import _lib
    from stable_renderer_b200.plan import OverlapPlan
    ids, x0 = _inputs(2, 128, 128, 128, seed=3)
    x = x0.cuda()
    plan = OverlapPlan(ids.cuda(), x.shape, key_capacity=128 * 128)
    with pytest.raises(_lib.SrxError):
        plan.step(x, 0.5, cached=True)


def test_corresponder_switches_to_cached_plan():
    """The sampler-loop drop-in buckets the ids once it knows more steps follow (SamplingCallbackContext.total_steps)."""
    from helpers import Ctx, EngineData
    from stable_renderer_b200.corresponder import OverlapCorresponder
    from stable_renderer_b200.corrmap import IDMap
    ids, x0 = _inputs(4, 256, 256, 128, seed=9)
    want = x0.numpy()
    for _ in range(4):
        want = O.overlap_step(want, ids.numpy(), None, ratio=0.3, accumulate="f64")
    x = x0.cuda()
    idm = IDMap(tensor=ids.cuda())
    oc = OverlapCorresponder(step_finished_inject_ratio=0.3)
    for i in range(4):
        oc.step_finished(EngineData(idm), Ctx(x, 900.0, step_index=i, total_steps=20))
    plan = next(iter(idm._plans.values()))
    assert plan.fused and plan.cached
    assert_close(t2n(x), want, 4e-5, 1e-5, "4 steps through the corresponder (cached plan)")
    # without a step count the first call streams and the second one buckets
    x2 = x0.cuda()
    idm2 = IDMap(tensor=ids.cuda())

    class Bare:
        noise, timestep = x2, 900.0
    oc.step_finished(EngineData(idm2), Bare())
    p2 = next(iter(idm2._plans.values()))
    assert not getattr(p2, "cached", False)
    oc.step_finished(EngineData(idm2), Bare())
    assert p2.cached


def test_cached_plan_peer_exchange_two_ranks_on_one_gpu():
    from stable_renderer_b200 import _lib
    from stable_renderer_b200.plan import OverlapPlan
    sms = _lib.load().srx_device_sm_count()
    F, H = 8, 256
    ids, x0 = _inputs(F, H, H, 256, seed=78)
    ids = ids.cuda()
    ref = x0.cuda()
    cap = 256 * 256
    pref = OverlapPlan(ids, ref.shape, key_capacity=cap)
    xs = [x0[r * F // 2:(r + 1) * F // 2].contiguous().cuda() for r in range(2)]
    plans = [OverlapPlan(ids[r * F // 2:(r + 1) * F // 2].contiguous(), xs[r].shape, key_capacity=cap) for r in range(2)]
    ptrs = [p.workspace.data_ptr() for p in plans]
    for r, p in enumerate(plans):
        p.set_grid(sms // 2)
        p.bind_peers(r, ptrs)
        p.cache_mark(p._ids)
    union = plans[0].need_map | plans[1].need_map      # what the MAX all-reduce does in a real frame-sharded run
    for p in plans:
        p.need_map.copy_(union)
        p.cache_emit(p._ids)
        p.cached = True
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for step in range(2):
        pref.step(ref, 0.5)
        torch.cuda.synchronize()
        for r in range(2):
            with torch.cuda.stream(streams[r]):
                plans[r].step(xs[r], 0.5, cached=True)
        torch.cuda.synchronize()
        assert_close(t2n(torch.cat(xs, dim=0)), t2n(ref), 2e-5, 5e-6, f"cached + peer exchange, step {step}")


@pytest.mark.parametrize("frames", [5, 160])
def test_fused_frame_grouping_paths(frames):
    """Gather / AdaIN scheduling: 5 frames -> groups of 29 CTAs per frame with idle CTAs left over; 160 frames -> more
    frames than CTAs, every CTA walks over several frames on its own (no statistics exchange)."""
    from stable_renderer_b200.plan import OverlapPlan
    H = 64
    ids, x0 = _inputs(frames, H, H, 64, seed=frames)
    want = O.overlap_step(x0.numpy(), ids.numpy(), None, ratio=0.4, accumulate="f64")
    x = x0.cuda()
    plan = OverlapPlan(ids.cuda(), x.shape, key_capacity=64 * 64)
    assert plan.fused
    plan.step(x, 0.4)
    plan.check()
    assert_close(t2n(x), want, RTOL, ATOL, f"{frames} frames")
    plan.build_cache()
    x2 = x0.cuda()
    plan.step(x2, 0.4, cached=True)
    assert_close(t2n(x2), want, RTOL, ATOL, f"{frames} frames, cached plan")


def test_fused_peer_exchange_three_ranks_on_one_gpu():
    """Three emulated ranks (a third of the SMs each): owner slices with a remainder (capacity 10,048 / 3), ring order of the
    pulls, totals added in rank order."""
    from stable_renderer_b200 import _lib
    from stable_renderer_b200.plan import OverlapPlan
    sms = _lib.load().srx_device_sm_count()
    F, H, tex = 9, 128, 100
    ids, x0 = _inputs(F, H, H, tex, seed=79)
    ids = ids.cuda()
    ref = x0.cuda()
    cap = tex * tex + 48                       # 10,048 after rounding up to 64: not divisible by 3
    pref = OverlapPlan(ids, ref.shape, key_capacity=cap)
    xs = [x0[r * 3:(r + 1) * 3].contiguous().cuda() for r in range(3)]
    plans = [OverlapPlan(ids[r * 3:(r + 1) * 3].contiguous(), xs[r].shape, key_capacity=cap) for r in range(3)]
    ptrs = [p.workspace.data_ptr() for p in plans]
    for r, p in enumerate(plans):
        p.set_grid(sms // 3)
        p.bind_peers(r, ptrs)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(3)]
    for step in range(2):
        pref.step(ref, 0.5)
        torch.cuda.synchronize()
        for r in range(3):
            with torch.cuda.stream(streams[r]):
                plans[r].step(xs[r], 0.5)
        torch.cuda.synchronize()
        assert_close(t2n(torch.cat(xs, dim=0)), t2n(ref), 2e-5, 5e-6, f"3-rank peer exchange, step {step}")


def test_fused_adain_flag_may_change_between_steps_of_one_plan():
    """adain is a per-call flag and plans are shared (cached on the IDMap): steps that skip AdaIN must not leave the
    statistics exchange of a later AdaIN step waiting (batch < SM count, so several CTAs share a frame)."""
    from stable_renderer_b200.plan import OverlapPlan
    ids, x0 = _inputs(4, 256, 256, 128, seed=21)
    want, parts = O.overlap_step(x0.numpy(), ids.numpy(), None, ratio=0.5, accumulate="f64", return_parts=True)
    plan = OverlapPlan(ids.cuda(), x0.shape, key_capacity=128 * 128)
    assert plan.fused
    for adain in (False, False, True, False, True, True):
        x = x0.cuda()
        plan.step(x, 0.5, adain=adain)
        plan.check()
        assert_close(t2n(x), want if adain else parts["blended"], RTOL, ATOL, f"adain={adain}")


def test_fused_set_grid_after_steps():
    """Changing the grid of a plan that has already stepped restarts the device-side step counter, barrier counters,
    accumulators and statistics records together."""
    from stable_renderer_b200 import _lib
    from stable_renderer_b200.plan import OverlapPlan
    sms = _lib.load().srx_device_sm_count()
    ids, x0 = _inputs(5, 256, 256, 128, seed=23)
    want = O.overlap_step(x0.numpy(), ids.numpy(), None, ratio=0.5, accumulate="f64")
    plan = OverlapPlan(ids.cuda(), x0.shape, key_capacity=128 * 128)
    for grid in (0, sms // 2, sms // 3, 0):
        plan.set_grid(grid)
        for _ in range(3):       # an odd number of steps leaves the "other" accumulator dirty
            x = x0.cuda()
            plan.step(x, 0.5)
        plan.check()
        assert_close(t2n(x), want, RTOL, ATOL, f"grid {grid}")


def test_fused_lost_peer_raises_instead_of_hanging():
    """A rank whose peer never steps must not hang inside the cooperative kernel: its waits give up after ~2 s, the step
    finishes (with invalid latents) and `check()` reports the lost peer."""
    from stable_renderer_b200 import _lib
    from stable_renderer_b200.plan import OverlapPlan
    sms = _lib.load().srx_device_sm_count()
    ids, x0 = _inputs(4, 128, 128, 64, seed=29)
    ids = ids.cuda()
    plans = [OverlapPlan(ids[r * 2:(r + 1) * 2].contiguous(), (2, 4, 16, 16), key_capacity=64 * 64) for r in range(2)]
    ptrs = [p.workspace.data_ptr() for p in plans]
    for r, p in enumerate(plans):
        p.set_grid(sms // 2)
        p.bind_peers(r, ptrs)
    x = x0[:2].contiguous().cuda()
    plans[0].step(x, 0.5)            # rank 1 never steps
    with pytest.raises(_lib.SrxError, match="never arrived|gave up"):
        plans[0].check()


def test_idmap_invalidate_rebinds_cached_plans_to_the_new_ids():
    """IDMap.invalidate() with a CPU tensor (the `from_directory` default): the cached plan must stream the NEW ids, in
    the streaming regime and after re-bucketing."""
    from helpers import Ctx, EngineData
    from stable_renderer_b200.corresponder import OverlapCorresponder
    from stable_renderer_b200.corrmap import IDMap
    ids_a, x0 = _inputs(4, 256, 256, 128, seed=31)
    ids_b, _ = _inputs(4, 256, 256, 128, seed=37)
    host = ids_a.clone()
    idm = IDMap(tensor=host)
    ed = EngineData(id_maps=idm)
    oc = OverlapCorresponder(step_finished_inject_ratio=0.5, key_capacity=128 * 128)
    for ids in (ids_a, ids_b, ids_a):
        host.copy_(ids)
        idm.invalidate()
        want = O.overlap_step(x0.numpy(), ids.numpy(), None, ratio=0.5, accumulate="f64")
        for s in range(3):          # step 0 streams, later steps run from the cached plan
            ctx = Ctx(x0.cuda(), step_index=s, total_steps=3)
            oc.step_finished(ed, ctx)
            assert_close(t2n(ctx.noise), want, RTOL, ATOL, f"step {s}")
