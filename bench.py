#!/usr/bin/env python
"""bench.py — latent-px/s (and overlap-steps/s) of the correspondence-map latent overlap step on N B200s.

`value` is the whole-job aggregate latent-px/s = (frames over all GPUs) * h * w * steps/s — the one of BASELINE.json's
two metrics that adds up over GPUs; `overlap_steps_per_sec` is reported beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one full `OverlapCorresponder.step_finished` on ids that are NEW for the step (streaming regime,
SURVEY.md §8d): key the id buffers, segment-reduce the latents per key, [all-reduce the key-indexed accumulator
when frames are sharded], gather/blend, AdaIN, in place.  Workloads (BASELINE.json configs):
    cfg1  16 frames  512^2 ids,  64x64x4 f32          cfg2  32 frames/GPU 512^2, 64x64x4 f32   (weak scaling)
    cfg3  96 frames 1024^2 ids, 128x128x4 bf16, frames sharded over the GPUs (strong scaling) — the DEFAULT: the config
          BASELINE.json's north_star target is quoted on; it fits one GPU, so N = 1/2/4/8 all run this named config
    cfg5  768 frames 512^2, 64x64x4 f32, 4 objects, sharded (strong scaling)
    bake  cfg4: 64 views 1024^2 RGB -> 4096^2 atlas, depth/normal weighted, views sharded (metric: views/s)

The default run also carries, in the same JSON line: `parity_check` (a small N-rank step through the same exchange path
checked against oracle/srx_oracle.py — the run exits non-zero when it fails), and `extra.cfg2` (streaming step, cached-plan
step and the 20-step job of BASELINE config 2) and, at N = 1, `extra.cfg4_bake` (BASELINE config 4).

Rank 0 prints ONE JSON line.  `value` is device-timed with inputs resident in HBM (CUDA events around exactly K
steps replayed from a CUDA graph, max over ranks); `e2e` goes through the public API with host buffers (pinned H2D of
ids + latents and D2H of the latents inside the timed region); `roofline` is the dominant kernel (the id-streaming
accumulate pass) timed per launch with CUDA events; `cpu_baseline` is oracle/torch_port.py — the reference's CPU torch
algorithm — on this box's host cores.  `--impl reference` times that CPU path alone.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

WORKLOADS = {
    #        frames  H     h    dtype            tex  n_obj scaling
    "cfg1": (16, 512, 64, torch.float32, 512, 1, "weak"),
    "cfg2": (32, 512, 64, torch.float32, 512, 1, "weak"),
    "cfg3": (96, 1024, 128, torch.bfloat16, 512, 1, "strong"),
    "cfg5": (768, 512, 64, torch.float32, 512, 4, "strong"),
}
RATIO = 0.5          # node default step_finished_inject_ratio (_nodes/samplers.py:81)
DTYPE_NAME = {torch.float32: "f32", torch.bfloat16: "bf16", torch.float16: "f16"}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML every ~10 ms while the timed regions run."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop, self._thr = [], set(), None, threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.nv = None
            log(f"[bench] NVML unavailable: {e}")

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self) -> dict:
        if self._thr is not None:
            self._stop.set()
            self._thr.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": int(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload: str):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, if one exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------------------
# distributed plumbing
# ---------------------------------------------------------------------------------------------------------------------
def init_dist(n_gpus: int):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's version banner must not land on stdout beside the JSON line
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif n_gpus > 1:
        raise SystemExit("for --gpus N > 1 launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
                         "--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")
    else:
        torch.cuda.set_device(0)
    return rank, local, world


def max_over_ranks(ms: float, world: int) -> float:
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world: int):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


# ---------------------------------------------------------------------------------------------------------------------
# overlap workloads
# ---------------------------------------------------------------------------------------------------------------------
def shard(frames_total: int, scaling: str, rank: int, world: int):
    if scaling == "weak":
        return frames_total, rank * frames_total, frames_total * world
    from stable_renderer_b200.sharding import frame_shard
    first, count = frame_shard(frames_total, rank, world)
    return count, first, frames_total


def workload_string(workload: str, frames_total: int) -> str:
    """`config.workload` — identical in our arm and in the reference arm."""
    _, H, h, dtype, _, _, _ = WORKLOADS[workload]
    return (f"{workload}: {frames_total} frames of {H}x{H}x4 int32 ids -> {h}x{h}x4 {DTYPE_NAME[dtype]} latents, "
            f"ratio {RATIO}, one overlap step per id batch pass (streaming regime)")


def percentile(v, q):
    s = sorted(v)
    return s[min(len(s) - 1, max(0, int(math.ceil(q * len(s))) - 1))]


def parity_check(rank: int, local: int, world: int, exchange: str, dtype: torch.dtype) -> dict:
    """A small N-rank overlap step (ragged frame shards, same exchange path as the timed run, streaming AND cached-plan
    regime) checked against the numpy oracle.  Outside every timed region.  oracle/ is used here as the checker only."""
    import numpy as np
    import torch.distributed as dist
    import srx_oracle as O
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.plan import OverlapPlan
    from stable_renderer_b200.sharding import frame_shard
    dev = torch.device("cuda", local)
    frames_total, H, h, tex = (2 * world + 1 if world > 1 else 5), 128, 16, 64
    ids = synthetic.make_ids(frames_total, H, H, tex_h=tex, tex_w=tex, frac_2048=0.05, seed=77)
    first, count = frame_shard(frames_total, rank, world)
    res = {"frames_total": frames_total, "regimes": ["streaming", "cached_plan"]}
    ok_all = True
    for dt, name in ((torch.float32, "f32"),) + (((dtype, DTYPE_NAME[dtype]),) if dtype != torch.float32 else ()):
        x0 = synthetic.make_latents(frames_total, 4, h, h, seed=3, dtype=dt)
        want = O.overlap_step(x0.float().numpy(), ids.numpy(), None, ratio=RATIO, accumulate="f64")[first:first + count]
        ids_dev = ids[first:first + count].contiguous().to(dev)
        plan = OverlapPlan(None, (count, 4, h, h), id_shape=ids_dev.shape, id_dtype=ids_dev.dtype, key_capacity=tex * tex,
                           device=dev, process_group=dist.group.WORLD if world > 1 else None, exchange=exchange)
        errs = []
        for regime in ("streaming", "cached_plan"):
            x = x0[first:first + count].contiguous().to(dev)
            if regime == "streaming":
                if world == 1 or plan.exchange == "peer":
                    plan.step(x, RATIO, ids=ids_dev)
                else:
                    plan.reduce(x, ids=ids_dev)
                    dist.all_reduce(plan.accumulator, op=dist.ReduceOp.SUM)
                    plan.gather(x, RATIO)
            else:
                if not (plan.fused and (world == 1 or plan.exchange == "peer")):
                    continue
                plan.build_cache(ids_dev)
                plan.step(x, RATIO, cached=True)
            plan.check()
            got = x.float().cpu().numpy()
            err = np.abs(got - want)
            # f32: 1e-5 relative (north_star) with an absolute floor for values near zero; 16-bit latents: 1e-2
            tol = (3e-6 + 1e-5 * np.abs(want)) if dt == torch.float32 else 1e-2 * np.maximum(1.0, np.abs(want))
            errs.append(float(err.max()))
            ok_all = ok_all and bool((err <= tol).all()) and bool(np.isfinite(got).all())
        plan.close()
        res[f"max_abs_err_{name}"] = max(errs)
    t = torch.tensor([max(v for k, v in res.items() if k.startswith("max_abs_err")), 0.0 if ok_all else 1.0],
                     dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for k in [k for k in res if k.startswith("max_abs_err")]:
            tk = torch.tensor([res[k]], dtype=torch.float64, device=dev)
            dist.all_reduce(tk, op=dist.ReduceOp.MAX)
            res[k] = float(tk.item())
    res["max_abs_err"] = float(t[0].item())
    res["ok"] = bool(t[1].item() == 0.0)
    res["ranks"] = world
    res["checker"] = "oracle/srx_oracle.py::overlap_step (float64 accumulation)"
    res["tolerance"] = "f32: 3e-6 + 1e-5*|ref|; f16/bf16: 1e-2*max(1,|ref|)"
    return res


def run_overlap(args, workload: str, rank: int, local: int, world: int, light: bool = False) -> dict:
    """light=True: the secondary record of another named config (fewer repetitions, no every-step-new-ids e2e)."""
    import torch.distributed as dist
    from stable_renderer_b200 import _lib, synthetic
    from stable_renderer_b200.corresponder import OverlapCorresponder
    from stable_renderer_b200.corrmap import IDMap
    from stable_renderer_b200.plan import OverlapPlan

    _lib.load()  # no extension, no benchmark
    frames_cfg, H, h, dtype, tex, n_obj, scaling = WORKLOADS[workload]
    F, f0, F_global = shard(frames_cfg, scaling, rank, world)
    dev = torch.device("cuda", local)
    key_capacity = tex * tex            # = CorrespondMap height*width: what the engine knows about the key space
    K, Wm = args.steps, args.warmup

    # ---- synthetic inputs ------------------------------------------------------------------------
    ids0 = synthetic.make_ids(F, H, H, tex_h=tex, tex_w=tex, n_obj=n_obj, frac_2048=0.05, seed=1234, device=dev,
                              frame_offset=f0)
    id_bytes = ids0.numel() * 4
    # rotate id buffers so that no step finds its ids in the 126 MB L2 (one buffer of >= 400 MB is "larger than L2" by itself)
    n_rot = 3 if id_bytes < (200 << 20) else 2 if id_bytes < (400 << 20) else 1
    ids_rot = [ids0] + [torch.roll(ids0, shifts=r, dims=0).contiguous() for r in range(1, n_rot)]
    x = synthetic.make_latents(F_global, 4, h, h, seed=0, dtype=dtype)[f0:f0 + F].contiguous().to(dev)
    x_init = x.clone()
    plan = OverlapPlan(None, x.shape, id_shape=ids0.shape, id_dtype=ids0.dtype, key_capacity=key_capacity, device=dev,
                       process_group=dist.group.WORLD if world > 1 else None, exchange=args.exchange,
                       split_kernels=args.split_kernels)
    acc = plan.accumulator
    peer = plan.exchange == "peer"

    def one_step(r: int):
        if world == 1 or peer:
            plan.step(x, RATIO, ids=ids_rot[r % n_rot])   # one persistent kernel (exchange over NVLink inside it when sharded)
        else:
            plan.reduce(x, ids=ids_rot[r % n_rot])
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
            plan.gather(x, RATIO)

    # ---- CUDA graphs holding EXACTLY K steps: q replays of an m-step graph + one r-step graph, no eager launch ------
    launch_mode = "cuda_graph"
    m = min(K, 48)
    if n_rot > 1 and m >= n_rot:
        m -= m % n_rot
    q, r_rem = divmod(K, m)
    graphs = []
    side = torch.cuda.Stream(device=dev)
    try:
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for r in range(max(n_rot, 2)):
                one_step(r)                         # warm everything outside capture (NCCL channels, occupancy queries)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for n_steps in (m, r_rem):
            if n_steps == 0:
                graphs.append(None)
                continue
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for r in range(n_steps):
                    one_step(r)
            graphs.append(g)
        torch.cuda.synchronize()
    except Exception as e:
        log(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); falling back to eager launches")
        graphs = []
        launch_mode = "eager"
        torch.cuda.synchronize()

    def run_k_steps():
        if graphs:
            for _ in range(q):
                graphs[0].replay()
            if graphs[1] is not None:
                graphs[1].replay()
        else:
            for i in range(K):
                one_step(i)

    sampler = ClockSampler(local)
    x.copy_(x_init)
    for i in range(Wm):
        one_step(i)
    if graphs:
        graphs[0].replay()                              # graph warm-up (first replay uploads the graph)
    sampler.start()
    barrier(world)
    # Two more untimed steps right in front of the start event: at N > 1 the step kernel is itself a rendezvous of all ranks, so
    # every rank's start event fires within microseconds of the others' on the device — without them the ranks' host-side skew
    # after the barrier (hundreds of microseconds) lands inside a timed region that is only K x 0.1 ms long.
    for i in range(2):
        one_step(i)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    run_k_steps()
    ev1.record()
    torch.cuda.synchronize()
    barrier(world)
    ms_total = max_over_ranks(ev0.elapsed_time(ev1), world)
    ms_step = ms_total / K
    plan.check()
    assert torch.isfinite(x.float()).all(), "latents became non-finite"

    # ---- dominant kernel alone, per-launch CUDA events (eager launches on the current stream) ----------------------
    n_k1 = min(K, 200 if not light else 50)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_k1)]
    elem = x.element_size()
    step_bytes = 16 * F * H * H + 2 * F * 4 * h * h * elem           # SURVEY.md §8d streaming-regime figure, per GPU
    fused_kernel = plan.fused and (world == 1 or peer)
    if fused_kernel:
        # the whole step is one kernel: ids streamed once, latents read + written once
        k1_name = "k_overlap_fused (persistent step: id streaming, key reduce, " + \
                  ("NVLink peer exchange, " if peer else "") + "gather/blend, AdaIN)"
        k1_bytes = step_bytes
        k1_call = lambda i: plan.step(x, RATIO, ids=ids_rot[i % n_rot])   # noqa: E731
    else:
        k1_name = "k_accum_r8 (id-streaming key/segment-reduce pass)"
        k1_bytes = 16 * F * H * H + F * 4 * h * h * elem             # ids streamed once + latents read once
        k1_call = lambda i: plan.reduce(x, ids=ids_rot[i % n_rot])    # noqa: E731
    x.copy_(x_init)
    barrier(world)
    for i in range(3):
        k1_call(i)
    for i, (a, b) in enumerate(evs):
        a.record()
        k1_call(i)
        b.record()
    torch.cuda.synchronize()
    k1_ms_eager = max_over_ranks(statistics.mean([a.elapsed_time(b) for a, b in evs]), world)
    clocks = sampler.stop()        # clocks are sampled during the device-timed regions; NVML polling would only
                                   # perturb the host-driven end-to-end loops below
    k1_ms = k1_ms_eager
    if fused_kernel:
        # one launch per step: the kernel's average launch duration over the timed region IS the step time
        k1_ms = ms_step
    else:
        acc.zero_()
    peak, peak_src = measured_hbm_peak()
    k1_gbs = k1_bytes / (k1_ms * 1e-3) / 1e9

    # ---- cached-plan regime: the 20 denoise steps of one sampling run share their ids (BASELINE config 2) ------------
    cached = None
    if fused_kernel:
        job_steps = 20
        # J independent sampling runs (own plan, pool, latents, accumulators) stepped round-robin, so that between two
        # steps of one run more than an L2's worth of other runs' data passes through: every cached step finds its
        # pool / latents / accumulators in HBM, as it would after a UNet forward.
        plans, xs_j = [plan], [x]
        plan.build_cache(ids_rot[0])
        kept, seen = plan.cache_entries()
        # footprint of one run, from rank-independent sizes (all ranks must build the same number of plans): two
        # accumulators, latents in + out, winners, and the pool estimated at 0.15 pairs per id pixel (cfg2 keeps 0.15)
        per_job = 2 * acc.numel() * 4 + 2 * x.numel() * elem + 4 * F * h * h + int(8 * 0.15 * F * H * H)
        n_jobs = max(2, min(24, int(math.ceil(1.5 * 126e6 / max(per_job, 1))) + 1))
        for j in range(1, n_jobs):
            pj = OverlapPlan(None, x.shape, id_shape=ids0.shape, id_dtype=ids0.dtype, key_capacity=key_capacity, device=dev,
                             process_group=dist.group.WORLD if world > 1 else None, exchange=args.exchange)
            plans.append(pj)
            xs_j.append(x_init.clone())
        x.copy_(x_init)

        def run_jobs(build: bool):
            if build:
                for j, pj in enumerate(plans):
                    pj.build_cache(ids_rot[j % n_rot])
            for _ in range(job_steps):
                for pj, xj in zip(plans, xs_j):
                    pj.step(xj, RATIO, cached=True)

        run_jobs(True)                                  # warm-up (also sizes the pools)
        barrier(world)
        for _ in range(2):                              # device-side alignment of the ranks, as above
            plans[0].step(xs_j[0], RATIO, cached=True)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        run_jobs(False)
        c1.record()
        torch.cuda.synchronize()
        cached_ms = max_over_ranks(c0.elapsed_time(c1), world) / (job_steps * n_jobs)
        barrier(world)
        j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw = time.perf_counter()
        j0.record()
        run_jobs(True)
        j1.record()
        torch.cuda.synchronize()
        tw = (time.perf_counter() - tw) * 1e3
        job_ms = max_over_ranks(max(j0.elapsed_time(j1), tw), world) / n_jobs
        for pj in plans[1:]:
            pj.close()
        cached_alg = 2 * F * 4 * h * h * elem + 8 * seen + 4 * F * h * h
        cached = {"ms_per_step": cached_ms, "steps_per_sec": 1e3 / cached_ms,
                  "latent_px_per_sec": F_global * h * h * 1e3 / cached_ms,
                  "job": {"steps": job_steps, "ms": job_ms, "steps_per_sec": job_steps * 1e3 / job_ms,
                          "latent_px_per_sec": F_global * h * h * job_steps * 1e3 / job_ms,
                          "includes": "bucketing pass on fresh ids (2 id passes) + 20 cached steps, eager launches"},
                  "pairs_kept_per_gpu": kept, "pairs_seen_per_gpu": seen,
                  "algorithmic_bytes_per_step": cached_alg,
                  "traffic_bytes_per_step": 2 * F * 4 * h * h * elem + 8 * kept + 4 * F * h * h,
                  "roofline_frac": cached_alg / (cached_ms * 1e-3) / 1e9 / peak,
                  "l2": f"{n_jobs} independent runs of {per_job >> 20} MiB each stepped round-robin: no step finds its pool, "
                        f"latents or accumulators in the 126 MB L2",
                  "note": "ids unchanged between the steps of a run (SURVEY.md 8d cached-plan regime)"}

    # ---- end to end through the public API with host buffers ------------------------------------------------------
    class _Ctx:
        pass

    class _ED:
        pass

    class _MapSize:
        height = tex
        width = tex

    ids_host = [t.cpu().pin_memory() for t in ids_rot]
    x_host = x_init.cpu().pin_memory()
    x_out = torch.empty_like(x_host).pin_memory()
    ids_dev = torch.empty_like(ids0)
    x_dev = torch.empty_like(x)

    e2e_stream = None
    if not light:
        idm = IDMap(tensor=ids_dev, masks=torch.zeros(1, 1, 1))        # masks are not used by the overlap step
        ed = _ED()
        ed.id_maps, ed.correspond_maps = idm, {(1, 0): _MapSize()}
        ctx = _Ctx()
        ctx.noise, ctx.timestep = x_dev, 900
        oc = OverlapCorresponder(step_finished_inject_ratio=RATIO, process_group=True if world > 1 else None,
                                 exchange=args.exchange, cache_plan=False)   # every e2e step brings new ids
        n_e2e = max(10, min(K, 200)) if id_bytes < (400 << 20) else 10

        def e2e_step(i: int):
            ids_dev.copy_(ids_host[i % n_rot], non_blocking=True)
            x_dev.copy_(x_host, non_blocking=True)
            oc.step_finished(ed, ctx)
            x_out.copy_(x_dev, non_blocking=True)

        for i in range(3):
            e2e_step(i)
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.perf_counter()
        e0.record()
        for i in range(n_e2e):
            e2e_step(i)
        e1.record()
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall
        barrier(world)
        e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), t_wall * 1e3), world) / n_e2e
        e2e_stream = {"value": F_global * h * h * 1e3 / e2e_ms, "unit": "latent-px/s",
                      "h2d_bytes_per_step": id_bytes + x.numel() * elem, "d2h_bytes_per_step": x.numel() * elem,
                      "ms_per_step": e2e_ms, "steps_per_sec": 1e3 / e2e_ms, "steps": n_e2e,
                      "api": "OverlapCorresponder.step_finished(engine_data, sampling_context)",
                      "regime": "every step brings a new id batch over PCIe (PCIe bound)"}

    # The named config: one sampling run = one id batch + 20 denoise steps.  Per run the ids cross PCIe once; every step
    # copies its latents in, calls step_finished (which buckets the ids on the first step and then runs from the cached
    # plan) and copies the latents out.  This is also how the reference arm is timed (keying cached per id batch).
    # Every run is timed on its own (wall clock around a synchronize); the value is the MEDIAN run, p90 beside it.
    e2e_job = None
    if fused_kernel:
        job_steps = 20
        n_runs = 50 if not light else 30
        oc2 = OverlapCorresponder(step_finished_inject_ratio=RATIO, process_group=True if world > 1 else None,
                                  exchange=args.exchange, cache_plan=True)
        # (uploading the next run's ids on a copy stream while this run steps was tried: the id copy and the per-step
        # latent copies then share the DMA engine chunk by chunk and run times become bimodal; sequential it is)
        idm2 = IDMap(tensor=ids_dev, masks=torch.zeros(1, 1, 1))
        ed2 = _ED()
        ed2.id_maps, ed2.correspond_maps = idm2, {(1, 0): _MapSize()}
        ctx2 = _Ctx()
        ctx2.noise, ctx2.timestep, ctx2.total_steps = x_dev, 900, job_steps

        ev_run = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

        def e2e_job_run(j: int):
            ev_run[0].record()
            ids_dev.copy_(ids_host[j % n_rot], non_blocking=True)
            ev_run[1].record()
            idm2.invalidate()
            for s in range(job_steps):
                ctx2.step_index = s
                x_dev.copy_(x_host, non_blocking=True)
                oc2.step_finished(ed2, ctx2)
                x_out.copy_(x_dev, non_blocking=True)
            ev_run[2].record()

        for j in range(2):
            e2e_job_run(j)
        torch.cuda.synchronize()
        run_ms, ids_ms, steps_ms = [], [], []
        for j in range(n_runs):
            barrier(world)
            t_wall = time.perf_counter()
            e2e_job_run(j)
            torch.cuda.synchronize()
            run_ms.append((time.perf_counter() - t_wall) * 1e3)
            ids_ms.append(ev_run[0].elapsed_time(ev_run[1]))
            steps_ms.append(ev_run[1].elapsed_time(ev_run[2]))
        if world > 1:
            tr = torch.tensor(run_ms, dtype=torch.float64, device=dev)
            dist.all_reduce(tr, op=dist.ReduceOp.MAX)
            run_ms = [float(v) for v in tr.cpu()]
        job_ms = statistics.median(run_ms)
        e2e_job = {"value": F_global * h * h * job_steps * 1e3 / job_ms, "unit": "latent-px/s",
                   "steps_per_sec": job_steps * 1e3 / job_ms, "ms_per_step": job_ms / job_steps, "ms_per_run": job_ms,
                   "ms_per_run_p90": percentile(run_ms, 0.9), "ms_per_run_mean": statistics.mean(run_ms),
                   "ms_per_run_min": min(run_ms), "runs": n_runs, "steps": job_steps * n_runs,
                   "ms_ids_h2d_per_run_rank0": statistics.median(ids_ms), "ms_steps_per_run_rank0": statistics.median(steps_ms),
                   "h2d_bytes_per_step": id_bytes // job_steps + x.numel() * elem, "d2h_bytes_per_step": x.numel() * elem,
                   "api": "OverlapCorresponder.step_finished(engine_data, sampling_context)",
                   "regime": f"runs of {job_steps} denoise steps on one id batch: ids H2D once per run ({id_bytes >> 20} MiB, "
                             "amortised above), latents H2D + D2H every step; median over the runs, each timed on the wall "
                             "clock (max over ranks)"}

    out = {
        "metric": "overlap_latent_px_per_sec", "value": F_global * h * h * 1e3 / ms_step, "unit": "latent-px/s",
        "n_gpus": world, "steps": K,
        "warmup": Wm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": DTYPE_NAME[dtype], "data": "synthetic",
        "config": {"workload": workload_string(workload, F_global),
                   "frames_per_gpu": F, "frames_total": F_global, "key_capacity": key_capacity,
                   "launch": launch_mode + (f" ({q} replays of a {m}-step graph" + (f" + one {r_rem}-step graph" if r_rem else "")
                                            + ", no eager launch in the timed region)" if graphs else ""),
                   "fast_path": bool(plan.fast_path),
                   "l2": (f"{n_rot} id buffers of {id_bytes >> 20} MiB rotate, so no step finds its ids in the 126 MB L2" if n_rot > 1
                          else f"the ids of one step ({id_bytes >> 20} MiB per GPU) are larger than the 126 MB L2"),
                   "kernels": "one persistent kernel per step" if fused_kernel else "split reduce / gather kernels",
                   "parallelism": (f"frames sharded over {world} GPU(s), key accumulator ({acc.numel() * 4 >> 10} KiB) "
                                   + (("exchanged inside the step kernel over NVLink: in-switch reduction (multimem.ld_reduce) of the "
                                       "owner's slice + multicast (multimem.st) of the totals" if getattr(plan, "nvls", False) else
                                       "exchanged inside the step kernel over NVLink peer memory (pull reduce-scatter + record all-gather)")
                                      if peer else "summed with one NCCL all-reduce between the reduce and gather kernels"))
                   if world > 1 else "single GPU"},
        "overlap_steps_per_sec": 1e3 / ms_step,
        "id_px_per_sec": F_global * H * H * 1e3 / ms_step,
        "clocks": clocks,
        # `e2e` follows the named config (one id batch per 20 denoise steps — also how the reference arm is timed: keying
        # cached per id batch); `e2e_every_step_new_ids` is the PCIe-bound worst case in which every step uploads new ids
        "e2e": e2e_job if e2e_job is not None else e2e_stream,
        "e2e_every_step_new_ids": e2e_stream,
        "cached_plan": cached,
        "gpu_launches": K if fused_kernel else (2 * K if plan.fast_path else 3 * K),
        "roofline": {"bound": "hbm", "kernel": k1_name,
                     "achieved": k1_gbs, "peak": peak, "unit": "GB/s", "frac": k1_gbs / peak,
                     "traffic": ncu_traffic(workload), "peak_source": peak_src,
                     "bytes_per_launch": k1_bytes, "ms_per_launch": k1_ms, "ms_per_launch_eager_events": k1_ms_eager,
                     "step_bytes_per_gpu": step_bytes,
                     "aggregate": {"bytes_per_step": step_bytes * world if scaling == "weak" else
                                   16 * F_global * H * H + 2 * F_global * 4 * h * h * elem,
                                   "peak": peak * world, "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak},
                     "step_frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak},
    }
    plan.close()
    del ids_rot, ids_host, ids_dev, ids0
    return out


def cpu_overlap_baseline(workload: str, budget_s: float, max_steps: int, frames_cap=None) -> dict:
    """The reference's CPU torch algorithm (oracle/torch_port.py) on this box's host cores, same synthetic workload."""
    from stable_renderer_b200 import synthetic
    from torch_port import CpuOverlapPort
    frames_cfg, H, h, dtype, tex, n_obj, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    F = frames_cfg if frames_cap is None else max(1, min(frames_cfg, frames_cap))
    ids = synthetic.make_ids(F, H, H, tex_h=tex, tex_w=tex, n_obj=n_obj, frac_2048=0.05, seed=1234)
    x = synthetic.make_latents(frames_cfg, 4, h, h, seed=0)[:F].contiguous()
    t0 = time.perf_counter()
    port = CpuOverlapPort(ids)
    t_plan = time.perf_counter() - t0
    port.step(x, RATIO)                                             # warm-up
    times = []
    t_begin = time.perf_counter()
    while len(times) < max_steps and (time.perf_counter() - t_begin < budget_s or len(times) < 3):
        t0 = time.perf_counter()
        x = port.step(x, RATIO)
        times.append(time.perf_counter() - t0)
    t_step = statistics.median(times)
    scale = frames_cfg / F                                          # per-entry work is linear in the frame count
    return {"value": frames_cfg * h * h / (t_step * scale), "unit": "latent-px/s", "steps_per_sec": 1.0 / (t_step * scale),
            "cores": cores, "kind": "port",
            "sample": f"{len(times)} timed steps of {F}/{frames_cfg} frames of {workload} (median {t_step * 1e3:.1f} ms/step"
                      f"{', scaled by frame count' if F != frames_cfg else ''}); keying cached as in the reference "
                      f"(plan build {t_plan * 1e3:.0f} ms, excluded); torch {torch.__version__} CPU ops, {cores} threads",
            "ms_per_step": t_step * scale * 1e3, "plan_ms": t_plan * 1e3}


# ---------------------------------------------------------------------------------------------------------------------
# bake workload (cfg4)
# ---------------------------------------------------------------------------------------------------------------------
def run_bake(args, rank: int, local: int, world: int) -> dict:
    from stable_renderer_b200 import _lib, synthetic
    from stable_renderer_b200.corrmap import CorrespondMap
    _lib.load()
    views_total, H, tex = 64, 1024, 4096
    assert views_total % world == 0
    V = views_total // world
    dev = torch.device("cuda", local)
    ids = synthetic.make_ids(V, H, H, tex_h=tex, tex_w=tex, k=1, frac_2048=0.0, seed=99, device=dev, frame_offset=rank * V)
    colors = torch.rand(V, H, H, 3, device=dev)
    nd = synthetic.make_normal_depth(V, H, H, device=dev, frame_offset=rank * V)
    cm = CorrespondMap(name="bench", k=1, height=tex, width=tex, channel_count=4, device=dev)
    K, Wm = args.steps, args.warmup
    import torch.distributed as dist
    mode = dict(mode="replace", weight_mode=args.bake_weight, normal_depth=nd if args.bake_weight.startswith("view") else None,
                process_group=dist.group.WORLD if world > 1 else None)
    for _ in range(Wm):
        cm.update(colors, ids, **mode)
    barrier(world)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        cm.update(colors, ids, **mode)
    e1.record()
    torch.cuda.synchronize()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world) / K
    # end to end through CorrespondMap.update with pinned host buffers: ids + colours (+ normal/depth) in, written flags out
    host = [t.cpu().pin_memory() for t in (ids, colors)] + ([nd.cpu().pin_memory()] if mode["normal_depth"] is not None else [])
    flags_out = torch.empty_like(cm._writtens, device="cpu").pin_memory()
    n_e2e = 3

    def e2e_step():
        d = [t.to(dev, non_blocking=True) for t in host]
        cm.update(d[1], d[0], **{**mode, "normal_depth": d[2] if len(d) > 2 else None})
        flags_out.copy_(cm._writtens, non_blocking=True)

    e2e_step()
    barrier(world)
    tw = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    torch.cuda.synchronize()
    barrier(world)
    e2e_ms = max_over_ranks((time.perf_counter() - tw) * 1e3, world) / n_e2e
    clocks = sampler.stop()
    peak, peak_src = measured_hbm_peak()
    weighted = args.bake_weight != "none"
    alg = V * H * H * (16 + 12 + (8 if args.bake_weight.startswith("view") else 0)) + 2 * 4 * tex * tex + tex * tex
    return {
        "metric": "bake_views_per_sec", "value": views_total * 1e3 / ms, "unit": "views/s", "n_gpus": world, "steps": K,
        "warmup": Wm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": f"cfg4 bake: {views_total} views {H}x{H} RGB f32 -> {tex}x{tex} fp16 RGBA atlas, weight {args.bake_weight}",
                   "views_per_gpu": V, "l2": "inputs per step (ids+colours) are 1.9 GB/GPU at N=1, far above the 126 MB L2"},
        "texels_per_sec": views_total * H * H * 1e3 / ms, "clocks": clocks,
        "e2e": {"value": views_total * 1e3 / e2e_ms, "unit": "views/s", "ms_per_step": e2e_ms, "steps": n_e2e,
                "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host),
                "d2h_bytes_per_step": flags_out.numel(), "api": "CorrespondMap.update(color_frames, id_maps, ...)"},
        "gpu_launches": 2 * K,
        "roofline": {"bound": "hbm",
                     "kernel": "k_bake_accum_pair + k_bake_finalize (one step)" if weighted else "k_bake_claim_pair + k_bake_write_texels (one step)",
                     "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak,
                     "traffic": ncu_traffic("bake_weighted" if weighted else "bake_none") if world == 1 else None,
                     "peak_source": peak_src, "bytes_per_launch": alg},
    }


# ---------------------------------------------------------------------------------------------------------------------
# legacy overlap workload: ResizeOverlap with the four weighting strategies
# ---------------------------------------------------------------------------------------------------------------------
def run_legacy(args, rank: int, local: int, world: int) -> dict:
    from stable_renderer_b200 import _lib, synthetic
    from stable_renderer_b200.overlap import CorrespondenceMap, ResizeOverlap, Scheduler, overlap_algorithm_factory
    if world > 1:
        raise SystemExit("the legacy overlap keeps per-key group members, not sums: replicas only (SURVEY.md §8e)")
    _lib.load()
    F, H, h = 16, 512, 64
    dev = torch.device("cuda", local)
    ids = [synthetic.make_ids(F, H, H, tex_h=512, tex_w=512, seed=1234 + r, device=dev, legacy_layout=True) for r in range(3)]
    cms = [CorrespondenceMap(t) for t in ids]                 # three id batches of 64 MiB rotate (larger than L2 together)
    frames = [synthetic.make_latents(1, 4, h, h, seed=i).to(dev) for i in range(F)]
    vn = torch.rand(F, H, H, 1, device=dev)
    alpha = Scheduler(interpolate_begin=0.9, interpolate_end=0.9, interpolate_type="constant")
    radius = Scheduler(interpolate_begin=0, interpolate_end=0, interpolate_type="constant")
    K, Wm = max(args.steps // 20, 20), args.warmup
    per = {}
    for name in ("average", "frame_distance", "pixel_distance", "perpendicular_view_normal"):
        ov = ResizeOverlap(alpha, radius, overlap_algorithm_factory(name), verbose=False)
        kw = {"view_normal_map": vn} if name == "perpendicular_view_normal" else {}
        for i in range(Wm):
            ov(frames, cms[i % 3], step=1, timestep=900, **kw)
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            ov(frames, cms[i % 3], step=1, timestep=900, **kw)
        e1.record()
        torch.cuda.synchronize()
        per[name] = e0.elapsed_time(e1) / K
    peak, peak_src = measured_hbm_peak()
    alg = 16 * F * H * H + 2 * F * 4 * h * h * 4
    ms = per["average"]
    return {"metric": "legacy_overlap_latent_px_per_sec", "value": F * h * h * 1e3 / ms, "unit": "latent-px/s", "n_gpus": 1,
            "steps": K, "warmup": Wm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"legacy ResizeOverlap.__call__: {F} frames of {H}x{H}x4 int32 (obj, mat, texX, texY) ids -> "
                                   f"{h}x{h}x4 f32 latents, alpha 0.9, radius 0; value = strategy 'average'",
                       "l2": "three id batches of 64 MiB rotate"},
            "ms_per_step_by_strategy": per, "e2e": None, "gpu_launches": 3 * K,
            "roofline": {"bound": "hbm", "kernel": "k_legacy_seed + k_legacy_accum + k_legacy_finalize (whole call, incl. torch.stack)",
                         "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak,
                         "traffic": None, "peak_source": peak_src, "bytes_per_launch": alg}}


def cpu_legacy_baseline() -> dict:
    """oracle restatement of the reference's per-trace Python loop (overlap.py:83-152) on a 2-frame 128x128 crop."""
    import numpy as np
    import srx_oracle as O
    from stable_renderer_b200 import synthetic
    F, H, h = 2, 128, 16
    ids = synthetic.make_ids(F, H, H, tex_h=128, tex_w=128, seed=1234, legacy_layout=True).numpy()
    x = np.random.default_rng(0).standard_normal((F, 1, 4, h, h))
    t0 = time.perf_counter()
    O.legacy_resize_overlap(x, ids, 0.9, "average")
    dt = time.perf_counter() - t0
    return {"value": F * h * h / dt, "unit": "latent-px/s", "cores": 1, "kind": "port",
            "sample": f"{F} frames of {H}x{H} ids -> {h}x{h} latents, strategy average, {dt * 1e3:.0f} ms (the reference loops over "
                      "every key in Python; work is proportional to the id pixels)"}


def cpu_bake_baseline(views: int = 4) -> dict:
    """The reference's CPU bake (oracle/torch_port.py::cpu_bake_port = CorrespondMap.update, mode 'replace') on a few views
    of the cfg4 shape; the reference has no weighted bake, so this is the closest CPU counterpart."""
    from stable_renderer_b200 import synthetic
    from torch_port import cpu_bake_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    H, tex = 1024, 4096
    ids = synthetic.make_ids(views, H, H, tex_h=tex, tex_w=tex, k=1, frac_2048=0.0, seed=99)
    colors = torch.rand(views, H, H, 3)
    masks = ((ids[..., 2] == 2048) | (ids == 0).all(-1)).float()
    values = torch.zeros(1, tex * tex, 4, dtype=torch.float16)
    writtens = torch.zeros(1, tex * tex, dtype=torch.bool)
    cpu_bake_port(values, writtens, colors[:1], ids[:1], masks[:1])
    t0 = time.perf_counter()
    cpu_bake_port(values, writtens, colors, ids, masks)
    dt = time.perf_counter() - t0
    return {"value": views / dt, "unit": "views/s", "cores": cores, "kind": "port",
            "sample": f"{views} of 64 views of cfg4 through CorrespondMap.update semantics (mode replace, unweighted), {dt * 1e3:.0f} ms"}


# ---------------------------------------------------------------------------------------------------------------------
def run_feature_overlap(local: int) -> dict:
    """SURVEY.md 8f-4 (the body of post_atten_inject): 16 frames of 512^2 ids, the three attention sizes of an SD1.5 UNet at 512^2,
    fp16 features.  Per size: one call with cached buckets (every layer / denoise step after the first of an id batch) and one that
    buckets the ids as well; algorithmic bytes = (rows gathered + 3 B h w) rows of c * 2 bytes (+ 16 B per id pixel when bucketing)."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.corrmap import IDMap
    from stable_renderer_b200.feature import feature_overlap
    dev = torch.device("cuda", local)
    F, H = 16, 512
    ids = synthetic.make_ids(F, H, H, tex_h=512, tex_w=512, n_obj=1, frac_2048=0.05, seed=7, device=dev)
    idm = IDMap(tensor=ids, frame_indices=list(range(F)))
    peak, _ = measured_hbm_peak()

    def timed(fn, reps=20, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    sizes = {}
    for h, c in ((64, 320), (32, 640), (16, 1280)):
        x = torch.randn(F, h * h, c, device=dev).to(torch.float16)
        kw = dict(map_size=(H, H), key_capacity=512 * 512, check=False)
        info = {}
        feature_overlap(x, idm, 0.6, info=info, **kw)
        row = c * 2
        apply_bytes = (info["rows_gathered"] + 3 * F * h * h) * row
        id_bytes = ids.numel() * ids.element_size()
        ms_cached = timed(lambda: feature_overlap(x, idm, 0.6, **kw))
        ms_bucket = timed(lambda: feature_overlap(x, idm, 0.6, cache_buckets=False, **kw))
        sizes[f"hw{h}x{h}_c{c}"] = {
            "ms_per_call": ms_cached, "algorithmic_bytes": apply_bytes, "rows_gathered": info["rows_gathered"],
            "achieved_gbs": apply_bytes / ms_cached / 1e6, "roofline_frac": apply_bytes / ms_cached / 1e6 / peak,
            "ms_per_call_bucketing_every_call": ms_bucket, "algorithmic_bytes_bucketing": apply_bytes + id_bytes}
    return {"workload": f"feature overlap: {F} frames of {H}x{H} ids, [B, h*w, c] fp16 features, ratio 0.6, up-sampling size = id map size",
            "api": "feature.feature_overlap(origin_values, id_map, ratio, map_size)  (OverlapCorresponder.post_atten_inject body)",
            "timing": "CUDA events around the public call (host side included), median of 20", "dtype": "f16", "sizes": sizes}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=list(WORKLOADS) + ["bake", "legacy"], default="cfg3")
    ap.add_argument("--bake-weight", default="view_normal_depth", choices=["none", "uniform", "view_normal", "view_normal_depth"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the secondary records (extra.cfg2, extra.cfg4_bake) of the default run")
    ap.add_argument("--exchange", choices=["auto", "peer", "nccl"], default="auto",
                    help="multi-GPU accumulator exchange: inside the step kernel over NVLink peer memory, or NCCL all-reduce")
    ap.add_argument("--split-kernels", action="store_true", help="run the split reduce / gather kernels instead of the persistent one")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        # the reference's own CPU path for this metric/config; rank 0 alone runs it
        if int(os.environ.get("RANK", "0")) != 0:
            return
        wl = "cfg3" if args.workload in ("bake", "legacy") else args.workload
        frames_cfg, H, h, dtype, tex, n_obj, scaling = WORKLOADS[wl]
        # bounded sample: enough frames per step that K + W steps end within ~2 minutes
        probe = cpu_overlap_baseline(wl, budget_s=2.0, max_steps=3, frames_cap=2)
        per_frame_s = probe["ms_per_step"] / 1e3 / frames_cfg
        frames_cap = int(max(1, min(frames_cfg, 120.0 / max(args.steps + args.warmup, 1) / max(per_frame_s, 1e-9))))
        res = cpu_overlap_baseline(wl, budget_s=1e9, max_steps=args.steps, frames_cap=frames_cap)
        world = max(args.gpus, 1)
        frames_total = frames_cfg * world if scaling == "weak" else frames_cfg
        value = res["value"]        # latent-px/s of the CPU path does not depend on how many frames one step holds
        steps_per_sec = res["steps_per_sec"] * (frames_cfg / frames_total)
        line = {"impl": "reference", "metric": "overlap_latent_px_per_sec", "value": value, "unit": "latent-px/s",
                "overlap_steps_per_sec": steps_per_sec,
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / steps_per_sec,
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_string(wl, frames_total),
                           "arm": "reference CPU torch path (oracle/torch_port.py: the reference's torch ops, keying cached per id "
                                  "batch as the reference does; float32 arithmetic on the same latents)"},
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": value, "unit": "latent-px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        line["cpu_baseline"]["value"] = value
        print(json.dumps(line), flush=True)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU baseline")
    rank, local, world = init_dist(args.gpus)
    t_start = time.perf_counter()
    parity = None
    if args.workload == "bake":
        out = run_bake(args, rank, local, world)
    elif args.workload == "legacy":
        out = run_legacy(args, rank, local, world)
    else:
        # N-rank parity step first (outside every timed region): a wrong answer must not produce a bench line
        parity = parity_check(rank, local, world, args.exchange, WORKLOADS[args.workload][3])
        out = run_overlap(args, args.workload, rank, local, world)
        out["parity_check"] = parity
        if not args.no_extras:
            extra = {}
            torch.cuda.empty_cache()
            if args.workload != "cfg2":
                a2 = argparse.Namespace(**vars(args))
                a2.steps, a2.warmup = max(60, min(args.steps, 600)), max(args.warmup, 5)
                r2 = run_overlap(a2, "cfg2", rank, local, world, light=True)
                extra["cfg2"] = {k: r2[k] for k in ("value", "unit", "ms_per_step", "steps", "scaling", "overlap_steps_per_sec",
                                                    "config", "roofline", "cached_plan", "e2e", "dtype")}
                torch.cuda.empty_cache()
            if world == 1:
                a4 = argparse.Namespace(**vars(args))
                a4.steps, a4.warmup = max(10, min(args.steps, 50)), 3
                for wname, key in (("view_normal_depth", "cfg4_bake"), ("none", "cfg4_bake_reference_modes")):
                    a4.bake_weight = wname
                    r4 = run_bake(a4, rank, local, world)
                    extra[key] = {k: r4[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "config", "roofline", "e2e",
                                                     "texels_per_sec", "dtype")}
                    torch.cuda.empty_cache()
                extra["feature_overlap"] = run_feature_overlap(local)
                torch.cuda.empty_cache()
            out["extra"] = extra
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline and args.workload == "bake":
            out["cpu_baseline"] = cpu_bake_baseline()
        if world == 1 and not args.no_cpu_baseline and args.workload == "legacy":
            out["cpu_baseline"] = cpu_legacy_baseline()
        if world == 1 and not args.no_cpu_baseline and args.workload not in ("bake", "legacy"):
            out["cpu_baseline"] = cpu_overlap_baseline(args.workload, budget_s=12.0, max_steps=40,
                                                       frames_cap=16 if args.workload in ("cfg3", "cfg5") else None)
        out["bench_wall_s"] = time.perf_counter() - t_start
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        log(f"[bench] PARITY CHECK FAILED: {parity}")
        sys.exit(3)


if __name__ == "__main__":
    main()
