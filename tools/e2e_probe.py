"""Where does the end-to-end time of a cached sampling run go?  python tools/e2e_probe.py"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from stable_renderer_b200 import synthetic  # noqa: E402
from stable_renderer_b200.corresponder import OverlapCorresponder  # noqa: E402
from stable_renderer_b200.corrmap import IDMap  # noqa: E402

dev = torch.device("cuda", 0)
F, H, h, tex = 32, 512, 64, 512
ids = synthetic.make_ids(F, H, H, tex_h=tex, tex_w=tex, frac_2048=0.05, seed=1234, device=dev)
x_host = synthetic.make_latents(F, 4, h, h, seed=0).pin_memory()
x_out = torch.empty_like(x_host).pin_memory()
x_dev = x_host.to(dev)


class Ctx:
    noise, timestep, total_steps, step_index = x_dev, 900, 20, 5


class ED:
    id_maps = IDMap(tensor=ids, masks=torch.zeros(1, 1, 1))

    class _M:
        height = tex
        width = tex
    correspond_maps = {(1, 0): _M()}


oc = OverlapCorresponder(step_finished_inject_ratio=0.5)
ed, ctx = ED(), Ctx()
for _ in range(3):
    oc.step_finished(ed, ctx)
torch.cuda.synchronize()


def timed(label, fn, n=300):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    t_host = (time.perf_counter() - t) / n * 1e6
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t) / n * 1e6
    print(f"{label:50s} host {t_host:8.1f} us/iter   with sync {t_all:8.1f} us/iter", flush=True)


timed("step_finished only (cached plan)", lambda: oc.step_finished(ed, ctx))
timed("H2D latents only", lambda: x_dev.copy_(x_host, non_blocking=True))
timed("D2H latents only", lambda: x_out.copy_(x_dev, non_blocking=True))


def full():
    x_dev.copy_(x_host, non_blocking=True)
    oc.step_finished(ed, ctx)
    x_out.copy_(x_dev, non_blocking=True)


timed("H2D + step_finished + D2H", full)
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    oc.step_finished(ed, ctx)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)

plan = next(iter(ED.id_maps._plans.values()))
timed("build_cache (bucketing pass)", lambda: plan.build_cache(), n=30)
ids_host = ids.cpu().pin_memory()
timed("H2D ids (128 MiB)", lambda: ids.copy_(ids_host, non_blocking=True), n=10)


def job():
    ids.copy_(ids_host, non_blocking=True)
    ED.id_maps.invalidate()
    for s in range(20):
        ctx.step_index = s
        full()


timed("whole run: ids H2D + 20 x (H2D, step, D2H)", job, n=10)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    job()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
