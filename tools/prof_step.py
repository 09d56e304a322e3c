"""Runs a few eager overlap steps (no CUDA graph) for ncu / compute-sanitizer: python tools/prof_step.py cfg2 [steps]
SRX_PROF_CACHED=1 runs the bucketing pass + cached-plan steps instead of streaming steps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import RATIO, WORKLOADS  # noqa: E402
from stable_renderer_b200 import synthetic  # noqa: E402
from stable_renderer_b200.plan import OverlapPlan  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
frames, H, h, dtype, tex, n_obj, _ = WORKLOADS[wl]
if len(sys.argv) > 3:
    frames = int(sys.argv[3])
dev = torch.device("cuda", 0)
ids = synthetic.make_ids(frames, H, H, tex_h=tex, tex_w=tex, n_obj=n_obj, frac_2048=0.05, seed=1234, device=dev)
ids2 = torch.roll(ids, 1, 0).contiguous()
x = synthetic.make_latents(frames, 4, h, h, seed=0, dtype=dtype).to(dev)
plan = OverlapPlan(None, x.shape, id_shape=ids.shape, id_dtype=ids.dtype, key_capacity=tex * tex, device=dev)
if os.environ.get("SRX_PROF_CACHED"):
    plan.build_cache(ids)
    for i in range(steps):
        plan.step(x, RATIO, cached=True)
else:
    for i in range(steps):
        plan.step(x, RATIO, ids=ids if i % 2 == 0 else ids2)
torch.cuda.synchronize()
plan.check()
print("ok", wl, frames, float(x.float().abs().mean()))
