"""Times the persistent step kernel per launch (CUDA events): python tools/fz_time.py cfg2 [iters]
SRX_FZ_DEBUG bits isolate parts (1 = no reductions, 2 = drain only, 4 = phase A only); SRX_SPLIT=1 times the split kernels."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from stable_renderer_b200 import synthetic  # noqa: E402
from stable_renderer_b200.plan import OverlapPlan  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
frames, H, h, dtype, tex, n_obj, _ = WORKLOADS[wl]
if len(sys.argv) > 3:
    frames = int(sys.argv[3])
dev = torch.device("cuda", 0)
ids = [synthetic.make_ids(frames, H, H, tex_h=tex, tex_w=tex, n_obj=n_obj, frac_2048=0.05, seed=1234, device=dev)]
ids.append(torch.roll(ids[0], 1, 0).contiguous())
ids.append(torch.roll(ids[0], 2, 0).contiguous())
x = synthetic.make_latents(frames, 4, h, h, seed=0, dtype=dtype).to(dev)
plan = OverlapPlan(None, x.shape, id_shape=ids[0].shape, id_dtype=ids[0].dtype, key_capacity=tex * tex, device=dev,
                   split_kernels=bool(os.environ.get("SRX_SPLIT")))
for i in range(5):
    plan.step(x, 0.5, ids=ids[i % 3])
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
for i, (a, b) in enumerate(evs):
    a.record()
    plan.step(x, 0.5, ids=ids[i % 3])
    b.record()
torch.cuda.synchronize()
t = [a.elapsed_time(b) for a, b in evs]
nbytes = ids[0].numel() * 4 + 2 * x.numel() * x.element_size()
med = statistics.median(t)
print(f"{wl} F={frames} fused={plan.fused} DEBUG={os.environ.get('SRX_FZ_DEBUG', '0')}: median {med * 1e3:.1f} us, "
      f"min {min(t) * 1e3:.1f} us, {nbytes / med / 1e6:.0f} GB/s (step bytes)", flush=True)
