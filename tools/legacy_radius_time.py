"""Times one legacy ResizeOverlap call with kernel_radius 1 (the ordered sweep, csrc/srx_legacy_ordered.cu) on config-1-like
sizes: 16 frames of 512x512 tuple-keyed ids -> 64x64x4 latents.  python tools/legacy_radius_time.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stable_renderer_b200 import synthetic  # noqa: E402
from stable_renderer_b200.overlap import CorrespondenceMap, ResizeOverlap, Scheduler, overlap_algorithm_factory  # noqa: E402

T, H, h = 16, 512, 64
ids = synthetic.make_ids(T, H, H, tex_h=512, tex_w=512, seed=1235, legacy_layout=True, dtype=torch.int32, device="cuda")
cmap = CorrespondenceMap.from_ids(ids)
frames = [synthetic.make_latents(1, 4, h, h, seed=i, device="cuda") for i in range(T)]
for radius in (0.0, 1.0):
    a_s = Scheduler(interpolate_begin=0.9, interpolate_end=0.9, interpolate_type="constant")
    r_s = Scheduler(interpolate_begin=radius, interpolate_end=radius, interpolate_type="constant")
    for strategy in ("average", "frame_distance"):
        ov = ResizeOverlap(a_s, r_s, overlap_algorithm_factory(strategy), verbose=False)
        ov(frames, cmap, step=0, timestep=500)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ov(frames, cmap, step=0, timestep=500)
        torch.cuda.synchronize()
        print(f"radius {int(radius)} {strategy}: {(time.perf_counter() - t0) * 1e3:.2f} ms per call, {len(cmap)} keys", flush=True)
