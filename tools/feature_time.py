"""Times the wide-channel feature overlap and the cell-similarity overlap on one GPU:
    python tools/feature_time.py [frames] [id_size]
Shapes follow the attention layers the hook sees on an SD1.5 UNet at 512x512: (hw, c) = (64*64, 320), (32*32, 640), (16*16, 1280).
Algorithmic bytes per call = the ids once + the features read once + the result written once."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from stable_renderer_b200 import synthetic  # noqa: E402
from stable_renderer_b200.corrmap import IDMap  # noqa: E402
from stable_renderer_b200.feature import feature_overlap, taichi_cells_overlap  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda", 0)
ids = synthetic.make_ids(F, H, H, tex_h=512, tex_w=512, n_obj=1, frac_2048=0.05, seed=7, device=dev)
idm = IDMap(tensor=ids, frame_indices=list(range(F)))


def timed(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for h, c in ((64, 320), (32, 640), (16, 1280)):
    for dt in (torch.float16, torch.float32):
        g = torch.Generator(device="cpu").manual_seed(h * c)
        x = torch.randn(F, h * h, c, generator=g).to(dev, dt)
        kw = dict(map_size=(H, H), key_capacity=512 * 512, check=False)
        info = {}
        feature_overlap(x, idm, 0.6, info=info, **kw)
        row = c * x.element_size()
        apply_bytes = (info["rows_gathered"] + 3 * F * h * h) * row
        id_bytes = ids.numel() * ids.element_size()
        med_b, best_b = timed(lambda: feature_overlap(x, idm, 0.6, cache_buckets=False, **kw))
        med, best = timed(lambda: feature_overlap(x, idm, 0.6, **kw))
        print(f"feature_overlap F={F} ids {H}x{H} hw={h}x{h} c={c} {str(dt)[6:]}: rows gathered {info['rows_gathered']} "
              f"({info['rows_gathered'] * row / 1e6:.0f} MB through L2) | buckets cached: median {med * 1e3:.1f} us, min {best * 1e3:.1f} us, "
              f"{apply_bytes / 1e6:.0f} MB algorithmic -> {apply_bytes / med / 1e6:.0f} GB/s | bucketing every call: median "
              f"{med_b * 1e3:.1f} us ({(apply_bytes + id_bytes) / 1e6:.0f} MB -> {(apply_bytes + id_bytes) / med_b / 1e6:.0f} GB/s)")

# cell-similarity overlap: b frames of a 64x64 id crop (4096 pixels), 8x8 cells, c = 320
b, P, cells, c = 16, 64 * 64, 64, 320
idf = ids[:, :64, :64, :].reshape(F, -1, 4)[:b].contiguous()
contrib = torch.rand(b, P, device=dev)
vals = torch.randn(b, cells, c, device=dev)
new = torch.zeros_like(vals)


def cells_call():
    new.zero_()
    taichi_cells_overlap(idf, vals, new, contrib)


med, best = timed(cells_call, reps=20, warm=3)
print(f"taichi_cells_overlap b={b} pixels={P} cells={cells} c={c}: median {med * 1e3:.1f} us, min {best * 1e3:.1f} us "
      f"(the reference's loop visits {(b * P) ** 2:.2e} pixel pairs)")
