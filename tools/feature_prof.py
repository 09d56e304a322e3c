"""A few cached-bucket feature overlap calls for ncu: python tools/feature_prof.py <h> <c> <f16|f32>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from stable_renderer_b200 import synthetic  # noqa: E402
from stable_renderer_b200.corrmap import IDMap  # noqa: E402
from stable_renderer_b200.feature import feature_overlap  # noqa: E402

h, c = int(sys.argv[1]), int(sys.argv[2])
dt = torch.float16 if sys.argv[3] == "f16" else torch.float32
F, H = 16, 512
dev = torch.device("cuda", 0)
ids = synthetic.make_ids(F, H, H, tex_h=512, tex_w=512, n_obj=1, frac_2048=0.05, seed=7, device=dev)
idm = IDMap(tensor=ids, frame_indices=list(range(F)))
x = torch.randn(F, h * h, c, device=dev).to(dt)
for _ in range(4):
    out = feature_overlap(x, idm, 0.6, map_size=(H, H), key_capacity=512 * 512, check=False)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
