"""Phase breakdown of the persistent step kernel from its in-kernel SM-clock timestamps.
    python tools/fz_trace.py cfg2            (one GPU)
    python -m torch.distributed.run --nproc-per-node N ... tools/fz_trace.py cfg2   (frame-sharded, peer exchange)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import WORKLOADS, shard  # noqa: E402
from stable_renderer_b200 import synthetic  # noqa: E402
from stable_renderer_b200.plan import OverlapPlan  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
frames_cfg, H, h, dtype, tex, n_obj, scaling = WORKLOADS[wl]
if os.environ.get('SRX_TRACE_FRAMES'):
    frames_cfg = int(os.environ['SRX_TRACE_FRAMES'])
F, f0, F_global = shard(frames_cfg, scaling, rank, world)
dev = torch.device("cuda", local)
ids = [synthetic.make_ids(F, H, H, tex_h=tex, tex_w=tex, n_obj=n_obj, frac_2048=0.05, seed=1234, device=dev, frame_offset=f0)]
ids.append(torch.roll(ids[0], 1, 0).contiguous())
ids.append(torch.roll(ids[0], 2, 0).contiguous())
x = synthetic.make_latents(F, 4, h, h, seed=0, dtype=dtype).to(dev)
plan = OverlapPlan(None, x.shape, id_shape=ids[0].shape, id_dtype=ids[0].dtype, key_capacity=tex * tex, device=dev,
                   process_group=dist.group.WORLD if (world > 1 and not os.environ.get("SRX_TRACE_NOPEER")) else None)
CACHED = bool(os.environ.get("SRX_TRACE_CACHED"))
if CACHED:
    plan.build_cache(ids[0])
names = ["A stream", "barrier0", "X pull+reduce", "signal", "B gather loop", "stats reduce", "C adain"]
acc = [[0.0] * 7 for _ in range(2)]
n = 30
mhz = 1965.0
ev_ms = 0.0
for i in range(n + 5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    e0.record()
    plan.step(x, 0.5, ids=ids[i % 3], cached=CACHED)
    e1.record()
    torch.cuda.synchronize()
    if i >= 5:
        ev_ms += e0.elapsed_time(e1) / n
    if i >= 5:
        tr = plan.read_trace()
        for w, key in enumerate(("first_cta", "last_cta")):
            t = tr[key]
            if world == 1:
                t = t[:3] + [t[2], t[2]] + t[5:]      # no exchange phase: stamps 3, 4 are not written
            if t[5] < t[4] or t[6] < t[5]:
                t = t[:5] + [t[7], t[7], t[7]]      # idle CTA in the gather phase
            for j in range(7):
                acc[w][j] += (t[j + 1] - t[j]) / mhz / n
# back-to-back eager launches without host syncs in between
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(60):
    plan.step(x, 0.5, ids=ids[i % 3], cached=CACHED)
e1.record()
torch.cuda.synchronize()
b2b = e0.elapsed_time(e1) / 60 * 1e3
ring = [r for r in plan.read_step_ring() if r[1] > 0 and r[0] != (1 << 64) - 1]
ring.sort()
spans = [(e - s) / 1e3 for s, e, _ in ring]
ss_parts = [0.0] * 7
for _, _, st in ring[-16:]:
    if world == 1:
        st = st[:3] + [st[2], st[2]] + st[5:]
    for j in range(7):
        ss_parts[j] += (st[j + 1] - st[j]) / mhz / 16
ss_txt = ", ".join(f"{nm} {v:.1f}" for nm, v in zip(names, ss_parts))
gaps = [(ring[i + 1][0] - ring[i][1]) / 1e3 for i in range(len(ring) - 1)]
tr = plan.read_trace()["first_cta"]
last_span = (tr[7] - tr[0]) / mhz
if world == 1:
    tr = tr[:3] + [tr[2], tr[2]] + tr[5:]
trl = tr
b2b_parts = ", ".join(f"{nm} {(trl[j + 1] - trl[j]) / mhz:.1f}" for j, nm in enumerate(names))
if world > 1:
    dist.barrier()
for r in range(world):
    if r == rank:
        for w, key in enumerate(("first CTA", "last CTA")):
            print(f"[rank {rank}] {wl} {key}: " + ", ".join(f"{nm} {v:.1f}" for nm, v in zip(names, acc[w])) +
                  f" | total {sum(acc[w]):.1f} us (exchange={plan.exchange})", flush=True)
        print(f"[rank {rank}] global-timer: kernel spans {[round(v, 1) for v in spans[-8:]]} us, gaps between kernels {[round(v, 1) for v in gaps[-8:]]} us; steady-state first-CTA phases: {ss_txt}", flush=True)
        print(f"[rank {rank}] event-timed single launch {ev_ms * 1e3:.1f} us; 60 back-to-back launches {b2b:.1f} us/step "
              f"(in-kernel span of the last one {last_span:.1f} us: {b2b_parts})", flush=True)
    if world > 1:
        dist.barrier()
if world > 1:
    dist.destroy_process_group()
