"""Worker of tests/test_gpu_bake.py::test_view_sharded_bake_two_gpus (launched with torch.distributed.run): every rank bakes
its own block of views through CorrespondMap.update(process_group=...) — reference modes (claims MAX-reduced, partial atlases
SUM-reduced) and the weighted bake (sums all-reduced) — and compares with a single-GPU bake of ALL views on the same device."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from stable_renderer_b200 import synthetic  # noqa: E402
from stable_renderer_b200.corrmap import CorrespondMap  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    F, H, tex, k = 3 * world + 1, 64, 48, 2                     # ragged: the last rank holds one view more
    per = [3] * world
    per[-1] += 1
    lo = sum(per[:rank])
    hi = lo + per[rank]
    ok = True
    for rnd in range(2):
        ids = synthetic.make_ids(F, H, H, tex_h=tex, tex_w=tex, k=k, frac_2048=0.0, seed=50 + rnd, frame_offset=2 * rnd).cuda()
        colors = synthetic.make_colors(F, H, H, 3, seed=51 + rnd).cuda()
        nd = synthetic.make_normal_depth(F, H, H, seed=52 + rnd).cuda()
        for mode, wm in (("replace", "none"), ("first", "none"), ("replace", "view_normal_depth")):
            key = (mode, wm)
            if rnd == 0:
                maps.setdefault(key, (CorrespondMap(name="ref", k=k, height=tex, width=tex, channel_count=4),
                                      CorrespondMap(name="shard", k=k, height=tex, width=tex, channel_count=4)))
            ref, cm = maps[key]
            kw = dict(mode=mode, ignore_obj_mat_id=True, weight_mode=wm)
            ref.update(colors, ids, normal_depth=nd if wm != "none" else None, **kw)
            cm.update(colors[lo:hi], ids[lo:hi], normal_depth=nd[lo:hi] if wm != "none" else None,
                      process_group=dist.group.WORLD, **kw)
            torch.cuda.synchronize()
            same_w = torch.equal(cm._writtens, ref._writtens)
            if wm == "none":
                same_v = torch.equal(cm._values.view(torch.int16), ref._values.view(torch.int16))
            else:
                same_v = torch.allclose(cm._values.float(), ref._values.float(), rtol=2e-3, atol=2e-3)
            if not (same_w and same_v):
                ok = False
                print(f"rank {rank}: mismatch in round {rnd} for {key}: written {same_w} values {same_v}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("BAKE_SHARD_OK" if int(flag.item()) == 1 else "BAKE_SHARD_FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


maps: dict = {}

if __name__ == "__main__":
    main()
