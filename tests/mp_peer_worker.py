"""Worker of tests/test_gpu_fused_step.py::test_fused_peer_exchange_two_gpus (launched with torch.distributed.run):
frames sharded over the ranks, accumulator exchange inside the step kernel over NVLink peer memory; every rank checks
its own frames against the numpy oracle run over ALL frames."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import srx_oracle as O  # noqa: E402
from stable_renderer_b200 import synthetic  # noqa: E402
from stable_renderer_b200.plan import OverlapPlan  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    exchange = os.environ.get("SRX_TEST_EXCHANGE", "peer")
    F, H, tex, steps = 4 * world, 256, 256, 3
    ids = synthetic.make_ids(F, H, H, tex_h=tex, tex_w=tex, frac_2048=0.05, seed=31)
    x0 = synthetic.make_latents(F, 4, H // 8, H // 8, seed=2)
    want = x0.numpy()
    for _ in range(steps):
        want = O.overlap_step(want, ids.numpy(), None, ratio=0.5, accumulate="f64")
    sl = slice(rank * F // world, (rank + 1) * F // world)
    x = x0[sl].contiguous().cuda()
    plan = OverlapPlan(ids[sl].contiguous().cuda(), x.shape, key_capacity=tex * tex, process_group=dist.group.WORLD,
                       exchange=exchange)
    assert plan.exchange == exchange, f"exchange {plan.exchange!r}, wanted {exchange!r}"
    for _ in range(steps):
        if plan.exchange == "peer":
            plan.step(x, 0.5)
        else:
            plan.reduce(x)
            dist.all_reduce(plan.accumulator, op=dist.ReduceOp.SUM)
            plan.gather(x, 0.5)
    plan.check()
    torch.cuda.synchronize()
    got = x.cpu().numpy()
    err = np.abs(got - want[sl])
    ok = bool((err <= 1e-5 + 5e-5 * np.abs(want[sl])).all())
    if plan.exchange == "peer":
        # cached-plan regime on the same plan: bucket the ids once (the ranks' winner-key maps are combined with a MAX
        # all-reduce), then run the same number of steps from the pool
        x2 = x0[sl].contiguous().cuda()
        plan.build_cache()
        for _ in range(steps):
            plan.step(x2, 0.5, cached=True)
        plan.check()
        torch.cuda.synchronize()
        err2 = np.abs(x2.cpu().numpy() - want[sl])
        ok = ok and bool((err2 <= 1e-5 + 5e-5 * np.abs(want[sl])).all())
        err = np.maximum(err, err2)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"max abs err rank0 {err.max():.3e}; exchange={plan.exchange} nvls={getattr(plan, 'nvls', False)}", flush=True)
        print("PEER_OK" if int(flag.item()) == 1 else "PEER_MISMATCH", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
