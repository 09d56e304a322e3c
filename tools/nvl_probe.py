"""NVLink peer-access probe: run with torch.distributed.run --nproc-per-node 2 tools/nvl_probe.py"""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "nvlprobe", "libnvlprobe.so"))
lib.nvl_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
MAXB = 64 << 20
buf = symm_mem.empty(MAXB, dtype=torch.uint8, device=f"cuda:{local}")
hdl = symm_mem.rendezvous(buf, group=dist.group.WORLD)
buf.zero_()
peer = int(hdl.buffer_ptrs[(rank + 1) % world])
tout = torch.zeros(2 * 148, dtype=torch.int64, device="cuda")
names = {0: "contiguous 16B loads", 4: "contiguous 32B loads", 1: "random 32B gathers", 2: "contiguous 16B stores + fence.sys",
         3: "contiguous 32B stores + fence.sys"}
torch.cuda.synchronize()
dist.barrier()
for size in (64 << 10, 1 << 20, 4 << 20, 16 << 20):
    for mode in (0, 4, 1, 2, 3):
        res = []
        for it in range(6):
            dist.barrier()
            if rank == 0:      # only rank 0 drives traffic: unidirectional numbers
                lib.nvl_probe(peer, peer, size, mode, 148, 512, tout.data_ptr(), torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                t = tout.view(148, 2).cpu()
                res.append((int(t[:, 1].max()) - int(t[:, 0].min())) / 1e3)
            torch.cuda.synchronize()
        if rank == 0:
            best = min(res[1:])
            print(f"{size >> 10:6d} KiB  {names[mode]:36s} {best:8.1f} us  {size / best / 1e3:8.1f} GB/s", flush=True)
# both directions at once (each rank loads from / stores to the other)
for size in (4 << 20,):
    for mode in (0, 2, 3):
        res = []
        for it in range(6):
            dist.barrier()
            lib.nvl_probe(peer, peer, size, mode, 148, 512, tout.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            t = tout.view(148, 2).cpu()
            res.append((int(t[:, 1].max()) - int(t[:, 0].min())) / 1e3)
        if rank == 0:
            best = min(res[1:])
            print(f"{size >> 10:6d} KiB  BIDIR {names[mode]:30s} {best:8.1f} us  {size / best / 1e3:8.1f} GB/s per direction", flush=True)
dist.barrier()
dist.destroy_process_group()
