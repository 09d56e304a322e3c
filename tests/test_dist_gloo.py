"""CPU tests (gloo, world_size 2) of the frame-sharded overlap step's host-side logic: the partition helpers, and the
exchange protocol itself — owner slices, sums in rank order, winner-key union for the cached plan — restated with numpy
and run across two real processes, against the single-process oracle over all frames."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import srx_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_shard_and_owner_slices_partition_everything():
    from stable_renderer_b200.sharding import frame_shard, owner_of, owner_slice
    for total in (1, 7, 32, 96, 768):
        for world in (1, 2, 3, 4, 8):
            spans = [frame_shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    for cap in (64, 1000, 262144, (1 << 24) + 256):
        for world in (1, 2, 5, 8):
            sl = [owner_slice(cap, r, world) for r in range(world)]
            assert sl[0][0] == 0 and sl[-1][1] == cap
            for (a0, a1), (b0, _) in zip(sl, sl[1:]):
                assert a1 == b0
            for slot in (0, cap // 3, cap - 1):
                o = owner_of(slot, cap, world)
                assert sl[o][0] <= slot < sl[o][1]
    with pytest.raises(ValueError):
        frame_shard(4, 2, 2)


def _partials(x, ids, K):
    """One rank's phase A: per-key sums / counts of its own frames, per-cell winner key (numpy restatement)."""
    B, C, h, w = x.shape
    vsi = O.vertex_screen_info(ids, None)
    sx, sy, fr = O.entry_cells(vsi, h, w)
    key = vsi[:, 3].astype(np.int64)
    vals = x[fr, :, sy, sx].astype(np.float64)
    sums = np.zeros((K, C))
    cnt = np.zeros(K)
    np.add.at(sums, key, vals)
    np.add.at(cnt, key, 1.0)
    cell = fr.astype(np.int64) * (h * w) + sy.astype(np.int64) * w + sx
    last = np.full(B * h * w, -1, dtype=np.int64)
    np.maximum.at(last, cell, np.arange(cell.size))
    winner = np.where(last >= 0, key[np.maximum(last, 0)], -1)
    return sums, cnt, winner


def _finish(x, sums, cnt, winner, ratio):
    B, C, h, w = x.shape
    x32 = x.astype(np.float32)
    flat = np.ascontiguousarray(x32.transpose(0, 2, 3, 1).reshape(-1, C))
    touched = np.nonzero(winner >= 0)[0]
    mean = (sums[winner[touched]] / cnt[winner[touched], None]).astype(np.float32)
    bl = flat.copy()
    bl[touched] = np.float32(1 - ratio) * flat[touched] + np.float32(ratio) * mean
    blended = bl.reshape(B, h, w, C).transpose(0, 3, 1, 2)
    return O.adain(x32, blended)


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.sharding import frame_shard, owner_slice
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        F, H, tex, ratio = 6, 64, 32, 0.5
        K = tex * tex
        ids = synthetic.make_ids(F, H, H, tex_h=tex, tex_w=tex, frac_2048=0.05, seed=5).numpy()
        x = synthetic.make_latents(F, 4, H // 8, H // 8, seed=6).numpy()
        want = O.overlap_step(x, ids, None, ratio=ratio, accumulate="f64")
        f0, fc = frame_shard(F, rank, world)
        xs, idss = x[f0:f0 + fc], ids[f0:f0 + fc]
        sums, cnt, winner = _partials(xs, idss, K)
        # exchange as in csrc/srx_fused.cu phase X: every rank can read every peer's partials (all_gather stands in for
        # NVLink peer loads); the owner adds its slice in rank order; totals are read back from the owners
        part = torch.from_numpy(np.concatenate([sums, cnt[:, None]], axis=1))
        parts = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(parts, part)
        b, e = owner_slice(K, rank, world)
        mine = torch.zeros(e - b, part.shape[1], dtype=part.dtype)
        for p in range(world):
            mine += parts[p][b:e]
        per = owner_slice(K, 0, world)[1]
        padded = torch.zeros(per, part.shape[1], dtype=part.dtype)
        padded[:e - b] = mine
        slices = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(slices, padded)
        total = torch.cat(slices)[:K].numpy()
        got = _finish(xs, total[:, :4], total[:, 4], winner, ratio)
        err = np.abs(got - want[f0:f0 + fc])
        assert (err <= 3e-6 + 1e-5 * np.abs(want[f0:f0 + fc])).all(), f"rank {rank}: max abs err {err.max():.3e}"

        # cached plan: pairs are kept only for keys that win a cell.  Filtering with the rank's OWN winners loses the
        # contributions to keys that win on another rank only; the MAX all-reduce of the byte maps fixes that.
        need_local = np.zeros(K, dtype=np.uint8)
        need_local[winner[winner >= 0]] = 1
        need = torch.from_numpy(need_local.copy())
        dist.all_reduce(need, op=dist.ReduceOp.MAX)
        need = need.numpy()
        kept = part.numpy() * need[:, None]
        keptl = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(keptl, torch.from_numpy(kept))
        total_kept = sum(t.numpy() for t in keptl)
        got2 = _finish(xs, total_kept[:, :4], np.maximum(total_kept[:, 4], 1e-30), winner, ratio)
        assert np.abs(got2 - got).max() < 1e-6, "union of winner maps must reproduce the unfiltered result"
        lost = int(((need == 1) & (need_local == 0) & (cnt > 0)).sum())
        cnts = torch.tensor([lost])
        dist.all_reduce(cnts)
        assert int(cnts.item()) > 0, "the test input must contain keys that win on the other rank only"
        open(os.path.join(tmp, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_frame_sharded_exchange_two_processes_gloo(tmp_path):
    port = 29650 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
