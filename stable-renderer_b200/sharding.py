"""Host-side arithmetic of the frame-sharded (multi-GPU) overlap step.

The reference is single-GPU; these helpers define the partition the kernels assume:
  * frames are split into contiguous blocks, one block per rank (`frame_shard`);
  * the dense key table is split into `world` owner slices of ceil(K / world) slots (`owner_slice`, `owner_of`) — the
    slice whose partial sums a rank pulls from its peers and whose totals it serves (csrc/srx_fused.cu, phase X);
  * a key matters for the cached plan if it wins a cell on ANY rank (`union_need_maps`)."""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np


def frame_shard(frames_total: int, rank: int, world: int) -> Tuple[int, int]:
    """(first frame, frame count) of `rank`: contiguous blocks, the first `frames_total % world` ranks get one more."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(int(frames_total), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def owner_slice(key_capacity: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) slots owned by `rank`: slices of ceil(K / world) slots, the last ones clipped to K."""
    per = -(-int(key_capacity) // int(world))
    begin = min(rank * per, key_capacity)
    return begin, min(begin + per, key_capacity)


def owner_of(slot: int, key_capacity: int, world: int) -> int:
    return int(slot) // (-(-int(key_capacity) // int(world)))


def union_need_maps(maps: Sequence[np.ndarray]) -> np.ndarray:
    """What the MAX all-reduce of the ranks' winner-key byte maps computes."""
    out = np.zeros_like(maps[0])
    for m in maps:
        out = np.maximum(out, m)
    return out
