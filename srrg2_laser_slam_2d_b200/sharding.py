"""Multi-GPU split of loop-closure verification.

Candidates are independent, so they are split contiguously over the ranks (all guesses of a candidate
stay on one rank: its cloud is read once); the single query cloud and the parameters are replicated.
The only exchange is one all-gather of each rank's 48-byte best record (ls2d_best) followed by the same
deterministic best-of on every rank, so 1-GPU and N-GPU answers are identical."""
from __future__ import annotations

import numpy as np

from ._abi import BEST_DTYPE, reduce_best


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of rank's share; the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_best(local_best: np.ndarray, group=None) -> np.ndarray:
    """All-gather the ranks' ls2d_best records through torch.distributed (NCCL on GPUs, gloo in the CPU
    tests) and return the global winner.  `local_best` is one BEST_DTYPE record."""
    import torch
    import torch.distributed as dist

    rec = np.zeros(1, BEST_DTYPE)
    rec[0] = local_best
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(rec.view(np.int32).copy()).to(dev)
    out = torch.zeros(mine.numel() * world, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    return reduce_best(np.frombuffer(out.cpu().numpy().tobytes(), dtype=BEST_DTYPE))


def shard_groups(group_offsets: np.ndarray, rank: int, world: int) -> tuple[int, int]:
    """All-pairs search (BASELINE.json configs[4]): the candidate-pair list is grouped by query (fixed) map and the
    ranks own contiguous runs of groups, cut so that every rank gets about the same number of PAIRS (a group is never
    split: its fixed map stays on one rank).  Returns the rank's [g_lo, g_hi)."""
    off = np.asarray(group_offsets, np.int64)
    n_groups, total = len(off) - 1, int(off[-1])
    cuts = [0]
    for r in range(1, world):
        # first group boundary at or after r/world of the pairs, never before the previous cut
        g = int(np.searchsorted(off, (total * r + world - 1) // world, side="left"))
        cuts.append(min(max(g, cuts[-1]), n_groups))
    cuts.append(n_groups)
    return cuts[rank], cuts[rank + 1]


def combine_group_best(gathered: np.ndarray) -> np.ndarray:
    """gathered: [world, n_groups] ls2d_best records, candidate = -1 where a rank does not own the group or accepted
    nothing.  Groups are owned by exactly one rank, so the combination is a per-group reduce_best."""
    gathered = np.asarray(gathered)
    world, n_groups = gathered.shape
    out = np.zeros(n_groups, BEST_DTYPE)
    out["candidate"], out["guess"] = -1, -1
    for g in range(n_groups):
        col = np.ascontiguousarray(gathered[:, g])
        if (col["candidate"] >= 0).any():
            out[g] = reduce_best(col)
    return out


def all_gather_group_best(local: np.ndarray, group=None) -> np.ndarray:
    """All-gather every rank's full-length per-group record array ([n_groups] BEST_DTYPE, -1 outside its own run)
    and combine.  20,000 groups x 48 B = 960 KB per rank."""
    import torch
    import torch.distributed as dist

    local = np.ascontiguousarray(local, dtype=BEST_DTYPE)
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(local.view(np.int32).copy()).to(dev)
    out = torch.zeros(world * mine.numel(), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    rec = np.frombuffer(out.cpu().numpy().tobytes(), dtype=BEST_DTYPE).reshape(world, len(local))
    return combine_group_best(rec)
