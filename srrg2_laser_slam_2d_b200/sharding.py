"""Multi-GPU split of loop-closure verification.

Candidates are independent, so they are split contiguously over the ranks (all guesses of a candidate
stay on one rank: its cloud is read once); the single query cloud and the parameters are replicated.
The only exchange is one all-gather of each rank's 32-byte best record (ls2d_best) followed by the same
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
    out = torch.zeros(8 * world, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    return reduce_best(np.frombuffer(out.cpu().numpy().tobytes(), dtype=BEST_DTYPE))
