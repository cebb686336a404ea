"""world_size-2 gloo test of the sharded-verification host logic (CPU only): contiguous candidate split,
all-gather of the 48-byte best records, identical deterministic winner on every rank."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from srrg2_laser_slam_2d_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from srrg2_laser_slam_2d_b200._abi import BEST_DTYPE
    from srrg2_laser_slam_2d_b200.sharding import all_gather_best, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    # a fake scored candidate table, identical on every rank; each rank reduces its own shard
    rng = np.random.default_rng(0)
    n_cand = 101
    inl = rng.integers(300, 600, n_cand)
    inl[[17, 83]] = 700                                    # a tie across the two shards
    chi = rng.uniform(1, 5, n_cand).astype(np.float32)
    chi[[17, 83]] = 2.0
    lo, hi = shard_range(n_cand, rank, world)
    k = lo + int(np.lexsort((np.arange(lo, hi), chi[lo:hi] / inl[lo:hi], -inl[lo:hi]))[0])
    rec = np.zeros(1, BEST_DTYPE)[0]
    rec["candidate"], rec["guess"], rec["n_inliers"], rec["n_corr"], rec["chi_inliers"] = k, 0, inl[k], inl[k], chi[k]
    best = all_gather_best(rec)
    q.put((rank, int(best["candidate"]), int(best["n_inliers"])))
    dist.destroy_process_group()


def test_two_rank_all_gather_picks_the_same_winner():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [(0, 17, 700), (1, 17, 700)]             # tie broken by the lowest candidate id on both ranks


def test_shard_groups_balances_pairs_and_never_splits_a_group():
    from srrg2_laser_slam_2d_b200.sharding import shard_groups
    rng = np.random.default_rng(3)
    for n_groups in (1, 5, 64, 1000):
        sizes = rng.integers(0, 20, n_groups)
        off = np.concatenate([[0], np.cumsum(sizes)])
        for world in (1, 2, 3, 8):
            spans = [shard_groups(off, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n_groups
            assert all(a[1] == b[0] and a[0] <= a[1] for a, b in zip(spans, spans[1:]))
            pairs = [off[hi] - off[lo] for lo, hi in spans]
            assert sum(pairs) == off[-1]
            if n_groups >= 64:
                assert max(pairs) <= off[-1] / world + sizes.max()      # within one group of the ideal share


def _group_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from srrg2_laser_slam_2d_b200._abi import BEST_DTYPE
    from srrg2_laser_slam_2d_b200.sharding import all_gather_group_best, shard_groups
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    sizes = rng.integers(0, 6, 40)
    off = np.concatenate([[0], np.cumsum(sizes)])
    inl = rng.integers(200, 400, off[-1])
    lo, hi = shard_groups(off, rank, world)
    local = np.zeros(40, BEST_DTYPE)
    local["candidate"], local["guess"] = -1, -1
    for g in range(lo, hi):                               # this rank's groups: best = most inliers, accepted if >= 300
        if off[g + 1] > off[g]:
            k = off[g] + int(np.argmax(inl[off[g]:off[g + 1]]))
            if inl[k] >= 300:
                local[g]["candidate"], local[g]["guess"] = 1000 + k, k - off[g]
                local[g]["n_inliers"], local[g]["n_corr"], local[g]["chi_inliers"] = inl[k], inl[k], 1.0
    best = all_gather_group_best(local)
    q.put((rank, best["candidate"].tolist()))
    dist.destroy_process_group()


def test_two_rank_group_best_equals_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_group_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(1)
    sizes = rng.integers(0, 6, 40)
    off = np.concatenate([[0], np.cumsum(sizes)])
    inl = rng.integers(200, 400, off[-1])
    want = []
    for g in range(40):
        seg = inl[off[g]:off[g + 1]]
        want.append(1000 + off[g] + int(np.argmax(seg)) if len(seg) and seg.max() >= 300 else -1)
    assert got[0][1] == want and got[1][1] == want
