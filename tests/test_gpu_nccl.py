"""ls2d_verify_sharded_nccl (include/ls2d.h) with a REAL ncclComm_t: the C-ABI collective of the sharded loop-closure
verification (BASELINE.json: "NCCL over NVLink used only to all-gather per-shard best-candidate poses and chi2").

One process per GPU; every rank verifies its contiguous shard of the candidates, the 48-byte ls2d_best records are
all-gathered over the caller's communicator and reduced with the deterministic rule, so every rank -- and the 1-GPU
ls2d_verify over all candidates -- names the same winner, bit for bit.  The 2-rank test needs two GPUs (skipped on a
1-GPU box; `gpurun --gpus 2` runs it: profiles/r02_nccl_2gpu.log); the 1-rank test exercises the same entry point
(communicator of size 1) on any GPU box."""
import os

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

N_CAND, N_GUESS = 96, 4
KW = dict(canvas_cols=1081, point_distance=1.414, normal_cos=0.8, cauchy_chi_threshold=0.05, max_iterations=30)


def _workload():
    from srrg2_laser_slam_2d_b200.synthetic import make_scan_pairs
    sp = make_scan_pairs(N_CAND, seed=91, motion_xy=0.3, motion_theta=0.15)   # candidate 0 is the query's true match
    rng = np.random.default_rng(8)
    guesses = (sp.gt_xyt[0][None, None, :] + rng.uniform(-0.1, 0.1, (N_CAND, N_GUESS, 3))).astype(np.float32)
    return sp, guesses


def _sharded(rank, world, comm):
    import torch
    from srrg2_laser_slam_2d_b200 import Gates, Handle, default_params
    from srrg2_laser_slam_2d_b200._abi import LS2D_FIXED, LS2D_MOVING
    from srrg2_laser_slam_2d_b200.sharding import shard_range
    sp, guesses = _workload()
    h = Handle(rank, default_params(**KW))
    h.upload_clouds(LS2D_FIXED, sp.fixed_pts, sp.fixed_off)
    h.upload_clouds(LS2D_MOVING, sp.moving_pts, sp.moving_off)
    gates = Gates(300, 0.1, 0.8)
    lo, hi = shard_range(N_CAND, rank, world)
    dev = torch.device("cuda", rank)
    cand = torch.arange(lo, hi, dtype=torch.int32, device=dev)
    gs = torch.from_numpy(guesses[lo:hi].copy()).to(dev)
    torch.cuda.synchronize(dev)
    best = h.verify_sharded_nccl(0, cand.data_ptr(), hi - lo, gs.data_ptr(), N_GUESS, gates, lo, comm.ptr, world)
    single = h.verify(0, None, guesses, gates)         # all candidates on this GPU alone
    bad_rc = None
    try:
        h.verify_sharded_nccl(0, cand.data_ptr(), hi - lo, gs.data_ptr(), N_GUESS, gates, lo, comm.ptr, world + 1)
    except Exception as e:                              # n_ranks must equal the communicator's size
        bad_rc = str(e)
    h.close()
    return best.tobytes(), single.tobytes(), int(best["candidate"]), bad_rc


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from srrg2_laser_slam_2d_b200.nccl_comm import NcclComm
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    comm = NcclComm(rank, world, rank)
    out = _sharded(rank, world, comm)
    comm.close()
    q.put((rank,) + out)
    dist.destroy_process_group()


def test_sharded_verify_over_a_single_rank_communicator():
    from srrg2_laser_slam_2d_b200.nccl_comm import NcclComm
    comm = NcclComm(0, 1, 0)
    best, single, cand, bad = _sharded(0, 1, comm)
    comm.close()
    assert best == single and cand == 0
    assert bad is not None and "invalid argument" in bad


def test_sharded_verify_two_ranks_names_the_single_gpu_winner():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, best, single, cand, bad in got:
        assert best == single and cand == 0, rank       # every rank: the winner of the 1-GPU run, bit for bit
        assert bad is not None and "invalid argument" in bad
    assert got[0][1] == got[1][1]
