"""ls2d_score_batch -- one linearisation per pair without a pose update -- against the oracle's first iteration.

Clouds of up to 1152 points run score_kernel (persistent CTAs, bulk-async copies of the next pair's clouds behind an
mbarrier, csrc/ls2d_score.cuh); the batches here are larger than the resident grid (148 SMs x 2 CTAs), so every CTA
walks several pairs through both stages; ragged and empty clouds, z-buffer ties, sensor offsets, id indirection and
both accumulation arithmetics are covered.  Counts, chi2 and H bit for bit."""
import numpy as np
import pytest

import golden_util as gu
from srrg2_laser_slam_2d_b200 import default_params
from srrg2_laser_slam_2d_b200._abi import LS2D_FIXED, LS2D_MOVING, score_reduction_threads
from srrg2_laser_slam_2d_b200.synthetic import FAR_POINT, make_scan_pairs



gpu = pytest.mark.gpu


def check(s, o, poses):
    for f in ("n_corr", "n_inliers", "n_kernelized"):
        assert np.array_equal(s[f], o[f]), f
    solved = o["status"] != 3       # the oracle's first iteration also SOLVES (status 3 = singular system); scoring does not
    assert np.array_equal(s["status"][solved], o["status"][solved])
    assert np.array_equal(gu.bits(s["chi_inliers"]), gu.bits(o["chi_inliers"]))
    assert np.array_equal(gu.bits(s["H"]), gu.bits(o["H"]))
    ulp = np.abs(gu.bits(s["chi_kernelized"]).astype(np.int64) - gu.bits(o["chi_kernelized"]).astype(np.int64))
    assert ulp.max(initial=0) <= 4 * max(1, int(s["n_kernelized"].max(initial=1)))
    assert np.array_equal(s["x"], poses[:, 0]) and np.array_equal(s["y"], poses[:, 1])   # pose untouched
    assert ((s["iterations"] == 1) | (s["status"] == 1)).all()


@gpu
@pytest.mark.parametrize("n_beams,cols,n_pairs,extra", [
    (721, 721, 900, {}), (1081, 1081, 700, {}), (361, 361, 1500, {}),
    (1081, 1081, 320, dict(with_sensor=1, sensor_in_robot=(0.2, -0.1, 0.15))),
    (1081, 1081, 320, dict(single_rounding_accumulation=1)),
    (1152, 1151, 64, {}),                      # the largest cloud / widest canvas score_kernel stages
    (1500, 1081, 24, {}), (1081, 1200, 24, {}), (1081, 1081, 24, dict(factor=1)),   # shapes the aligner kernels score
])
def test_score_batch_is_the_first_linearisation(handle_factory, oracle, n_beams, cols, n_pairs, extra):
    sp = make_scan_pairs(n_pairs, n_beams=n_beams, seed=12 + n_beams, chunk=256)
    kw = dict(canvas_cols=cols, normal_cos=0.9)
    kw.update(extra)
    gp = default_params(**kw)
    h = handle_factory(gp)
    h.upload_clouds(LS2D_FIXED, sp.fixed_pts, sp.fixed_off)
    h.upload_clouds(LS2D_MOVING, sp.moving_pts, sp.moving_off)
    poses = (sp.gt_xyt + np.float32([0.01, -0.01, 0.005])).astype(np.float32)
    s = h.score_batch(poses)
    prm = oracle.default_params(max_iterations=1, **kw)
    o, _ = oracle.align_batch(prm, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, poses,
                              sum_mode=oracle.SUM_TREE, tree_threads=score_reduction_threads(n_beams, params=gp),
                              n_threads=oracle.max_threads())
    check(s, o, poses)
    assert (s["n_corr"] > 100).mean() > 0.9


@gpu
def test_score_ragged_empty_and_tied_clouds_with_id_indirection(handle_factory, oracle):
    sp = make_scan_pairs(40, n_beams=721, seed=99)
    rng = np.random.default_rng(4)
    f_clouds = [sp.fixed_pts[sp.fixed_off[p]:sp.fixed_off[p + 1]] for p in range(40)]
    m_clouds = [sp.moving_pts[sp.moving_off[p]:sp.moving_off[p + 1]] for p in range(40)]
    for p in range(40):                                   # ragged: every cloud its own length, some empty
        f_clouds[p] = f_clouds[p][:rng.integers(0, 722)] if p % 5 else f_clouds[p][:0]
        m_clouds[p] = m_clouds[p][:rng.integers(1, 722)] if p % 7 else m_clouds[p][:0]
    f_clouds[3] = np.concatenate([f_clouds[3][:200], f_clouds[3][:200]])          # every winner tied: lowest index wins
    m_clouds[3] = np.concatenate([m_clouds[4][:150], m_clouds[4][:150], m_clouds[4][:150]])
    m_clouds[6] = np.tile(np.float32(FAR_POINT), (300, 1))                        # nothing projects
    pack = lambda cl: (np.concatenate(cl).astype(np.float32), np.concatenate([[0], np.cumsum([len(c) for c in cl])]).astype(np.int32))
    fp, fo = pack(f_clouds)
    mp, mo = pack(m_clouds)
    kw = dict(canvas_cols=721, normal_cos=0.9, min_num_correspondences=3)
    gp = default_params(**kw)
    h = handle_factory(gp)
    h.upload_clouds(LS2D_FIXED, fp, fo)
    h.upload_clouds(LS2D_MOVING, mp, mo)
    n = 1200                                              # > resident grid: the persistent loop and both stages
    fid = rng.integers(0, 40, n).astype(np.int32)
    mid = rng.integers(0, 40, n).astype(np.int32)
    mid[:40], fid[:40] = np.arange(40), np.arange(40)
    poses = rng.uniform(-0.05, 0.05, (n, 3)).astype(np.float32)
    s = h.score_batch(poses, fid, mid)
    prm = oracle.default_params(max_iterations=1, **kw)
    o, _ = oracle.align_batch(prm, fp, fo, mp, mo, poses, fid, mid, sum_mode=oracle.SUM_TREE,
                              tree_threads=score_reduction_threads(721, params=gp), n_threads=oracle.max_threads())
    check(s, o, poses)
    assert (s["status"] == 1).any() and (s["status"] == 0).any()


@gpu
def test_score_of_a_pose_equals_the_aligners_first_iteration_statistics(handle_factory):
    """the scoring pass and the aligner's first round see the same correspondences: counts equal, chi2 close (their
    reduction shapes differ)"""
    sp = make_scan_pairs(64, n_beams=1081, seed=7)
    h = handle_factory(default_params(canvas_cols=1081, normal_cos=0.9))
    h.upload_clouds(LS2D_FIXED, sp.fixed_pts, sp.fixed_off)
    h.upload_clouds(LS2D_MOVING, sp.moving_pts, sp.moving_off)
    s = h.score_batch(sp.init_xyt)
    _, it = h.align_batch(sp.init_xyt, want_iters=True)
    assert np.array_equal(s["n_corr"], it["n_corr"][:, 0]) and np.array_equal(s["n_inliers"], it["n_inliers"][:, 0])
    assert np.allclose(s["chi_inliers"], it["chi_inliers"][:, 0], rtol=1e-4)


@gpu
def test_gated_square_root_is_the_correctly_rounded_one(handle_factory):
    """every binary32 operand a squared-range gate can let through (range_min >= 1 mm ... range_max <= 1e5 m):
    the branch-free square root of the kernels == __fsqrt_rn, bit for bit"""
    h = handle_factory()
    n, bad = h.selftest_gated_sqrt(1e-7, 1.1e10)
    assert n > 4e8 and bad == 0, (n, bad)


def test_squared_range_gate_is_the_reference_gate(oracle):
    """CPU: lo / hi of make_range_gate2 are exact -- checked through the host build of the header"""
    import ctypes as C
    from srrg2_laser_slam_2d_b200._abi import MATHCHECK_PATH
    L = C.CDLL(MATHCHECK_PATH)
    L.ls2d_host_range_gate2.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    for rmin, rmax in [(0.3, 20.0), (0.01, 30.0), (0.0, 1000.0), (0.1, 5.6), (1.0, 1e5), (0.29999998, 19.999998)]:
        lo, hi = C.c_float(), C.c_float()
        L.ls2d_host_range_gate2(rmin, rmax, C.byref(lo), C.byref(hi))
        rmin32, rmax32 = np.float32(rmin), np.float32(rmax)
        for t, inside in ((lo.value, True), (np.nextafter(np.float32(lo.value), np.float32(-1)), False)):
            if t >= 0:
                assert (np.sqrt(np.float32(t)) >= rmin32) == inside
        assert np.sqrt(np.float32(hi.value)) <= rmax32 < np.sqrt(np.nextafter(np.float32(hi.value), np.float32(np.inf)))


@gpu
def test_lock_free_cell_hand_back_is_deterministic(handle_factory):
    """score_kernel and icp_multi2_kernel let the winner of a column hand the z-buffer cells back inside the interval
    in which the losers still test them (losers only ever see the winner's rho or EMPTY; racecheck reports these
    accesses, profiles/r02/sanitizer.md).  40 repetitions on 4096 pairs, with ties in the clouds, must be one result."""
    from srrg2_laser_slam_2d_b200._abi import multi_reduction_threads
    sp = make_scan_pairs(4096, n_beams=721, seed=404, chunk=512)
    fixed = sp.fixed_pts.copy()
    fixed[1::7] = fixed[0::7][:len(fixed[1::7])]                 # duplicated points: equal rho in a column (decision D3)
    gp = default_params(canvas_cols=721, normal_cos=0.8, max_iterations=4)
    h = handle_factory(gp)
    h.upload_clouds(LS2D_FIXED, fixed, sp.fixed_off)
    h.upload_clouds(LS2D_MOVING, sp.moving_pts, sp.moving_off)
    h.upload_clouds(2, sp.fixed_pts, sp.fixed_off)
    poses = (sp.gt_xyt + np.float32([0.01, -0.01, 0.005])).astype(np.float32)
    first_s = h.score_batch(poses).tobytes()
    sl = [default_params(canvas_cols=721, normal_cos=0.8, max_iterations=4, with_sensor=1, sensor_in_robot=(0.1, 0.0, 0.0)),
          default_params(canvas_cols=721, normal_cos=0.8, max_iterations=4, with_sensor=1, sensor_in_robot=(-0.1, 0.0, 0.1))]
    assert multi_reduction_threads(sl, 721, 721, True) == 256 | 1 << 16 | 1 << 17   # icp_multi2_kernel serves this shape
    first_m = h.align_multi(sl, [LS2D_FIXED, 2], [LS2D_MOVING, LS2D_MOVING], poses).tobytes()
    for _ in range(40):
        assert h.score_batch(poses).tobytes() == first_s
        assert h.align_multi(sl, [LS2D_FIXED, 2], [LS2D_MOVING, LS2D_MOVING], poses).tobytes() == first_m
