"""Oracle self-consistency of the multi-slice aligner (decisions D14-D17): several laser slices with their own
sensor_in_robot plus the odometry prior in one 3x3 system (MULTI.json:700-730; LASER_0.json:502-506).  CPU only."""
import math

import numpy as np

from srrg2_laser_slam_2d_b200.synthetic import make_multi_sensor_pairs, make_scan_pairs

SENSORS = ((0.2, 0.05, 0.1), (-0.2, 0.0, math.pi))


def slices_for(oracle, msp, **kw):
    """MULTI.json tracking aligner: laser_0 with Cauchy 0.01 / normal_cos 0.9, laser_1 without robustifier / 0.8"""
    base = dict(canvas_cols=721, max_iterations=10, min_num_correspondences=5, with_sensor=1)
    base.update(kw)
    s0 = oracle.default_params(normal_cos=0.9, cauchy_chi_threshold=0.01, sensor_in_robot=msp.sensors[0], **base)
    s1 = oracle.default_params(normal_cos=0.8, cauchy_chi_threshold=-1.0, sensor_in_robot=msp.sensors[1], **base)
    return [s0, s1]


def sets_of(msp):
    fixed = [(msp.fixed_pts[s], msp.fixed_off[s]) for s in range(len(msp.fixed_pts))]
    moving = [(msp.moving_pts, msp.moving_off)] * len(msp.fixed_pts)
    return fixed, moving


def test_one_slice_without_prior_is_the_single_slice_aligner(oracle):
    sp = make_scan_pairs(6, n_beams=721, seed=21)
    prm = oracle.default_params(canvas_cols=721, normal_cos=0.9)
    for mode in (oracle.SUM_SEQUENTIAL, oracle.SUM_TREE):
        a, ai = oracle.align_batch(prm, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, sp.init_xyt,
                                   sum_mode=mode, tree_threads=256)
        b, bi = oracle.align_multi_batch([prm], [(sp.fixed_pts, sp.fixed_off)], [(sp.moving_pts, sp.moving_off)],
                                         sp.init_xyt, sum_mode=mode, tree_threads=256)
        assert a.tobytes() == b.tobytes() and ai.tobytes() == bi.tobytes()


def test_prior_jacobian_matches_finite_differences(oracle):
    rng = np.random.default_rng(3)
    for _ in range(20):
        z = rng.uniform(-1, 1, 3)
        X = rng.uniform(-1, 1, 3)
        pr = oracle.make_prior(np.eye(3), z=z)
        e0, J = oracle.prior_error_and_jacobian(pr, X)
        T = lambda v: np.array([[math.cos(v[2]), -math.sin(v[2]), v[0]], [math.sin(v[2]), math.cos(v[2]), v[1]],
                                [0, 0, 1]])
        for k in range(3):
            d = np.zeros(3)
            d[k] = 1e-3
            Xp = T(X) @ T(d)                         # X <- X * v2t(dx)  (VariableSE2Right)
            P = np.linalg.inv(T(z)) @ Xp
            ep = np.array([P[0, 2], P[1, 2], math.atan2(P[1, 0], P[0, 0])])
            num = (ep - e0) / 1e-3
            num[2] = (math.remainder(ep[2] - e0[2], 2 * math.pi)) / 1e-3
            assert np.allclose(num, J[:, k], atol=2e-3)
    # at X == Z the error vanishes
    e, _ = oracle.prior_error_and_jacobian(oracle.make_prior(np.eye(3), z=(0.3, -0.2, 0.5)), (0.3, -0.2, 0.5))
    assert np.abs(e).max() < 1e-6


def test_two_sensor_slices_converge_to_ground_truth(oracle):
    msp = make_multi_sensor_pairs(6, sensors=SENSORS, seed=11)
    fixed, moving = sets_of(msp)
    res, its = oracle.align_multi_batch(slices_for(oracle, msp), fixed, moving, msp.init_xyt)
    assert (res["status"] == 0).all()
    err = np.abs(np.stack([res["x"], res["y"], res["theta"]], 1) - msp.gt_xyt)
    assert err[:, :2].max() < 5e-3 and err[:, 2].max() < 3e-3
    # both slices contribute: more correspondences than either alone
    one, _ = oracle.align_multi_batch(slices_for(oracle, msp)[:1], fixed[:1], moving[:1], msp.init_xyt)
    assert (res["n_corr"] > one["n_corr"]).all()


def test_slice_below_min_correspondences_is_skipped(oracle):
    msp = make_multi_sensor_pairs(3, sensors=SENSORS, seed=12)
    fixed, moving = sets_of(msp)
    sl = slices_for(oracle, msp)
    both, _ = oracle.align_multi_batch(sl, fixed, moving, msp.init_xyt)
    sl[1].min_num_correspondences = 100000          # laser_1 can never contribute (D14)
    skipped, _ = oracle.align_multi_batch(sl, fixed, moving, msp.init_xyt)
    only0, _ = oracle.align_multi_batch(sl[:1], fixed[:1], moving[:1], msp.init_xyt)
    assert skipped.tobytes() == only0.tobytes() and skipped.tobytes() != both.tobytes()
    sl[0].min_num_correspondences = 100000          # nobody contributes => NotEnoughCorrespondences
    none, _ = oracle.align_multi_batch(sl, fixed, moving, msp.init_xyt)
    assert (none["status"] == 1).all() and (none["iterations"] == 0).all() and (none["n_corr"] > 0).all()


def test_prior_pulls_the_estimate_and_counts_as_one_inlier(oracle):
    msp = make_multi_sensor_pairs(4, sensors=SENSORS, seed=13)
    fixed, moving = sets_of(msp)
    sl = slices_for(oracle, msp)
    free, _ = oracle.align_multi_batch(sl, fixed, moving, msp.init_xyt)
    weak = oracle.make_prior(np.diag([1.0, 1.0, 1.0]))
    res, _ = oracle.align_multi_batch(sl, fixed, moving, msp.init_xyt, prior=weak, prior_z=msp.odom_xyt)
    assert (res["status"] == 0).all()
    # stats: the prior is one more inlier factor, never a correspondence (D15)
    assert np.array_equal(res["n_inliers"] + res["n_kernelized"], res["n_corr"] + 1)
    assert np.abs(np.stack([res["x"], res["y"], res["theta"]], 1) - msp.gt_xyt).max() < 6e-3
    # an overwhelming prior pins the estimate on its measurement
    z = msp.gt_xyt + np.float32(0.03)
    strong = oracle.make_prior(np.diag([1e9, 1e9, 1e9]))
    pinned, _ = oracle.align_multi_batch(sl, fixed, moving, msp.init_xyt, prior=strong, prior_z=z)
    assert np.abs(np.stack([pinned["x"], pinned["y"], pinned["theta"]], 1) - z).max() < 1e-4
    assert np.abs(np.stack([free["x"], free["y"], free["theta"]], 1) - z).max() > 1e-2


def test_prior_full_information_matrix_and_robustifier(oracle):
    msp = make_multi_sensor_pairs(2, sensors=SENSORS, seed=14)
    fixed, moving = sets_of(msp)
    sl = slices_for(oracle, msp, max_iterations=1)
    A = np.array([[4.0, 0.5, 0.2], [0.5, 3.0, -0.1], [0.2, -0.1, 2.0]], np.float32)
    z = msp.gt_xyt + np.float32(0.1)
    base, _ = oracle.align_multi_batch(sl, fixed, moving, msp.init_xyt)
    res, _ = oracle.align_multi_batch(sl, fixed, moving, msp.init_xyt, prior=oracle.make_prior(A), prior_z=z)
    # one iteration at X = identity: H grows by J^T A J with J = blockdiag(R(z)^T, 1), chi by e^T A e
    for p in range(2):
        c, s = math.cos(z[p, 2]), math.sin(z[p, 2])
        Zi = np.linalg.inv(np.array([[c, -s, z[p, 0]], [s, c, z[p, 1]], [0, 0, 1]]))
        e = np.array([Zi[0, 2], Zi[1, 2], -z[p, 2]])
        J = np.eye(3)
        J[:2, :2] = Zi[:2, :2]
        dH = J.T @ A.astype(np.float64) @ J
        got = res["H"][p] - base["H"][p]
        assert np.allclose(got, dH[np.triu_indices(3)], rtol=2e-3, atol=2e-3)
        assert abs((res["chi_inliers"][p] - base["chi_inliers"][p]) - e @ A @ e) < 1e-3
    # with a Cauchy kernel on the prior slice the factor is kernelized (chi = e^T A e >> tau)
    rob, _ = oracle.align_multi_batch(sl, fixed, moving, msp.init_xyt,
                                      prior=oracle.make_prior(A, cauchy_chi_threshold=0.01), prior_z=z)
    assert np.array_equal(rob["n_kernelized"], base["n_kernelized"] + 1)
    assert np.array_equal(rob["n_inliers"], base["n_inliers"])
