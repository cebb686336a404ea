"""Oracle self-consistency: analytic cases the restatement must satisfy.
CPU only.  Each test names the reference behaviour it pins."""
import math

import numpy as np
import pytest

from srrg2_laser_slam_2d_b200.synthetic import make_scan_pairs, reference_demo_scene

FLT_MAX = np.finfo(np.float32).max


def wall_cloud(n=200, x=3.0, half=2.0):
    """points on the wall x = const facing the origin"""
    y = np.linspace(-half, half, n, dtype=np.float32)
    return np.stack([np.full(n, x, np.float32), y, -np.ones(n, np.float32), np.zeros(n, np.float32)], -1)


def corner_cloud(n=400):
    """two perpendicular walls (x = 3 and y = 2) seen from the origin: constrains x, y and theta"""
    a = wall_cloud(n // 2, 3.0, 2.0)
    x = np.linspace(-2.0, 3.0, n // 2, dtype=np.float32)
    b = np.stack([x, np.full(n // 2, 2.0, np.float32), np.zeros(n // 2, np.float32), -np.ones(n // 2, np.float32)], -1)
    return np.concatenate([a, b])


def transform(cloud, xyt):
    c, s = math.cos(xyt[2]), math.sin(xyt[2])
    out = cloud.astype(np.float64).copy()
    out[:, 0] = c * cloud[:, 0] - s * cloud[:, 1] + xyt[0]
    out[:, 1] = s * cloud[:, 0] + c * cloud[:, 1] + xyt[1]
    out[:, 2] = c * cloud[:, 2] - s * cloud[:, 3]
    out[:, 3] = s * cloud[:, 2] + c * cloud[:, 3]
    return out.astype(np.float32)


# ------------------------------------------------------------------ projector (A.1)
def test_projector_empty_cells_and_range_gates(oracle):
    prm = oracle.default_params(canvas_cols=360, range_min=1.0, range_max=5.0)
    pts = np.array([[0.5, 0, -1, 0], [2.0, 0, -1, 0], [6.0, 0, -1, 0], [0, 3.0, 0, -1]], np.float32)
    img = oracle.project(prm, (0, 0, 0), pts)
    hit = np.flatnonzero(img["source_idx"] >= 0)
    assert set(img["source_idx"][hit]) == {1, 3}          # 0.5 m < range_min, 6 m > range_max rejected
    assert np.all(img["depth"][img["source_idx"] < 0] == FLT_MAX)  # decision D4
    # u = K00*theta + K01 with K00 = C/(max-min), K01 = C/2: theta = 0 -> column C/2, theta = pi/2 -> 3C/4
    assert img["source_idx"][180] == 1 and img["source_idx"][270] == 3
    assert img["depth"][180] == np.float32(2.0) and img["depth"][270] == np.float32(3.0)


def test_projector_zbuffer_nearest_then_first(oracle):
    prm = oracle.default_params(canvas_cols=90)
    pts = np.array([[5, 0, -1, 0], [3, 0, -1, 0], [4, 0, -1, 0], [3, 0, 0, -1]], np.float32)
    img = oracle.project(prm, (0, 0, 0), pts)
    cell = img[45]
    assert cell["source_idx"] == 1 and cell["depth"] == np.float32(3.0)   # nearest wins; first of the tie wins (D3)
    assert cell["nx"] == -1.0


def test_projector_boundary_range_is_inclusive(oracle):
    prm = oracle.default_params(canvas_cols=90, range_min=0.3, range_max=20.0)
    pts = np.array([[0.3, 0, -1, 0]], np.float32)
    assert oracle.project(prm, (0, 0, 0), pts)["source_idx"][45] == 0      # rho == range_min kept (D2b)
    pts = np.array([[20.0, 0, -1, 0]], np.float32)
    assert oracle.project(prm, (0, 0, 0), pts)["source_idx"][45] == 0      # rho == range_max kept


def test_projector_column_out_of_canvas_is_dropped(oracle):
    # theta = +pi maps to u = C exactly -> column C -> outside [0, C)
    prm = oracle.default_params(canvas_cols=100)
    pts = np.array([[-2.0, 0.0, 1, 0]], np.float32)
    img = oracle.project(prm, (0, 0, 0), pts)
    assert (img["source_idx"] >= 0).sum() == 0
    # a narrow field of view drops everything outside it
    prm = oracle.default_params(canvas_cols=100, angle_col_min=-0.5, angle_col_max=0.5)
    pts = np.array([[2.0, 0.0, -1, 0], [0.0, 2.0, 0, -1]], np.float32)
    img = oracle.project(prm, (0, 0, 0), pts)
    assert list(img["source_idx"][img["source_idx"] >= 0]) == [0]


def test_projector_camera_pose_is_inverted(oracle):
    """setCameraPose(T) projects T^-1 * p (the finder passes local_map_in_sensor.inverse(),
    correspondence_finder_projective_2d.cpp:47, and then reads `transformed` in the fixed frame)."""
    prm = oracle.default_params(canvas_cols=360)
    cam = (1.0, 0.5, 0.3)
    p_world = np.array([[3.0, 1.0, -1, 0]], np.float32)
    img = oracle.project(prm, cam, p_world)
    k = int(np.flatnonzero(img["source_idx"] >= 0)[0])
    c, s = math.cos(0.3), math.sin(0.3)
    dx, dy = 3.0 - 1.0, 1.0 - 0.5
    ex, ey = c * dx + s * dy, -s * dx + c * dy
    assert abs(img["px"][k] - ex) < 1e-5 and abs(img["py"][k] - ey) < 1e-5
    assert abs(img["depth"][k] - math.hypot(ex, ey)) < 1e-5


def test_demo_scene_projection_shape(oracle):
    """apps/synthetic_scene_generator.cpp: 1024 bins over +-0.4 pi from (0.2, 0.2, 0.1): every bin inside the
    3.5 m circle sees either the circle or the corner."""
    scene = reference_demo_scene()
    assert scene.shape == (2048 + 1023, 4)
    prm = oracle.default_params(canvas_cols=1024, angle_col_min=np.float32(-math.pi * 0.4),
                                angle_col_max=np.float32(math.pi * 0.4), range_min=0.01)
    img = oracle.project(prm, (0.2, 0.2, 0.1), scene)
    hit = img["source_idx"] >= 0
    assert hit.sum() > 800
    # the 3 m corner leg sticks out of the circle and shows through its sparse sampling
    assert img["depth"][hit].max() < 4.6 and img["depth"][hit].min() > 1.5


# ------------------------------------------------------------------ finder (A.2)
def test_finder_thresholds_are_strict(oracle):
    """fabs(dd) > point_distance and dot < normal_cos reject; equality passes (.cpp:65,69)."""
    prm = oracle.default_params(canvas_cols=90, point_distance=0.5, normal_cos=0.8)
    fixed = np.array([[3.0, 0, -1, 0]], np.float32)
    fi, mi, _, _ = oracle.find_correspondences(prm, fixed, np.array([[3.5, 0, -1, 0]], np.float32), (0, 0, 0))
    assert len(fi) == 1                                                      # |dd| == 0.5 passes
    fi, _, _, _ = oracle.find_correspondences(prm, fixed, np.array([[3.5001, 0, -1, 0]], np.float32), (0, 0, 0))
    assert len(fi) == 0
    n = np.array([[3.0, 0, -0.8, 0.6]], np.float32)                           # dot = 0.8 passes
    fi, _, _, _ = oracle.find_correspondences(prm, fixed, n, (0, 0, 0))
    assert len(fi) == 1
    n = np.array([[3.0, 0, -0.79, 0.61]], np.float32)
    fi, _, _, _ = oracle.find_correspondences(prm, fixed, n, (0, 0, 0))
    assert len(fi) == 0


def test_finder_orders_by_column_and_reports_source_indices(oracle):
    prm = oracle.default_params(canvas_cols=721)
    cloud = corner_cloud(300)
    perm = np.random.default_rng(0).permutation(len(cloud))
    fi, mi, fimg, mimg = oracle.find_correspondences(prm, cloud, cloud[perm], (0, 0, 0))
    assert len(fi) > 100
    cols_f = {int(i): k for k, i in enumerate(fimg["source_idx"]) if i >= 0}
    order = [cols_f[int(i)] for i in fi]
    assert order == sorted(order)                                            # ascending column (.cpp:55-74)
    assert np.array_equal(perm[mi], fi)                                      # same physical point on both sides


def test_finder_empty_inputs(oracle):
    prm = oracle.default_params(canvas_cols=90)
    empty = np.zeros((0, 4), np.float32)
    fi, mi, fimg, _ = oracle.find_correspondences(prm, empty, wall_cloud(10), (0, 0, 0))
    assert len(fi) == 0 and (fimg["source_idx"] < 0).all()
    fi, mi, _, mimg = oracle.find_correspondences(prm, wall_cloud(10), empty, (0, 0, 0))
    assert len(fi) == 0 and (mimg["source_idx"] < 0).all()


# ------------------------------------------------------------------ factor (A.3)
@pytest.mark.parametrize("factor", [0, 1])
@pytest.mark.parametrize("with_sensor", [0, 1])
def test_jacobian_matches_finite_differences(oracle, with_sensor, factor):
    """J is the derivative of e for the post-multiplied increment X <- X * v2t(dx) (nicp_post.m:96); factor 1 = the
    point-to-point factor (decision D19)."""
    prm = oracle.default_params(with_sensor=with_sensor, sensor_in_robot=(0.2, -0.1, 0.3), factor=factor)
    rng = np.random.default_rng(3)
    for _ in range(20):
        X = rng.uniform(-1, 1, 3)
        a, b = rng.uniform(-math.pi, math.pi, 2)
        pf = np.array([*rng.uniform(-5, 5, 2), math.cos(a), math.sin(a)], np.float32)
        pm = np.array([*rng.uniform(-5, 5, 2), math.cos(b), math.sin(b)], np.float32)
        e0, J = oracle.error_and_jacobian(prm, X, pf, pm)
        Xiso = oracle.v2t(*X)
        num = np.zeros((3, 3))
        h = 1e-3
        for k in range(3):
            for sgn in (1, -1):
                d = np.zeros(3)
                d[k] = sgn * h
                Xp = oracle.lib().orc_compose(Xiso, oracle.v2t(*d))
                xyt = np.zeros(3, np.float32)
                oracle.lib().orc_t2v(Xp, xyt.ctypes.data)
                e, _ = oracle.error_and_jacobian(prm, xyt, pf, pm)
                num[:, k] += sgn * e.astype(np.float64) / (2 * h)
        assert np.allclose(J, num, atol=2e-2, rtol=2e-2), (J, num)


def test_error_is_zero_at_ground_truth(oracle):
    prm = oracle.default_params()
    X = (0.3, -0.2, 0.4)
    pm = np.array([2.0, 1.0, math.cos(1.0), math.sin(1.0)], np.float32)
    pf = transform(pm[None], X)[0]
    e, _ = oracle.error_and_jacobian(prm, X, pf, pm)
    assert np.abs(e).max() < 1e-6


# ------------------------------------------------------------------ aligner (A.4 - A.7)
def test_identical_clouds_give_zero_update(oracle):
    prm = oracle.default_params(canvas_cols=721, max_iterations=5)
    cloud = corner_cloud(600)
    res, its = oracle.align(prm, cloud, cloud, (0, 0, 0))
    assert res["status"] == 0 and res["iterations"] == 5
    assert abs(res["x"]) < 1e-6 and abs(res["y"]) < 1e-6 and abs(res["theta"]) < 1e-6
    assert res["chi_inliers"] < 1e-9 and res["n_kernelized"] == 0 and res["n_inliers"] == res["n_corr"]


def test_corner_translation_converges_to_ground_truth(oracle):
    prm = oracle.default_params(canvas_cols=721, max_iterations=15, point_distance=0.5, normal_cos=0.9)
    fixed = corner_cloud(800)
    gt = (0.05, -0.04, 0.03)
    inv = oracle.lib().orc_inverse(oracle.v2t(*gt))
    xyt = np.zeros(3, np.float32)
    oracle.lib().orc_t2v(inv, xyt.ctypes.data)
    moving = transform(fixed, xyt)          # moving = gt^-1 * fixed  =>  moving_in_fixed = gt
    res, its = oracle.align(prm, fixed, moving, (0, 0, 0))
    assert res["status"] == 0
    assert abs(res["x"] - gt[0]) < 2e-3 and abs(res["y"] - gt[1]) < 2e-3 and abs(res["theta"] - gt[2]) < 2e-3
    assert its["chi_inliers"][-1] + its["chi_kernelized"][-1] < its["chi_inliers"][0] + its["chi_kernelized"][0]


def test_status_codes(oracle):
    cloud = corner_cloud(400)
    far = cloud.copy()
    far[:, 0] += 100.0
    prm = oracle.default_params(canvas_cols=721, max_iterations=4)
    res, _ = oracle.align(prm, cloud, far, (0, 0, 0))
    assert res["status"] == 1                                                                 # NotEnoughCorrespondences
    assert res["iterations"] == 0 and res["n_corr"] == 0
    prm = oracle.default_params(canvas_cols=721, max_iterations=4, min_num_inliers=100000)
    res, _ = oracle.align(prm, cloud, cloud, (0, 0, 0))
    assert res["status"] == 2                                                                 # NotEnoughInliers
    wall = wall_cloud(200)
    wall[:, 2:] = 0.0            # zero normals on both sides: J has only the (zero) point-to-line row
    prm = oracle.default_params(canvas_cols=721, max_iterations=4, normal_cos=-1.0)
    res, _ = oracle.align(prm, wall, wall, (0, 0, 0))
    assert res["status"] == 3 and res["iterations"] == 0                                      # singular H


def test_min_num_correspondences_gate_is_inclusive(oracle):
    cloud = corner_cloud(400)
    prm = oracle.default_params(canvas_cols=721, max_iterations=2)
    n_corr = oracle.align(prm, cloud, cloud, (0, 0, 0))[0]["n_corr"]
    prm = oracle.default_params(canvas_cols=721, max_iterations=2, min_num_correspondences=int(n_corr))
    assert oracle.align(prm, cloud, cloud, (0, 0, 0))[0]["status"] == 1      # n_corr <= min  =>  stop
    prm = oracle.default_params(canvas_cols=721, max_iterations=2, min_num_correspondences=int(n_corr) - 1)
    assert oracle.align(prm, cloud, cloud, (0, 0, 0))[0]["status"] == 0


def test_cauchy_kernel_bookkeeping(oracle):
    """chi < tau => inlier; else kernelized with rho = tau*ln(1+chi/tau) (decision D6)."""
    fixed = wall_cloud(300)
    moving = fixed.copy()
    moving[:, 0] += 0.2                      # point-to-line error 0.2 -> chi = 0.04 >= tau = 0.01
    prm = oracle.default_params(canvas_cols=721, max_iterations=1, cauchy_chi_threshold=0.01)
    res, _ = oracle.align(prm, fixed, moving, (0, 0, 0))
    assert res["n_inliers"] == 0 and res["n_kernelized"] == res["n_corr"] > 0
    expect = res["n_corr"] * 0.01 * math.log(1 + 0.04 / 0.01)
    assert abs(res["chi_kernelized"] - expect) < 1e-3 * expect
    prm = oracle.default_params(canvas_cols=721, max_iterations=1, cauchy_chi_threshold=-1.0)
    res, _ = oracle.align(prm, fixed, moving, (0, 0, 0))
    assert res["n_kernelized"] == 0 and abs(res["chi_inliers"] - res["n_corr"] * 0.04) < 1e-3


def test_tree_and_sequential_sums_agree_within_tolerance(oracle):
    sp = make_scan_pairs(8, n_beams=721, seed=5)
    prm = oracle.default_params(canvas_cols=721, normal_cos=0.9)
    seq, _ = oracle.align_batch(prm, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, sp.init_xyt)
    tree, _ = oracle.align_batch(prm, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, sp.init_xyt,
                                 sum_mode=oracle.SUM_TREE, tree_threads=256)
    assert np.array_equal(seq["n_corr"], tree["n_corr"])
    assert np.abs(seq["x"] - tree["x"]).max() < 1e-5 and np.abs(seq["theta"] - tree["theta"]).max() < 1e-5
    assert np.allclose(seq["chi_inliers"], tree["chi_inliers"], rtol=1e-3)


@pytest.mark.parametrize("shape", [288 | 1 << 16, 288 | 1 << 16 | 1 << 17, 544 | 1 << 16, 256 | 1 << 17])
def test_kernel_shaped_sum_modes_agree_with_the_reference_order(oracle, shape):
    """every ORC_SUM_TREE variant the kernels use (warp-combine flag bit 16, fused accumulation arithmetic of decision
    D18 bit 17) against the reference's sequential order: same integer outcomes at the first linearisation, poses and
    chi2 inside BASELINE.json's tolerances; the fused mode really is a different arithmetic (some bits differ)"""
    sp = make_scan_pairs(16, n_beams=1081, seed=21)
    prm = oracle.default_params(canvas_cols=1081, normal_cos=0.9)
    seq, sit = oracle.align_batch(prm, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, sp.init_xyt)
    tree, tit = oracle.align_batch(prm, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, sp.init_xyt,
                                   sum_mode=oracle.SUM_TREE, tree_threads=shape)
    assert np.array_equal(sit["n_corr"][:, 0], tit["n_corr"][:, 0])
    assert np.abs(seq["x"] - tree["x"]).max() < 1e-5 and np.abs(seq["y"] - tree["y"]).max() < 1e-5
    assert np.abs(seq["theta"] - tree["theta"]).max() < 2e-6
    assert np.allclose(seq["chi_inliers"], tree["chi_inliers"], rtol=1e-3)
    if shape >> 17 & 1:
        unfused, _ = oracle.align_batch(prm, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, sp.init_xyt,
                                        sum_mode=oracle.SUM_TREE, tree_threads=shape & ~(1 << 17))
        assert not np.array_equal(unfused["chi_inliers"].view(np.uint32), tree["chi_inliers"].view(np.uint32))


def test_synthetic_pairs_converge_near_ground_truth(oracle):
    sp = make_scan_pairs(6, seed=9)
    prm = oracle.default_params(canvas_cols=1081, normal_cos=0.9)
    res, _ = oracle.align_batch(prm, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, sp.init_xyt,
                                n_threads=2)
    assert (res["status"] == 0).all()
    assert np.abs(res["x"] - sp.gt_xyt[:, 0]).max() < 5e-3
    assert np.abs(res["y"] - sp.gt_xyt[:, 1]).max() < 5e-3
    assert np.abs(res["theta"] - sp.gt_xyt[:, 2]).max() < 3e-3


def test_reference_synthetic_fixture_shape(oracle):
    """tests/fixtures.hpp:38-47 of the reference: 100 beams, range 1.0 m, -1..+1 rad, increment 0.02.  All 100
    points land in distinct columns of a 721-column full-circle projector (the finder's view of that scan)."""
    ang = (-1.0 + 0.02 * np.arange(100)).astype(np.float32)
    pts = np.stack([np.cos(ang), np.sin(ang), -np.cos(ang), -np.sin(ang)], -1).astype(np.float32)
    prm = oracle.default_params(canvas_cols=721)
    img = oracle.project(prm, (0, 0, 0), pts)
    assert (img["source_idx"] >= 0).sum() == 100
    assert np.allclose(img["depth"][img["source_idx"] >= 0], 1.0, atol=1e-6)


# ------------------------------------------------------------------ verification gates (A.8)
def test_acceptance_gates_and_best_of(oracle):
    r = np.zeros(5, oracle.RESULT_DTYPE)
    r["status"] = [0, 0, 0, 2, 0]
    r["n_inliers"] = [400, 500, 500, 900, 299]
    r["n_corr"] = [450, 520, 520, 900, 300]
    r["chi_inliers"] = [4.0, 10.0, 5.0, 1.0, 1.0]
    assert oracle.best_of(r, 300, 0.1, 0.8) == 2          # most inliers, then lowest chi per inlier
    r["chi_inliers"][2] = 10.0
    assert oracle.best_of(r, 300, 0.1, 0.8) == 1          # full tie keeps the lowest id
    r["chi_inliers"][:] = 1000.0
    assert oracle.best_of(r, 300, 0.1, 0.8) == -1
