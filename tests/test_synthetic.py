"""Seeded workload generator: shape, determinism, invalid-beam encoding (CPU only)."""
import numpy as np

from srrg2_laser_slam_2d_b200.synthetic import FAR_POINT, make_scan_pairs


def test_shapes_and_determinism():
    a = make_scan_pairs(5, n_beams=361, seed=3)
    b = make_scan_pairs(5, n_beams=361, seed=3)
    c = make_scan_pairs(5, n_beams=361, seed=4)
    assert a.fixed_pts.shape == (5 * 361, 4) and a.fixed_pts.dtype == np.float32
    assert np.array_equal(a.fixed_off, np.arange(6) * 361)
    assert np.array_equal(a.fixed_pts, b.fixed_pts) and np.array_equal(a.moving_pts, b.moving_pts)
    assert not np.array_equal(a.fixed_pts, c.fixed_pts)
    assert np.abs(a.gt_xyt).max() <= 0.05 and (a.init_xyt == 0).all()


def test_invalid_beams_are_far_points_and_normals_are_unit():
    sp = make_scan_pairs(4, seed=1)
    pts = sp.fixed_pts
    far = (pts == np.array(FAR_POINT, np.float32)).all(1)
    assert 0.005 < far.mean() < 0.1                      # ~2 % dropouts + a few failed normal fits
    good = pts[~far]
    assert np.abs(np.hypot(good[:, 2], good[:, 3]) - 1.0).max() < 1e-5
    rho = np.hypot(good[:, 0], good[:, 1])
    assert rho.min() > 0.05 and rho.max() < 20.0
    assert ((good[:, 0] * good[:, 2] + good[:, 1] * good[:, 3]) <= 0).all()   # normals face the sensor


def test_loop_closure_guesses_are_perturbed_ground_truth():
    sp = make_scan_pairs(6, n_beams=181, seed=2, motion_xy=0.3, motion_theta=0.2, init_noise_xy=0.1,
                         init_noise_theta=0.05)
    d = np.abs(sp.init_xyt - sp.gt_xyt)
    assert d[:, :2].max() < 0.2 and d[:, 2].max() < 0.06 and d.max() > 0
