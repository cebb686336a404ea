"""CPU tests of the oracle's RawDataPreprocessorProjective2D restatement.

The only number the reference's own test-suite pins on this row: the `Synthetic` fixture
(/root/reference/srrg2_laser_slam_2d/tests/fixtures.hpp:38-47: 100 beams of 1.0 m over -1 .. +1 rad, voxel 0.01)
must give 100 points (tests/test_measurement_adaptor.cpp:36)."""
import numpy as np

import oracle_binding as ob


def synthetic_fixture():
    angle_min, angle_max, incr = np.float32(-1), np.float32(1), np.float32(0.02)
    n = int((angle_max - angle_min) / incr)  # fixtures.hpp:47, binary32 arithmetic
    sp = ob.default_scan_params(angle_min=-1.0, angle_max=1.0, msg_range_min=0.0, msg_range_max=1000.0,
                                range_min=0.0, range_max=1000.0, voxelize_resolution=0.01)
    return sp, np.full(n, 1.0, np.float32)


def test_reference_synthetic_fixture_gives_100_points():
    sp, ranges = synthetic_fixture()
    assert len(ranges) == 100
    cloud = ob.preprocess_scan(sp, ranges)
    assert len(cloud) == 100  # test_measurement_adaptor.cpp:36
    # a circle of radius 1 around the sensor: points on it, normals radial and facing the sensor
    rho = np.hypot(cloud[:, 0], cloud[:, 1])
    assert np.allclose(rho, 1.0, atol=1e-6)
    radial = cloud[:, :2] / rho[:, None]
    assert np.all((cloud[:, 2:] * radial).sum(1) < -0.98)  # one-sided windows at the scan ends tilt a little
    assert np.allclose(np.hypot(cloud[:, 2], cloud[:, 3]), 1.0, atol=1e-6)


def test_valid_only_branch_keeps_beam_order_and_symmetric_angles():
    sp, ranges = synthetic_fixture()
    sp.voxelize_resolution = 0.0
    cloud = ob.preprocess_scan(sp, ranges)
    assert len(cloud) == 100
    az = np.arctan2(cloud[:, 1], cloud[:, 0])
    assert np.all(np.diff(az) > 0)
    # sensor matrix K = [1/res, n/2] (raw_data_preprocessor_projective_2d.cpp:87-90): beam c sits at (c - n/2) * res
    assert np.allclose(az, (np.arange(100) - 50) * 0.02, atol=2e-6)


def test_range_limits_are_the_tighter_of_message_and_params():
    sp, ranges = synthetic_fixture()
    sp.voxelize_resolution = 0.0
    ranges = ranges.copy()
    ranges[10:20] = 50.0
    ranges[40] = 0.05
    sp.msg_range_max, sp.range_max = 30.0, 1000.0
    sp.msg_range_min, sp.range_min = 0.0, 0.1
    cloud = ob.preprocess_scan(sp, ranges)
    # 11 beams rejected; the neighbours of the gaps still have >= 5 points in their windows
    assert len(cloud) == 89


def test_sparse_points_have_no_normal_and_are_dropped():
    sp, _ = synthetic_fixture()
    sp.voxelize_resolution = 0.0
    ranges = np.full(100, 20.0, np.float32)  # 0.4 m between neighbours > normal_point_distance 0.3
    assert len(ob.preprocess_scan(sp, ranges)) == 0
    ranges[:50] = 1.0
    cloud = ob.preprocess_scan(sp, ranges)
    assert len(cloud) == 50


def test_wall_normals_face_the_sensor():
    n = 361
    sp = ob.default_scan_params(angle_min=-np.pi / 2, angle_max=np.pi / 2, voxelize_resolution=0.0)
    az = (np.arange(n, dtype=np.float32) - n / 2) * np.float32(np.pi / n)
    ranges = (2.0 / np.maximum(np.cos(az), 0.2)).astype(np.float32)  # a wall at x = 2
    ranges[np.cos(az) < 0.25] = 100.0
    cloud = ob.preprocess_scan(sp, ranges)
    assert len(cloud) > 200
    inner = cloud[5:-5]
    assert np.allclose(inner[:, 0], 2.0, atol=1e-5)
    assert np.allclose(inner[:, 2], -1.0, atol=1e-4) and np.allclose(inner[:, 3], 0.0, atol=2e-3)


def test_voxelize_merges_and_sorts():
    n = 721
    sp = ob.default_scan_params()
    rng = np.random.default_rng(5)
    az = (np.arange(n, dtype=np.float32) - n / 2) * np.float32((sp.angle_max - sp.angle_min) / n)
    ranges = (1.5 + 0.3 * np.sin(3 * az) + rng.normal(0, 0.003, n)).astype(np.float32)
    full = ob.preprocess_scan(ob.default_scan_params(voxelize_resolution=0.0), ranges)
    vox = ob.preprocess_scan(sp, ranges)
    assert 0 < len(vox) < len(full) == n
    key = np.trunc(vox[:, :2] / np.float32(0.02)).astype(np.int64)
    order = np.lexsort((key[:, 1], key[:, 0]))
    assert np.array_equal(order, np.arange(len(vox)))          # output is sorted by voxel
    assert np.allclose(np.hypot(vox[:, 2], vox[:, 3]), 1.0, atol=1e-6)
    # every full-resolution point has a voxel representative within one cell diagonal
    d = np.abs(full[:, None, :2] - vox[None, :, :2]).max(2).min(1)
    assert d.max() < 0.04


def test_batch_matches_single_and_threads():
    rng = np.random.default_rng(11)
    sp = ob.default_scan_params()
    ranges = rng.uniform(0.5, 6.0, (12, 1)).astype(np.float32) + rng.normal(0, 0.01, (12, 721)).astype(np.float32)
    ranges[rng.random(ranges.shape) < 0.02] = 0.0
    sp.range_min = 0.1
    pts1, cnt1 = ob.preprocess_scans(sp, ranges, n_threads=1)
    pts4, cnt4 = ob.preprocess_scans(sp, ranges, n_threads=4)
    assert np.array_equal(cnt1, cnt4) and np.array_equal(pts1.view(np.uint32), pts4.view(np.uint32))
    for s in range(12):
        one = ob.preprocess_scan(sp, ranges[s])
        assert len(one) == cnt1[s] and np.array_equal(one.view(np.uint32), pts1[s, :cnt1[s]].view(np.uint32))


def test_empty_scan():
    sp = ob.default_scan_params()
    assert len(ob.preprocess_scan(sp, np.zeros(0, np.float32))) == 0
    assert len(ob.preprocess_scan(sp, np.full(64, 1e9, np.float32))) == 0
