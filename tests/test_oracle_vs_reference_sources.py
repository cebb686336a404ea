"""The oracle's restatement of the reference's IN-REPO sources against those sources themselves (CPU only).

oracle/_ref/libls2d_ref.so is built by oracle/Makefile from three reference files compiled where they lie under
/root/reference, unmodified -- correspondence_finder_projective_2d.cpp, merger_projective_2d.cpp,
scene_clipper_projective_2d.cpp, raw_data_preprocessor_projective_2d.cpp -- against stand-in headers (oracle/ref_shim/) for the absent upstream libraries.
The upstream pieces (polar projector, isometry algebra, point arithmetic) are the oracle's own restatement in both
arms, so what these tests pin is exactly what lives in the reference repository: the finder's gates, their
strictness, ordering and caching (.cpp:18-77), the merger's per-column decision tree and its ordered appends
(.cpp:9-100), the clipper's collection, voxelize branch and move to the robot frame (.cpp:11-65), the pre-processor's
range limits, sensor matrix, voxelize / valid-only branches and message handling (.cpp:13-51, 53-104).  The projector / factor / solver
arithmetic stays unpinned."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from srrg2_laser_slam_2d_b200.synthetic import make_raw_scans, make_scan_pairs



def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


CASES = [dict(canvas_cols=1081, point_distance=0.5, normal_cos=0.9),     # LASER_0.json tracking finder
         dict(canvas_cols=721, point_distance=1.414, normal_cos=0.8),    # LASER_0.json loop-closure finder
         dict(canvas_cols=361, point_distance=0.05, normal_cos=0.999),   # gates that reject a lot
         dict(canvas_cols=1081, angle_col_min=-2.35619, angle_col_max=2.35619, point_distance=0.5, normal_cos=-1.0)]


@pytest.mark.parametrize("kw", CASES)
def test_finder_gates_order_and_caching(oracle, ref, kw):
    prm = oracle.default_params(**kw)
    n_beams = 721 if kw["canvas_cols"] == 721 else 1081
    sp = make_scan_pairs(24, n_beams=n_beams, seed=11 + kw["canvas_cols"], motion_xy=0.3, motion_theta=0.15)
    rng = np.random.default_rng(3)
    total = 0
    for p in range(24):
        f = sp.fixed_pts[sp.fixed_off[p]:sp.fixed_off[p + 1]]
        m = sp.moving_pts[sp.moving_off[p]:sp.moving_off[p + 1]]
        poses = np.concatenate([rng.uniform(-0.3, 0.3, (2, 3)), sp.gt_xyt[p:p + 1]]).astype(np.float32)
        fi = np.full(prm.canvas_cols, -7, np.int32)
        mi = np.full(prm.canvas_cols, -7, np.int32)
        # three compute() calls on one object: the fixed image is projected once and reused (.cpp:37-44)
        k = ref.ref_find_correspondences(C.byref(prm), _p(f), len(f), _p(m), len(m), _p(poses), 3, _p(fi), _p(mi))
        ofi, omi, _, _ = oracle.find_correspondences(prm, f, m, poses[2])
        assert k == len(ofi)
        assert np.array_equal(fi[:k], ofi) and np.array_equal(mi[:k], omi)
        total += k
    assert total > 24 * 20  # the cases do produce correspondences


def test_finder_throws_on_missing_inputs(ref):
    assert ref.ref_finder_throws_without_inputs() == 1


@pytest.mark.parametrize("cols,thr", [(721, 0.2), (1081, 0.2), (721, 0.02), (361, 0.5)])
def test_merger_decision_tree_and_appends(oracle, ref, cols, thr):
    prm = oracle.default_params(canvas_cols=cols)
    sp = make_scan_pairs(16, n_beams=721, seed=900 + cols, motion_xy=0.4, motion_theta=0.2, range_noise=0.03)
    n_changed = 0
    for p in range(16):
        scene = sp.fixed_pts[sp.fixed_off[p]:sp.fixed_off[p + 1]].copy()
        meas = sp.moving_pts[sp.moving_off[p]:sp.moving_off[p + 1]]
        for xyt in (sp.gt_xyt[p], sp.gt_xyt[p] + np.float32([0.3, -0.2, 0.1])):   # aligned and misaligned
            xyt = np.ascontiguousarray(xyt, np.float32)
            want, counters = oracle.merge(prm, thr, scene, meas, xyt)
            buf = np.zeros((len(scene) + cols, 4), np.float32)
            buf[:len(scene)] = scene
            n = ref.ref_merge(C.byref(prm), thr, _p(buf), len(scene), _p(meas), len(meas), _p(xyt))
            assert n == len(want)
            assert np.array_equal(_bits(buf[:n]), _bits(want))
            n_changed += int(counters.sum())
            scene = want  # keep merging into the grown scene, as the tracker does
    assert n_changed > 0


@pytest.mark.parametrize("cols,sensor,voxel", [(721, (0.0, 0.0, 0.0), 0.0), (1081, (0.2, 0.2, 0.1), 0.0),
                                               (361, (-0.1, 0.05, -0.3), 0.0), (1081, (0.0, 0.0, 0.0), 0.1),
                                               (721, (0.2, 0.2, 0.1), 0.05)])
def test_clipper_collects_voxelizes_and_moves_to_the_robot_frame(oracle, ref, cols, sensor, voxel):
    prm = oracle.default_params(canvas_cols=cols)
    sp = make_scan_pairs(12, n_beams=1081, seed=77 + cols)
    for p in range(12):
        scene = np.concatenate([sp.fixed_pts[sp.fixed_off[p]:sp.fixed_off[p + 1]],
                                sp.moving_pts[sp.moving_off[p]:sp.moving_off[p + 1]]])   # a denser "local map"
        robot = np.ascontiguousarray(sp.gt_xyt[p] * 3, np.float32)
        sens = np.float32(sensor)
        want = oracle.clip_scene(prm, scene, robot, sens, voxel)
        out = np.zeros((cols, 4), np.float32)
        k = ref.ref_clip(C.byref(prm), _p(scene), len(scene), _p(robot), _p(sens), voxel, _p(out))
        assert k == len(want) and k > 0
        assert np.array_equal(_bits(out[:k]), _bits(want))


def _ref_preprocess(ref, sp, ranges):
    ranges = np.ascontiguousarray(ranges, np.float32)
    out = np.zeros((len(ranges), 4), np.float32)
    k = ref.ref_preprocess_scan(C.byref(sp), _p(ranges), len(ranges), _p(out))
    return out[:k]


def test_preprocessor_synthetic_fixture_of_the_reference_tests(oracle, ref):
    """tests/fixtures.hpp:38-47 through the reference's own RawDataPreprocessorProjective2D::setRawData + compute:
    100 points (tests/test_measurement_adaptor.cpp:36), the oracle's cloud bit for bit"""
    from test_oracle_preprocess import synthetic_fixture
    sp, ranges = synthetic_fixture()
    got = _ref_preprocess(ref, sp, ranges)
    assert len(got) == 100
    assert np.array_equal(_bits(got), _bits(oracle.preprocess_scan(sp, ranges)))


@pytest.mark.parametrize("voxel", [0.0, 0.02, 0.1])
@pytest.mark.parametrize("limits", [dict(msg_range_min=0.1, msg_range_max=30.0, range_min=0.3, range_max=20.0),
                                    dict(msg_range_min=0.5, msg_range_max=8.0, range_min=0.0, range_max=1000.0)])
def test_preprocessor_range_limits_sensor_matrix_and_branches(oracle, ref, voxel, limits):
    raw = make_raw_scans(6, n_beams=1081, seed=4242)
    sp = oracle.default_scan_params(angle_min=raw.angle_min, angle_max=raw.angle_max, voxelize_resolution=voxel,
                                    **limits)
    n_pts = 0
    for ranges in np.concatenate([raw.fixed_ranges, raw.moving_ranges]):
        want = oracle.preprocess_scan(sp, ranges)
        got = _ref_preprocess(ref, sp, ranges)
        assert len(got) == len(want)
        assert np.array_equal(_bits(got), _bits(want))
        n_pts += len(want)
    assert n_pts > 12 * 100
