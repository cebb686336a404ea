"""The rows of the path chained the way MultiTracker2D chains them (apps/visual_test_tracker_2d.cpp:167-179:
setRawData -> preprocessRawData -> clip -> align -> merge), frame after frame on one growing local map, on the device
and in the oracle: every intermediate cloud and every pose of the two runs must be bit-identical, and the tracked
trajectory must follow the ground truth.  (The reference's own integration test, tests/test_slam.cpp, needs a dataset
and a configuration that are not in the repository.)"""
import numpy as np
import pytest

import golden_util as gu
from srrg2_laser_slam_2d_b200 import default_params
from srrg2_laser_slam_2d_b200._abi import LS2D_FIXED, LS2D_MOVING, default_scan_params, reduction_threads
from srrg2_laser_slam_2d_b200.synthetic import make_scan_sequence

pytestmark = pytest.mark.gpu


def v2t(v):
    c, s = np.cos(v[2]), np.sin(v[2])
    return np.array([[c, -s, v[0]], [s, c, v[1]], [0, 0, 1]])


def t2v(T):
    return np.array([T[0, 2], T[1, 2], np.arctan2(T[1, 0], T[0, 0])])


@pytest.mark.parametrize("voxel", [0.02, 0.0])
def test_tracker_chain_is_bit_identical_and_follows_ground_truth(handle_factory, oracle, voxel):
    n_frames, cols = 40, 721
    seq = make_scan_sequence(n_frames, n_beams=721, seed=11)
    skw = dict(angle_min=seq.angle_min, angle_max=seq.angle_max, voxelize_resolution=voxel)
    akw = dict(canvas_cols=cols, normal_cos=0.9, max_iterations=10)                 # LASER_0.json tracking values
    sp, osp = default_scan_params(**skw), oracle.default_scan_params(**skw)
    prm = oracle.default_params(**akw)
    h = handle_factory(default_params(**akw))
    zero = np.zeros((1, 3), np.float32)

    def run(gpu: bool):
        poses = [np.zeros(3, np.float32)]                                         # robot_in_local_map, frame 0 = origin
        scene = None
        trace = []
        for f in range(n_frames):
            ranges = seq.ranges[f:f + 1]
            if gpu:
                pts, cnt = h.preprocess_scans(sp, ranges)
                meas = pts[0, :cnt[0]].copy()
            else:
                meas = oracle.preprocess_scan(osp, ranges[0])
            if scene is None:
                scene = meas.copy()
                trace.append((meas, scene.copy(), poses[-1]))
                continue
            guess = poses[-1]                                                     # constant-position motion model
            if gpu:
                h.upload_clouds(2, scene, np.array([0, len(scene)], np.int32))
                clip = h.clip_scenes(2, [0], guess[None, :])[0]
                h.upload_clouds(LS2D_FIXED, meas, np.array([0, len(meas)], np.int32))
                h.upload_clouds(LS2D_MOVING, clip, np.array([0, len(clip)], np.int32))
                res = h.align_batch(zero)[0]
            else:
                clip = oracle.clip_scene(prm, scene, guess, (0.0, 0.0, 0.0))
                res, _ = oracle.align(prm, meas, clip, zero[0], sum_mode=oracle.SUM_TREE,
                                      tree_threads=reduction_threads(max(len(meas), len(clip)), cols))
            assert res["status"] == 0
            # X maps the clipped scene (robot frame at the guess) onto the measurement (true robot frame)
            X = v2t(np.array([res["x"], res["y"], res["theta"]], np.float64))
            pose = t2v(v2t(guess.astype(np.float64)) @ np.linalg.inv(X)).astype(np.float32)
            if gpu:
                scene, _ = h.merge_scene(scene, meas, pose, 0.2)
            else:
                scene, _ = oracle.merge(prm, 0.2, scene, meas, pose)
            poses.append(pose)
            trace.append((meas, scene.copy(), pose))
        return trace

    got, ref = run(True), run(False)
    for f, (g, r) in enumerate(zip(got, ref)):
        for a, b in zip(g, r):
            assert a.shape == b.shape and np.array_equal(gu.bits(a), gu.bits(b)), f
    # the local map grew, and the tracked pose follows the drive
    assert len(got[-1][1]) > 1.3 * len(got[0][1])
    T0 = np.linalg.inv(v2t(seq.poses[0]))
    worst = 0.0
    for f in range(n_frames):
        gt = t2v(T0 @ v2t(seq.poses[f]))
        d = got[f][2].astype(np.float64) - gt
        d[2] = (d[2] + np.pi) % (2 * np.pi) - np.pi
        worst = max(worst, float(np.hypot(d[0], d[1])), float(abs(d[2])))
    assert worst < 0.05, worst
