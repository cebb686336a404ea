"""GPU parity of the local-map maintenance modules next to the aligner:
SceneClipperProjective2D (R/mapping/scene_clipper_projective_2d.cpp) and MergerProjective2D
(R/mapping/merger_projective_2d.cpp), through the C ABI, bit-exact against the oracle."""
import numpy as np
import pytest

import golden_util as gu
from srrg2_laser_slam_2d_b200 import default_params
from srrg2_laser_slam_2d_b200._abi import LS2D_FIXED, LS2D_MOVING
from srrg2_laser_slam_2d_b200.synthetic import FAR_POINT, make_scan_pairs

pytestmark = pytest.mark.gpu


def local_map(sp, k, n=3):
    """a local-map-like scene: n scans of the same room merged by concatenation (unordered, overlapping)"""
    pts = [sp.fixed_pts[sp.fixed_off[k]:sp.fixed_off[k + 1]], sp.moving_pts[sp.moving_off[k]:sp.moving_off[k + 1]]]
    return np.concatenate(pts[:n])


@pytest.mark.parametrize("cols,sensor", [(721, (0.0, 0.0, 0.0)), (721, (0.2, 0.1, 0.3)), (1081, (0.1, -0.2, -0.4)),
                                         (90, (0.0, 0.0, 0.0))])
def test_scene_clipper_bit_exact(handle_factory, oracle, cols, sensor):
    sp = make_scan_pairs(8, n_beams=900, seed=61)
    scenes = [local_map(sp, k) for k in range(8)]
    off = np.concatenate([[0], np.cumsum([len(s) for s in scenes])]).astype(np.int32)
    kw = dict(canvas_cols=cols)
    h = handle_factory(default_params(**kw))
    h.upload_clouds(LS2D_FIXED, np.concatenate(scenes), off)
    rng = np.random.default_rng(5)
    ids = np.array([0, 3, 3, 7, 1, 5], np.int32)
    robots = rng.uniform(-0.3, 0.3, (len(ids), 3)).astype(np.float32)
    got = h.clip_scenes(LS2D_FIXED, ids, robots, sensor)
    prm = oracle.default_params(**kw)
    for r, cid in enumerate(ids):
        ref = oracle.clip_scene(prm, scenes[cid], robots[r], sensor)
        assert got[r].shape == ref.shape and len(ref) > cols // 4
        assert np.array_equal(gu.bits(got[r]), gu.bits(ref))


def test_scene_clipper_edge_cases(handle_factory, oracle):
    kw = dict(canvas_cols=361)
    h = handle_factory(default_params(**kw))
    far = np.tile(np.array(FAR_POINT, np.float32), (40, 1))
    sp = make_scan_pairs(1, n_beams=361, seed=2)
    clouds = [far, sp.fixed_pts[:0], sp.fixed_pts]
    off = np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).astype(np.int32)
    h.upload_clouds(LS2D_MOVING, np.concatenate(clouds), off)
    got = h.clip_scenes(LS2D_MOVING, [0, 1, 2], np.zeros((3, 3), np.float32))
    assert len(got[0]) == 0 and len(got[1]) == 0                       # nothing visible / empty scene
    ref = oracle.clip_scene(oracle.default_params(**kw), sp.fixed_pts, (0, 0, 0), (0, 0, 0))
    assert np.array_equal(gu.bits(got[2]), gu.bits(ref))
    # clipping from the pose a scan was taken at keeps (almost) every valid beam: the scan is its own z-buffer
    valid = (sp.fixed_pts[:, 0] < 1e5).sum()
    assert len(ref) >= 0.7 * valid


@pytest.mark.parametrize("cols,thr", [(721, 0.2), (1081, 0.2), (361, 0.05), (361, 5.0)])
def test_merger_bit_exact(handle_factory, oracle, cols, thr):
    sp = make_scan_pairs(6, n_beams=800, seed=71)
    kw = dict(canvas_cols=cols)
    h = handle_factory(default_params(**kw))
    prm = oracle.default_params(**kw)
    for k in range(6):
        scene = sp.fixed_pts[sp.fixed_off[k]:sp.fixed_off[k + 1]]
        meas = sp.moving_pts[sp.moving_off[k]:sp.moving_off[k + 1]]
        # the measurement was taken from gt; merge with a slightly wrong estimate so all four branches fire
        xyt = sp.gt_xyt[k] + np.array([0.03, -0.02, 0.01], np.float32)
        got, gc = h.merge_scene(scene, meas, xyt, thr)
        ref, rc = oracle.merge(prm, thr, scene, meas, xyt)
        assert np.array_equal(gc, rc), (gc, rc)
        assert got.shape == ref.shape and np.array_equal(gu.bits(got), gu.bits(ref))
        assert len(ref) >= len(scene)


def test_merger_grows_a_local_map_over_a_sequence(handle_factory, oracle):
    """tracker-like use: start from the first scan, merge the following ones; GPU and oracle stay bit-identical
    over the whole sequence (each step consumes the previous step's output)."""
    sp = make_scan_pairs(5, n_beams=721, seed=81)
    kw = dict(canvas_cols=721)
    h = handle_factory(default_params(**kw))
    prm = oracle.default_params(**kw)
    g = o = sp.fixed_pts[sp.fixed_off[0]:sp.fixed_off[1]]
    meas0 = sp.moving_pts[sp.moving_off[0]:sp.moving_off[1]]
    for step in range(4):
        xyt = sp.gt_xyt[0] * (1.0 + 0.5 * step)
        g, gc = h.merge_scene(g, meas0, xyt)
        o, oc = oracle.merge(prm, 0.2, o, meas0, xyt)
        assert np.array_equal(gc, oc) and np.array_equal(gu.bits(g), gu.bits(o))
    assert len(g) > 721


def test_merger_rejects_too_small_capacity(handle_factory):
    import ctypes as C
    from srrg2_laser_slam_2d_b200 import Ls2dError
    h = handle_factory(default_params(canvas_cols=361))
    scene = np.zeros((10, 4), np.float32)
    size = C.c_int32(10)
    rc = h._L.ls2d_merge_scene(h._h, scene.ctypes.data_as(C.c_void_p), C.byref(size), 100, None, 0,
                               np.zeros(3, np.float32).ctypes.data_as(C.c_void_p), 0.2, None)
    assert rc == -1


def test_tracker_rows_golden(handle_factory):
    """frozen fixture of the rows around the aligner (tests/golden/tracker_721_l0.npz): raw scans -> pre-processor ->
    clipper -> merger on the device, bit for bit"""
    from srrg2_laser_slam_2d_b200._abi import default_scan_params
    d = gu.load_raw("tracker_721_l0")
    kw = dict(angle_min=float(d["angles"][0]), angle_max=float(d["angles"][1]))
    h = handle_factory(default_params(canvas_cols=721))
    meas, cm = h.preprocess_scans(default_scan_params(**kw), d["fixed_ranges"])
    scene, cs = h.preprocess_scans(default_scan_params(voxelize_resolution=0.0, **kw), d["moving_ranges"])
    scenes = [scene[k, :cs[k]] for k in range(3)]
    off = np.concatenate([[0], np.cumsum(cs)]).astype(np.int32)
    h.upload_clouds(2, np.concatenate(scenes), off)
    clips = h.clip_scenes(2, [0, 1, 2], d["robot_in_local_map"], d["sensor_in_robot"])
    for k in range(3):
        assert np.array_equal(gu.bits(meas[k, :cm[k]]), gu.bits(d[f"meas_{k}"]))
        assert np.array_equal(gu.bits(scenes[k]), gu.bits(d[f"scene_{k}"]))
        assert np.array_equal(gu.bits(clips[k]), gu.bits(d[f"clip_{k}"]))
        merged, counters = h.merge_scene(scenes[k], meas[k, :cm[k]], d["gt_xyt"][k], 0.2)
        assert np.array_equal(gu.bits(merged), gu.bits(d[f"merged_{k}"]))
        assert np.array_equal(counters, d[f"merge_counters_{k}"])
    fx = default_scan_params(angle_min=-1.0, angle_max=1.0, msg_range_min=0.0, msg_range_max=1000.0, range_min=0.0,
                             range_max=1000.0, voxelize_resolution=0.01)
    cloud, n = h.preprocess_scans(fx, np.full((1, 100), 1.0, np.float32))
    assert n[0] == 100 and np.array_equal(gu.bits(cloud[0, :100]), gu.bits(d["synthetic_fixture_cloud"]))


@pytest.mark.parametrize("cols,res,sensor", [(721, 0.1, (0.2, 0.1, 0.3)), (1081, 0.05, (0.0, 0.0, 0.0)),
                                             (361, 0.5, (0.1, -0.2, -0.4)), (721, 0.02, (0.0, 0.0, 0.0))])
def test_scene_clipper_voxelize_branch_bit_exact(handle_factory, oracle, cols, res, sensor):
    """SceneClipperProjective2D with voxelize_resolution > 0 (R/mapping/scene_clipper_projective_2d.cpp:36-48):
    the winners are voxelized with res_coeffs (res, res, 0.1, 0.1) in the sensor frame, then moved to the robot"""
    sp = make_scan_pairs(6, n_beams=900, seed=63)
    scenes = [local_map(sp, k) for k in range(6)]
    off = np.concatenate([[0], np.cumsum([len(s) for s in scenes])]).astype(np.int32)
    kw = dict(canvas_cols=cols)
    h = handle_factory(default_params(**kw))
    h.upload_clouds(LS2D_FIXED, np.concatenate(scenes), off)
    rng = np.random.default_rng(6)
    ids = np.array([0, 5, 2, 2, 4], np.int32)
    robots = rng.uniform(-0.3, 0.3, (len(ids), 3)).astype(np.float32)
    got = h.clip_scenes(LS2D_FIXED, ids, robots, sensor, voxelize_resolution=res)
    prm = oracle.default_params(**kw)
    for r, cid in enumerate(ids):
        ref = oracle.clip_scene(prm, scenes[cid], robots[r], sensor, res)
        plain = oracle.clip_scene(prm, scenes[cid], robots[r], sensor)
        assert got[r].shape == ref.shape and 10 < len(ref) <= len(plain)
        assert np.array_equal(gu.bits(got[r]), gu.bits(ref))
