"""The oracle must reproduce the frozen fixtures of tests/golden/ bit for bit (CPU only).

Parity unpinned: the reference holds no golden vectors for this path and cannot be built here, so the
fixtures were produced by this oracle (tools/make_golden.py); this test guards its decision points
against accidental change and checks the fixtures' internal consistency."""
import numpy as np
import pytest

import golden_util as gu


@pytest.mark.parametrize("name", gu.ALIGN_CASES)
def test_oracle_reproduces_alignment_golden(oracle, name):
    d = gu.load(name)
    prm = gu.make_params(oracle.default_params, d)
    res, its = oracle.align_batch(prm, d["fixed_pts"], d["fixed_off"], d["moving_pts"], d["moving_off"],
                                  d["init_xyt"])
    assert set(res.dtype.names) == set(d["results"].dtype.names)
    for f in res.dtype.names:
        assert np.array_equal(res[f], d["results"][f]), f
    for f in its.dtype.names:
        assert np.array_equal(its[f], d["iters"][f]), f


@pytest.mark.parametrize("name", gu.ALIGN_CASES)
def test_golden_correspondences_are_consistent(oracle, name):
    d = gu.load(name)
    n_pairs = len(d["init_xyt"])
    for p in range(n_pairs):
        n = int(d["corr_n"][p])
        fi, mi = d["corr_fixed_idx"][p, :n], d["corr_moving_idx"][p, :n]
        fsrc, msrc = d["fixed_source_idx"][p], d["moving_source_idx"][p]
        # every correspondence pairs the two winners of one column, in ascending column order
        cols = [int(np.flatnonzero(fsrc == i)[0]) for i in fi]
        assert cols == sorted(cols)
        assert all(msrc[c] == j for c, j in zip(cols, mi))
        assert (d["corr_fixed_idx"][p, n:] == -1).all()
        # first-iteration n_corr of the aligner equals the finder's count at the initial guess
        assert d["iters"]["n_corr"][p, 0] == n


def test_oracle_reproduces_demo_scene_projection(oracle):
    d = gu.load("demo_scene_projection")
    prm = gu.make_params(oracle.default_params, d)
    img = oracle.project(prm, d["camera_pose"], d["scene"])
    assert np.array_equal(img["source_idx"], d["source_idx"])
    assert np.array_equal(gu.bits(img["depth"]), gu.bits(d["depth"]))


def test_oracle_reproduces_multi_slice_golden(oracle):
    """MULTI.json-shaped aligner: two laser slices with their own sensor_in_robot + the odometry prior"""
    from test_oracle_multi import sets_of  # noqa: F401  (same slice recipe as tools/make_golden.py)
    d = gu.load_raw("multi_721_mu")
    base = dict(canvas_cols=721, max_iterations=10, min_num_correspondences=5, with_sensor=1, point_distance=0.5)
    sl = [oracle.default_params(normal_cos=0.9, cauchy_chi_threshold=0.01,
                                sensor_in_robot=tuple(float(v) for v in d["sensors"][0]), **base),
          oracle.default_params(normal_cos=0.8, cauchy_chi_threshold=-1.0,
                                sensor_in_robot=tuple(float(v) for v in d["sensors"][1]), **base)]
    fixed = [(d["fixed_pts_0"], d["fixed_off"]), (d["fixed_pts_1"], d["fixed_off"])]
    moving = [(d["moving_pts"], d["moving_off"])] * 2
    res, its = oracle.align_multi_batch(sl, fixed, moving, d["init_xyt"], prior=oracle.make_prior(d["prior_info"]),
                                        prior_z=d["odom_xyt"])
    assert res.tobytes() == d["results"].tobytes() and its.tobytes() == d["iters"].tobytes()
    res, its = oracle.align_multi_batch(sl, fixed, moving, d["init_xyt"])
    assert res.tobytes() == d["results_no_prior"].tobytes() and its.tobytes() == d["iters_no_prior"].tobytes()
    assert (d["results"]["status"] == 0).all()
    assert np.abs(np.stack([d["results"][k] for k in ("x", "y", "theta")], 1) - d["gt_xyt"]).max() < 5e-3


def test_oracle_reproduces_tracker_rows_golden(oracle):
    """pre-processor -> clipper -> merger fixture (tools/make_golden.py --mapping-only), LASER_0.json values"""
    d = gu.load_raw("tracker_721_l0")
    kw = dict(angle_min=float(d["angles"][0]), angle_max=float(d["angles"][1]))
    sp_vox, sp_full = oracle.default_scan_params(**kw), oracle.default_scan_params(voxelize_resolution=0.0, **kw)
    prm = oracle.default_params(canvas_cols=721)
    for k in range(3):
        meas = oracle.preprocess_scan(sp_vox, d["fixed_ranges"][k])
        scene = oracle.preprocess_scan(sp_full, d["moving_ranges"][k])
        assert np.array_equal(gu.bits(meas), gu.bits(d[f"meas_{k}"]))
        assert np.array_equal(gu.bits(scene), gu.bits(d[f"scene_{k}"]))
        clip = oracle.clip_scene(prm, scene, d["robot_in_local_map"][k], d["sensor_in_robot"])
        assert np.array_equal(gu.bits(clip), gu.bits(d[f"clip_{k}"]))
        merged, counters = oracle.merge(prm, 0.2, scene, meas, d["gt_xyt"][k])
        assert np.array_equal(gu.bits(merged), gu.bits(d[f"merged_{k}"]))
        assert np.array_equal(counters, d[f"merge_counters_{k}"])
        assert 50 < len(meas) <= len(scene) <= 721 and counters[1] > 0       # voxelisation shrinks, merging merges
    assert len(d["synthetic_fixture_cloud"]) == 100                        # the reference's own pinned count
