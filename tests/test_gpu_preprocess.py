"""GPU parity of the raw-scan pre-processor (RawDataPreprocessorProjective2D,
R/sensor_processing/raw_data_preprocessor_projective_2d.cpp) through the C ABI: bit-exact against the oracle."""
import numpy as np
import pytest

import golden_util as gu
from srrg2_laser_slam_2d_b200 import default_params
from srrg2_laser_slam_2d_b200._abi import LS2D_FIXED, LS2D_MOVING, default_scan_params
from srrg2_laser_slam_2d_b200.synthetic import make_raw_scans

pytestmark = pytest.mark.gpu


def both_params(oracle, **kw):
    return default_scan_params(**kw), oracle.default_scan_params(**kw)


def check(h, oracle, kw, ranges):
    sp, osp = both_params(oracle, **kw)
    got, cnt = h.preprocess_scans(sp, ranges)
    ref, rcnt = oracle.preprocess_scans(osp, ranges, n_threads=4)
    assert np.array_equal(cnt, rcnt)
    for s in range(len(ranges)):
        assert np.array_equal(gu.bits(got[s, :cnt[s]]), gu.bits(ref[s, :cnt[s]])), s
    return cnt


def test_reference_synthetic_fixture(handle_factory, oracle):
    # /root/reference/srrg2_laser_slam_2d/tests/fixtures.hpp:38-47 + test_measurement_adaptor.cpp:36 -> 100 points
    h = handle_factory(default_params())
    kw = dict(angle_min=-1.0, angle_max=1.0, msg_range_min=0.0, msg_range_max=1000.0, range_min=0.0,
              range_max=1000.0, voxelize_resolution=0.01)
    cnt = check(h, oracle, kw, np.full((1, 100), 1.0, np.float32))
    assert cnt[0] == 100


@pytest.mark.parametrize("n_beams,res", [(1081, 0.02), (1081, 0.0), (721, 0.02), (721, 0.05), (360, 0.0), (2048, 0.1)])
def test_preprocess_bit_exact(handle_factory, oracle, n_beams, res):
    raw = make_raw_scans(24, n_beams=n_beams, seed=91 + n_beams)
    h = handle_factory(default_params())
    kw = dict(angle_min=raw.angle_min, angle_max=raw.angle_max, voxelize_resolution=res)
    cnt = check(h, oracle, kw, np.concatenate([raw.fixed_ranges, raw.moving_ranges]))
    assert cnt.min() > 20 and cnt.max() <= n_beams


def test_preprocess_edge_cases(handle_factory, oracle):
    h = handle_factory(default_params())
    rng = np.random.default_rng(3)
    n = 300
    ranges = np.stack([
        np.full(n, 65.0, np.float32),                                   # nothing in range
        np.full(n, 25.0, np.float32),                                   # too sparse for any normal
        np.where(np.arange(n) % 7 == 0, 2.0, 65.0).astype(np.float32),  # isolated beams
        np.full(n, 0.0, np.float32),                                    # all points at the origin (range_min 0 accepts them)
        rng.uniform(0.5, 0.6, n).astype(np.float32),
        np.full(n, 1.0, np.float32),
    ])
    kw = dict(angle_min=-2.0, angle_max=2.0)
    cnt = check(h, oracle, kw, ranges)
    assert cnt[0] == 0 and cnt[1] == 0 and cnt[2] == 0
    check(h, oracle, dict(kw, voxelize_resolution=0.0), ranges)
    check(h, oracle, dict(kw, normal_min_points=1, voxelize_resolution=0.3), ranges)
    got, c0 = h.preprocess_scans(default_scan_params(), np.zeros((0, 64), np.float32))
    assert len(c0) == 0


def test_raw_scans_to_resident_set_and_align(handle_factory, oracle):
    """raw ranges -> resident cloud sets -> aligner: same result as aligning the oracle's pre-processed clouds"""
    raw = make_raw_scans(16, n_beams=1081, seed=7)
    kw = dict(angle_min=raw.angle_min, angle_max=raw.angle_max, voxelize_resolution=0.0)
    sp, osp = both_params(oracle, **kw)
    akw = dict(canvas_cols=1081, normal_cos=0.9, max_iterations=10)
    h = handle_factory(default_params(**akw))
    h.preprocess_scans_to_set(LS2D_FIXED, sp, raw.fixed_ranges)
    h.preprocess_scans_to_set(LS2D_MOVING, sp, raw.moving_ranges)
    fpts, foff = h.download_clouds(LS2D_FIXED, 16, 16 * 1081)
    rf, cf = oracle.preprocess_scans(osp, raw.fixed_ranges)
    rm, cm = oracle.preprocess_scans(osp, raw.moving_ranges)
    assert np.array_equal(np.diff(foff), cf)
    ref_f = np.concatenate([rf[s, :cf[s]] for s in range(16)])
    ref_m = np.concatenate([rm[s, :cm[s]] for s in range(16)])
    assert np.array_equal(gu.bits(fpts), gu.bits(ref_f))
    got = h.align_batch(raw.init_xyt)
    from srrg2_laser_slam_2d_b200._abi import reduction_threads
    off_f = np.concatenate([[0], np.cumsum(cf)]).astype(np.int32)
    off_m = np.concatenate([[0], np.cumsum(cm)]).astype(np.int32)
    ref, _ = oracle.align_batch(oracle.default_params(**akw), ref_f, off_f, ref_m, off_m, raw.init_xyt,
                                sum_mode=oracle.SUM_TREE, tree_threads=reduction_threads(1081))
    for f in ("status", "n_corr", "n_inliers", "n_kernelized"):
        assert np.array_equal(got[f], ref[f]), f
    for f in ("x", "y", "theta", "chi_inliers"):
        assert np.array_equal(gu.bits(got[f]), gu.bits(ref[f])), f
    err = np.abs(np.stack([got["x"], got["y"], got["theta"]], 1) - raw.gt_xyt).max()
    assert err < 0.02


@pytest.mark.parametrize("res,sensor", [(0.02, None), (0.0, (0.1, -0.05, 0.2))])
def test_track_batch_matches_the_oracle_chain(handle_factory, oracle, res, sensor):
    """ls2d_track_batch = pre-process -> clip -> align with only ranges / ids / poses crossing the bus: identical to the
    oracle's RawDataPreprocessorProjective2D -> SceneClipperProjective2D -> MultiAligner2D chain."""
    from srrg2_laser_slam_2d_b200._abi import reduction_threads
    n, nb = 12, 1081
    raw = make_raw_scans(n, n_beams=nb, seed=23)
    kw = dict(angle_min=raw.angle_min, angle_max=raw.angle_max)
    sp, osp = both_params(oracle, voxelize_resolution=res, **kw)
    _, osp_map = both_params(oracle, voxelize_resolution=0.0, **kw)
    akw = dict(canvas_cols=1081, normal_cos=0.9, max_iterations=10)
    if sensor is not None:
        akw.update(with_sensor=1, sensor_in_robot=sensor)
    # local maps: the full-resolution cloud of the scan taken at P * delta, in that pose's frame
    maps, cm = oracle.preprocess_scans(osp_map, raw.moving_ranges)
    scenes = [maps[k, :cm[k]] for k in range(n)]
    off = np.concatenate([[0], np.cumsum(cm)]).astype(np.int32)
    rng = np.random.default_rng(1)
    ids = rng.permutation(n).astype(np.int32)
    robots = rng.uniform(-0.02, 0.02, (n, 3)).astype(np.float32)       # predicted robot_in_local_map
    ranges = raw.fixed_ranges[ids]
    h = handle_factory(default_params(**akw))
    h.upload_clouds(2, np.concatenate(scenes), off)
    got = h.track_batch(sp, ranges, 2, ids, robots)
    prm = oracle.default_params(**akw)
    sen = sensor if sensor is not None else (0.0, 0.0, 0.0)
    fix, mov = [], []
    for f in range(n):
        fix.append(oracle.preprocess_scan(osp, ranges[f]))
        mov.append(oracle.clip_scene(prm, scenes[ids[f]], robots[f], sen))
    foff = np.concatenate([[0], np.cumsum([len(c) for c in fix])]).astype(np.int32)
    moff = np.concatenate([[0], np.cumsum([len(c) for c in mov])]).astype(np.int32)
    mpts, moff_dev = h.download_clouds(LS2D_MOVING, n, n * 1081)
    assert np.array_equal(moff_dev, moff) and np.array_equal(gu.bits(mpts), gu.bits(np.concatenate(mov)))
    # the sets are sized by their upper bound (n_beams / canvas_cols): that fixes the kernel's reduction shape
    ref, _ = oracle.align_batch(prm, np.concatenate(fix), foff, np.concatenate(mov), moff, np.zeros((n, 3), np.float32),
                                sum_mode=oracle.SUM_TREE, tree_threads=reduction_threads(1081))
    for f in ("status", "n_corr", "n_inliers", "n_kernelized"):
        assert np.array_equal(got[f], ref[f]), f
    for f in ("x", "y", "theta", "chi_inliers"):
        assert np.array_equal(gu.bits(got[f]), gu.bits(ref[f])), f
    assert (got["status"] == 0).all()


@pytest.mark.parametrize("iso,sensor", [(False, None), (True, (0.15, -0.05, 0.1))])
def test_track_batch_in_chunks_equals_the_three_calls(handle_factory, iso, sensor):
    """2600 frames: ls2d_track_batch cuts the batch into chunks (uploads overlap the kernels, two compute lanes);
    results and the two cloud sets it leaves behind are those of pre-process -> clip -> align called one after the
    other on the whole batch."""
    n, nb = 2600, 721
    raw = make_raw_scans(64, n_beams=nb, seed=5)
    rng = np.random.default_rng(9)
    pick = rng.integers(0, 64, n)
    ranges = np.ascontiguousarray(raw.fixed_ranges[pick])
    ids = pick.astype(np.int32)                               # the local map of the pose the frame was taken near
    robots = rng.uniform(-0.02, 0.02, (n, 3)).astype(np.float32)
    init = rng.uniform(-0.01, 0.01, (n, 3)).astype(np.float32)
    kw = dict(angle_min=raw.angle_min, angle_max=raw.angle_max)
    sp = default_scan_params(voxelize_resolution=0.02, **kw)
    sp_map = default_scan_params(voxelize_resolution=0.0, **kw)
    akw = dict(canvas_cols=721, max_iterations=6)
    sen = np.zeros(3, np.float32)
    if sensor is not None:
        akw.update(with_sensor=1, sensor_in_robot=sensor)
        sen = np.float32(sensor)
    if iso:  # poses as (tx, ty, c, s): the handle takes the format from the arrays' last dimension
        to_iso = lambda p: np.stack([p[..., 0], p[..., 1], np.cos(p[..., 2]), np.sin(p[..., 2])], -1).astype(np.float32)
        robots, init, sen = to_iso(robots), to_iso(init), to_iso(sen)
    h = handle_factory(default_params(**akw))
    h.preprocess_scans_to_set(2, sp_map, raw.moving_ranges)
    got = h.track_batch(sp, ranges, 2, ids, robots, init).copy()
    fpts, foff = h.download_clouds(LS2D_FIXED, n, n * nb)
    mpts, moff = h.download_clouds(LS2D_MOVING, n, n * 721)
    h.preprocess_scans_to_set(LS2D_FIXED, sp, ranges)
    h.clip_scenes_to_set(2, ids, robots, sen, LS2D_MOVING)
    fpts2, foff2 = h.download_clouds(LS2D_FIXED, n, n * nb)
    mpts2, moff2 = h.download_clouds(LS2D_MOVING, n, n * 721)
    want = h.align_batch(init)
    assert np.array_equal(foff, foff2) and np.array_equal(moff, moff2)
    assert np.array_equal(gu.bits(fpts), gu.bits(fpts2)) and np.array_equal(gu.bits(mpts), gu.bits(mpts2))
    assert got.tobytes() == want.tobytes()
    assert (got["status"] == 0).mean() > 0.9


def test_voxelize_with_uneven_buckets_takes_the_sort(handle_factory, oracle):
    """a wall at x = const seen square on: hundreds of voxels share one ix, the buckets of the counting sort are too
    uneven and the segments go through the bitonic sort instead -- same cloud, bit for bit"""
    nb = 1081
    kw = dict(angle_min=-1.2, angle_max=1.2)
    az = np.linspace(-1.2, 1.2, nb, dtype=np.float64)
    rng = np.random.default_rng(2)
    scans = []
    for d in (2.0, 3.5, 1.2):
        r = d / np.cos(az)                                    # the wall x = d
        r[::97] = 0.0                                         # a few dropped beams
        r[520:560] = 29.0 + 0.01 * np.arange(40)              # a doorway: far returns stretch the x extent, so the
                                                              # 2048 buckets cannot also split the wall along y
        scans.append((r + rng.normal(0, 1e-4, nb)).astype(np.float32))
    ranges = np.stack(scans)
    h = handle_factory()
    for res in (0.02, 0.05):
        check(h, oracle, dict(voxelize_resolution=res, **kw), ranges)


def test_beam_count_limits(handle_factory, oracle):
    """n_beams <= 8192 without voxelisation; with it the scan has to fit one CTA's shared memory (<= 6000 beams),
    beyond that the call says so instead of falling back to anything"""
    from srrg2_laser_slam_2d_b200._abi import Ls2dError
    h = handle_factory()
    for nb, res in ((6000, 0.02), (8192, 0.0)):
        raw = make_raw_scans(3, n_beams=nb, seed=nb)
        check(h, oracle, dict(angle_min=raw.angle_min, angle_max=raw.angle_max, voxelize_resolution=res), raw.fixed_ranges)
    raw = make_raw_scans(1, n_beams=8192, seed=1)
    sp = default_scan_params(angle_min=raw.angle_min, angle_max=raw.angle_max, voxelize_resolution=0.02)
    with pytest.raises(Ls2dError):
        h.preprocess_scans(sp, raw.fixed_ranges)
