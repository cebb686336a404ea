"""GPU parity of the multi-slice aligner (ls2d_align_multi: several laser slices + odometry prior in one 3x3
system; MULTI.json:700-730, LASER_0.json:502-506) against the CPU oracle, through the C ABI."""
import math

import numpy as np
import pytest

import golden_util as gu
from srrg2_laser_slam_2d_b200 import default_params
from srrg2_laser_slam_2d_b200._abi import LS2D_FIXED, LS2D_MOVING, make_prior
from srrg2_laser_slam_2d_b200.synthetic import make_multi_sensor_pairs, make_scan_pairs
from test_gpu_parity import INT_FIELDS, assert_bit_exact, tolerance_rate

pytestmark = pytest.mark.gpu

from srrg2_laser_slam_2d_b200._abi import multi_reduction_threads  # noqa: E402

MULTI_T = multi_reduction_threads()   # the general icp_multi_kernel's per-slice reduction shape (threads per pair)


def shape_of(slices, n_beams, shared=True):
    """per-slice reduction shape ls2d_align_multi runs for these slices on make_multi_sensor_pairs clouds (fixed: one
    scan of n_beams points per slice; moving: the shared local map of all sensors' scans).  Two slices with fixed
    clouds of <= 768 points, canvases below 768 columns and ONE shared moving set of <= 1536 points run the
    register-resident icp_multi2_kernel, everything else the general kernel."""
    return multi_reduction_threads(slices, n_beams, len(slices) * n_beams if len(slices) > 1 else n_beams, shared)
SENSORS = ((0.2, 0.05, 0.1), (-0.2, 0.0, math.pi))


def slices_for(factory, sensors, cols=721, **kw):
    base = dict(canvas_cols=cols, max_iterations=10, min_num_correspondences=5, with_sensor=1)
    base.update(kw)
    return [factory(normal_cos=0.9, cauchy_chi_threshold=0.01, sensor_in_robot=tuple(float(v) for v in sensors[0]), **base),
            factory(normal_cos=0.8, cauchy_chi_threshold=-1.0, sensor_in_robot=tuple(float(v) for v in sensors[1]), **base)]


def upload_multi(h, fixed0, fixed1, fixed_off, moving, moving_off):
    h.upload_clouds(0, fixed0, fixed_off)
    h.upload_clouds(2, fixed1, fixed_off)
    h.upload_clouds(1, moving, moving_off)


def test_golden_multi_slice_with_prior(handle_factory, oracle):
    d = gu.load_raw("multi_721_mu")
    h = handle_factory()
    upload_multi(h, d["fixed_pts_0"], d["fixed_pts_1"], d["fixed_off"], d["moving_pts"], d["moving_off"])
    sl = slices_for(default_params, d["sensors"], point_distance=0.5)
    osl = slices_for(oracle.default_params, d["sensors"], point_distance=0.5)
    fixed = [(d["fixed_pts_0"], d["fixed_off"]), (d["fixed_pts_1"], d["fixed_off"])]
    moving = [(d["moving_pts"], d["moving_off"])] * 2
    for with_prior, key in ((True, "results"), (False, "results_no_prior")):
        kw = dict(prior=make_prior(d["prior_info"]), prior_z=d["odom_xyt"]) if with_prior else {}
        okw = dict(prior=oracle.make_prior(d["prior_info"]), prior_z=d["odom_xyt"]) if with_prior else {}
        g, gi = h.align_multi(sl, [0, 2], [1, 1], d["init_xyt"], want_iters=True, **kw)
        assert shape_of(sl, 721) == 256 | 1 << 16 | 1 << 17  # the MULTI.json shape runs icp_multi2_kernel (fused sums)
        o, oi = oracle.align_multi_batch(osl, fixed, moving, d["init_xyt"], sum_mode=oracle.SUM_TREE,
                                         tree_threads=shape_of(sl, 721), **okw)
        assert_bit_exact(g, o, gi, oi)                      # kernel's summation order: every bit
        ref = d[key]                                        # frozen fixture: the reference's sequential order
        for f in INT_FIELDS:
            assert np.array_equal(g[f], ref[f]), f
        # north_star tolerances (1e-5 m, 1e-6 rad, 1e-4 relative chi2) as a rate over the fixture's pairs, and a hard
        # bound an order of magnitude above them
        rate, _ = tolerance_rate(g, ref)
        assert rate >= 0.75, rate
        assert np.abs(g["x"] - ref["x"]).max() <= 1e-4 and np.abs(g["theta"] - ref["theta"]).max() <= 1e-5


def test_one_slice_equals_the_fused_single_slice_kernel(handle_factory, oracle):
    """n_slices = 1 without prior must reproduce ls2d_align_batch's decisions; sums follow the multi kernel's tree"""
    sp = make_scan_pairs(24, n_beams=721, seed=31)
    kw = dict(canvas_cols=721, normal_cos=0.9)
    h = handle_factory(default_params(**kw))
    h.upload_clouds(LS2D_FIXED, sp.fixed_pts, sp.fixed_off)
    h.upload_clouds(LS2D_MOVING, sp.moving_pts, sp.moving_off)
    single = h.align_batch(sp.init_xyt)
    g, gi = h.align_multi([default_params(**kw)], [0], [1], sp.init_xyt, want_iters=True)
    o, oi = oracle.align_batch(oracle.default_params(**kw), sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off,
                               sp.init_xyt, sum_mode=oracle.SUM_TREE, tree_threads=MULTI_T)
    assert_bit_exact(g, o, gi, oi)
    rate, same = tolerance_rate(g, single)
    assert same == 1.0 and rate >= 0.95


def test_seeded_batch_two_sensors_prior_ids_and_skips(handle_factory, oracle):
    msp = make_multi_sensor_pairs(48, sensors=SENSORS, n_beams=541, seed=33)
    h = handle_factory()
    upload_multi(h, msp.fixed_pts[0], msp.fixed_pts[1], msp.fixed_off[0], msp.moving_pts, msp.moving_off)
    fixed = [(msp.fixed_pts[s], msp.fixed_off[s]) for s in range(2)]
    moving = [(msp.moving_pts, msp.moving_off)] * 2
    info = np.array([[50.0, 5.0, 1.0], [5.0, 60.0, -2.0], [1.0, -2.0, 200.0]], np.float32)
    rng = np.random.default_rng(0)
    fid = rng.integers(0, 48, 64).astype(np.int32)      # arbitrary (fixed, moving) pairings, more pairs than clouds
    mid = fid.copy()
    init = msp.init_xyt[fid]
    z = msp.odom_xyt[fid]
    for variant in range(3):
        sl = slices_for(default_params, msp.sensors, cols=541)
        osl = slices_for(oracle.default_params, msp.sensors, cols=541)
        pk, opk = {}, {}
        if variant == 0:        # full information matrix
            pk, opk = dict(prior=make_prior(info), prior_z=z), dict(prior=oracle.make_prior(info), prior_z=z)
        elif variant == 1:      # robustified prior, laser_1 can never contribute (D14)
            pk = dict(prior=make_prior(info, 0.001), prior_z=z)
            opk = dict(prior=oracle.make_prior(info, 0.001), prior_z=z)
            sl[1].min_num_correspondences = osl[1].min_num_correspondences = 100000
        else:                   # nobody contributes
            sl[0].min_num_correspondences = osl[0].min_num_correspondences = 100000
            sl[1].min_num_correspondences = osl[1].min_num_correspondences = 100000
        g, gi = h.align_multi(sl, [0, 2], [1, 1], init, fixed_id=fid, moving_id=mid, want_iters=True, **pk)
        o, oi = oracle.align_multi_batch(osl, fixed, moving, init, fixed_id=fid, moving_id=mid,
                                         sum_mode=oracle.SUM_TREE, tree_threads=shape_of(sl, 541), **opk)
        assert_bit_exact(g, o, gi, oi)
        seq, _ = oracle.align_multi_batch(osl, fixed, moving, init, fixed_id=fid, moving_id=mid, **opk)
        rate, same = tolerance_rate(g, seq)
        assert same == 1.0 and rate >= 0.95, (variant, rate, same)
        if variant == 2:
            assert (g["status"] == 1).all() and (g["n_corr"] > 0).all()
    err = np.abs(np.stack([g["x"], g["y"], g["theta"]], 1))
    assert err.max() == 0.0                                 # variant 2 never moved the estimate


def test_multi_argument_errors(handle_factory):
    from srrg2_laser_slam_2d_b200 import Ls2dError
    h = handle_factory()
    p = default_params()
    with pytest.raises(Ls2dError):                          # sets never uploaded
        h.align_multi([p], [0], [1], np.zeros((1, 3), np.float32))
    sp = make_scan_pairs(2, n_beams=181, seed=1)
    h.upload_clouds(0, sp.fixed_pts, sp.fixed_off)
    h.upload_clouds(1, sp.moving_pts, sp.moving_off)
    with pytest.raises(Ls2dError):                          # set id out of range
        h.align_multi([p], [9], [1], np.zeros((1, 3), np.float32))
    with pytest.raises(Ls2dError):                          # cloud id out of range
        h.align_multi([p], [0], [1], np.zeros((1, 3), np.float32), fixed_id=[5], moving_id=[0])
    with pytest.raises(Ls2dError):                          # a prior needs its measurements
        h.align_multi([p], [0], [1], np.zeros((1, 3), np.float32), prior=make_prior(np.eye(3)))
    assert len(h.align_multi([p], [0], [1], np.zeros((0, 3), np.float32))) == 0


@pytest.mark.parametrize("n_beams,cols,n_slices", [(1081, 1081, 2), (541, 900, 2), (361, 361, 3)])
def test_shapes_the_general_multi_kernel_runs(handle_factory, oracle, n_beams, cols, n_slices):
    """clouds above 768 points, canvases of 768 columns and more, or a third slice: icp_multi_kernel (stash in shared
    memory), bit for bit like the register-resident kernel on its shapes"""
    msp = make_multi_sensor_pairs(12, sensors=SENSORS, n_beams=n_beams, seed=35)
    h = handle_factory()
    upload_multi(h, msp.fixed_pts[0], msp.fixed_pts[1], msp.fixed_off[0], msp.moving_pts, msp.moving_off)
    sl = slices_for(default_params, msp.sensors, cols=cols)
    osl = slices_for(oracle.default_params, msp.sensors, cols=cols)
    fsets, msets = [0, 2], [1, 1]
    fixed = [(msp.fixed_pts[s], msp.fixed_off[s]) for s in range(2)]
    moving = [(msp.moving_pts, msp.moving_off)] * 2
    if n_slices == 3:                                       # a third slice looking at slice 0's clouds with other gates
        sl.append(default_params(canvas_cols=cols, normal_cos=0.95, point_distance=0.3, max_iterations=10))
        osl.append(oracle.default_params(canvas_cols=cols, normal_cos=0.95, point_distance=0.3, max_iterations=10))
        fsets, msets = fsets + [0], msets + [1]
        fixed, moving = fixed + [fixed[0]], moving + [moving[0]]
    assert multi_reduction_threads(sl, n_beams, 2 * n_beams, True) == MULTI_T
    z = msp.odom_xyt
    info = (100.0, 0.0, 0.0, 100.0, 0.0, 400.0)
    g, gi = h.align_multi(sl, fsets, msets, msp.init_xyt, prior=make_prior(info), prior_z=z, want_iters=True)
    o, oi = oracle.align_multi_batch(osl, fixed, moving, msp.init_xyt, prior=oracle.make_prior(info), prior_z=z,
                                     sum_mode=oracle.SUM_TREE, tree_threads=MULTI_T)
    assert_bit_exact(g, o, gi, oi)
    assert (g["status"] == 0).all()


def test_two_slices_with_moving_sets_of_their_own_run_the_general_kernel(handle_factory, oracle):
    msp = make_multi_sensor_pairs(12, sensors=SENSORS, n_beams=541, seed=36)
    h = handle_factory()
    upload_multi(h, msp.fixed_pts[0], msp.fixed_pts[1], msp.fixed_off[0], msp.moving_pts, msp.moving_off)
    h.upload_clouds(3, msp.moving_pts, msp.moving_off)      # the same clouds, but another set
    sl = slices_for(default_params, msp.sensors, cols=541)
    osl = slices_for(oracle.default_params, msp.sensors, cols=541)
    fixed = [(msp.fixed_pts[s], msp.fixed_off[s]) for s in range(2)]
    moving = [(msp.moving_pts, msp.moving_off)] * 2
    assert shape_of(sl, 541, shared=False) == MULTI_T and shape_of(sl, 541) == 256 | 1 << 16 | 1 << 17
    g, gi = h.align_multi(sl, [0, 2], [1, 3], msp.init_xyt, want_iters=True)
    o, oi = oracle.align_multi_batch(osl, fixed, moving, msp.init_xyt, sum_mode=oracle.SUM_TREE, tree_threads=MULTI_T)
    assert_bit_exact(g, o, gi, oi)
    g2, _ = h.align_multi(sl, [0, 2], [1, 1], msp.init_xyt, want_iters=True)     # shared: the register-resident kernel
    rate, same = tolerance_rate(g2, g)
    assert same == 1.0 and rate >= 0.9
