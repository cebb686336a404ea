"""Matrix-form poses at the C ABI (LS2D_POSE_ISO), on the device.

A caller of the reference holds Isometry2f objects; their (c, s) after a few compositions are not the cosf / sinf of any
angle, so a (x, y, theta) boundary would hand the device other bits than the reference's modules see.  With
LS2D_POSE_ISO the device uses the caller's isometry verbatim: the finder, clipper and merger of the reference's OWN
compiled sources (oracle/_ref), fed 20-step accumulated isometries, and the kernels agree bit for bit -- indices,
depths and clouds -- and the aligner started from such isometries follows the oracle's trajectory bit for bit."""
import ctypes as C

import numpy as np
import pytest

from srrg2_laser_slam_2d_b200 import default_params
from srrg2_laser_slam_2d_b200._abi import LS2D_FIXED, LS2D_MOVING, POSE_ISO, POSE_XYT, reduction_threads
from srrg2_laser_slam_2d_b200.synthetic import make_scan_pairs
from test_pose_formats import accumulated, roundtrip

pytestmark = pytest.mark.gpu


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("cols,n_beams", [(1081, 1081), (721, 721)])
def test_accumulated_isometries_reference_sources_vs_device(handle_factory, oracle, ref, cols, n_beams):
    kw = dict(canvas_cols=cols, normal_cos=0.9)
    prm = oracle.default_params(**kw)
    h = handle_factory(default_params(**kw))
    sp = make_scan_pairs(16, n_beams=n_beams, seed=31 + cols, motion_xy=0.3, motion_theta=0.15)
    h.upload_clouds(LS2D_FIXED, sp.fixed_pts, sp.fixed_off)
    h.upload_clouds(LS2D_MOVING, sp.moving_pts, sp.moving_off)
    n_flips = 0
    for p in range(16):
        f = sp.fixed_pts[sp.fixed_off[p]:sp.fixed_off[p + 1]]
        m = sp.moving_pts[sp.moving_off[p]:sp.moving_off[p + 1]]
        T = accumulated(oracle, 300 + p, sp.gt_xyt[p])
        iso4 = oracle.iso_array([T])
        # ---- finder: the reference's compiled CorrespondenceFinderProjective2f vs ls2d_find_correspondences
        fi, mi = np.full(cols, -7, np.int32), np.full(cols, -7, np.int32)
        k = ref.ref_find_correspondences_iso(C.byref(prm), _p(f), len(f), _p(m), len(m), _p(iso4), 1, _p(fi), _p(mi))
        gfi, gmi = h.find_correspondences(p, p, iso4[0])
        assert h.pose_format == POSE_ISO
        assert k == len(gfi) > 50 and np.array_equal(fi[:k], gfi) and np.array_equal(mi[:k], gmi)
        # ---- projector: pixel indices and depth bits of the moving image, camera = T^-1 as an isometry
        cam = oracle.iso_array([oracle.inverse(T)])[0]
        img = oracle.project(prm, cam, m)
        idx, depth = h.project(LS2D_MOVING, p, cam)
        assert np.array_equal(idx, img["source_idx"]) and np.array_equal(_bits(depth), _bits(img["depth"]))
        # the (x, y, theta) boundary would NOT have given these bits
        xyt = np.zeros(3, np.float32)
        oracle.lib().orc_t2v(oracle.inverse(T), xyt.ctypes.data)
        idx3, depth3 = h.project(LS2D_MOVING, p, xyt)
        assert h.pose_format == POSE_XYT
        n_flips += int((_bits(depth3) != _bits(depth)).sum())
        # ---- clipper (robot_in_local_map and sensor_in_robot both accumulated) and its voxelize branch
        S = accumulated(oracle, 400 + p, (0.2, 0.2, 0.1), steps=5)
        s4 = oracle.iso_array([S])[0]
        scene = np.concatenate([f, m])
        h.upload_clouds(2, scene, np.array([0, len(scene)], np.int32))
        for voxel in (0.0, 0.05):
            out = np.zeros((cols, 4), np.float32)
            k = ref.ref_clip_iso(C.byref(prm), _p(scene), len(scene), _p(iso4), _p(s4), voxel, _p(out))
            got = h.clip_scenes(2, [0], iso4, s4, voxelize_resolution=voxel)[0]
            assert k == len(got) > 50 and np.array_equal(_bits(out[:k]), _bits(got))
        # ---- merger
        buf = np.zeros((len(f) + cols, 4), np.float32)
        buf[:len(f)] = f
        k = ref.ref_merge_iso(C.byref(prm), 0.2, _p(buf), len(f), _p(m), len(m), _p(iso4))
        got, _ = h.merge_scene(f, m, iso4[0], 0.2)
        assert k == len(got) and np.array_equal(_bits(buf[:k]), _bits(got))
    assert n_flips > 100       # the round trip does change depth bits: the reason the format exists


def test_aligner_from_accumulated_isometries_is_bit_exact(handle_factory, oracle):
    """initial guesses AND sensor_in_robot as accumulated isometries; trajectory, final (c, s) and H bit for bit"""
    from test_gpu_parity import assert_bit_exact
    sp = make_scan_pairs(96, n_beams=1081, seed=77)
    S = oracle.iso_array([accumulated(oracle, 9, (0.2, 0.2, 0.1), steps=7)])[0]
    init = oracle.iso_array([accumulated(oracle, 500 + p, sp.init_xyt[p] + np.float32([0.02, -0.01, 0.01]))
                             for p in range(96)])
    for with_sensor in (0, 2):
        kw = dict(canvas_cols=1081, normal_cos=0.9, with_sensor=with_sensor)
        gp, op = default_params(**kw), oracle.default_params(**kw)
        if with_sensor:
            from srrg2_laser_slam_2d_b200._abi import set_sensor
            set_sensor(gp, S), oracle.set_sensor(op, S)
            assert gp.with_sensor == 2 and op.with_sensor == 2
        h = handle_factory(gp)
        h.upload_clouds(LS2D_FIXED, sp.fixed_pts, sp.fixed_off)
        h.upload_clouds(LS2D_MOVING, sp.moving_pts, sp.moving_off)
        g, gi = h.align_batch(init, want_iters=True)
        o, oi = oracle.align_batch(op, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, init,
                                   sum_mode=oracle.SUM_TREE, tree_threads=reduction_threads(1081, params=gp),
                                   n_threads=oracle.max_threads())
        assert_bit_exact(g, o, gi, oi)
        assert (g["status"] == 0).mean() > 0.9
        # the reported isometry is the state itself: theta is its atan2f
        assert np.array_equal(_bits(g["theta"]), _bits(oracle.libm_atan2f(g["s"], g["c"])))


def test_both_formats_agree_when_the_pose_is_a_v2t(handle_factory, oracle):
    """(x, y, theta) and v2t(x, y, theta) handed over as (tx, ty, c, s) are the same pose to the device"""
    sp = make_scan_pairs(48, n_beams=721, seed=78)
    kw = dict(canvas_cols=721, normal_cos=0.9)
    h = handle_factory(default_params(**kw))
    h.upload_clouds(LS2D_FIXED, sp.fixed_pts, sp.fixed_off)
    h.upload_clouds(LS2D_MOVING, sp.moving_pts, sp.moving_off)
    a = h.align_batch(sp.init_xyt)
    iso4 = oracle.iso_array([oracle.v2t(*[float(v) for v in x]) for x in sp.init_xyt])
    b = h.align_batch(iso4)
    assert a.tobytes() == b.tobytes()
    best_a = h.verify(0, np.arange(8, dtype=np.int32), np.tile(sp.init_xyt[:8, None, :], (1, 2, 1)),
                      __import__("srrg2_laser_slam_2d_b200").Gates(50, 1.0, 0.1))
    best_b = h.verify(0, np.arange(8, dtype=np.int32), np.tile(iso4[:8, None, :], (1, 2, 1)),
                      __import__("srrg2_laser_slam_2d_b200").Gates(50, 1.0, 0.1))
    assert best_a.tobytes() == best_b.tobytes()


def test_classify_correspondences_matches_the_factor_error(handle_factory, oracle):
    """MultiAligner2D.keep_only_inlier_correspondences: inlier = chi < cauchy_chi_threshold at the estimate"""
    sp = make_scan_pairs(8, n_beams=721, seed=79, motion_xy=0.2)
    for factor in (0, 1):
        kw = dict(canvas_cols=721, normal_cos=0.8, cauchy_chi_threshold=0.01, factor=factor)
        h = handle_factory(default_params(**kw))
        prm = oracle.default_params(**kw)
        h.upload_clouds(LS2D_FIXED, sp.fixed_pts, sp.fixed_off)
        h.upload_clouds(LS2D_MOVING, sp.moving_pts, sp.moving_off)
        n_in = n_out = 0
        for p in range(8):
            f = sp.fixed_pts[sp.fixed_off[p]:sp.fixed_off[p + 1]]
            m = sp.moving_pts[sp.moving_off[p]:sp.moving_off[p + 1]]
            X = sp.init_xyt[p]
            fi, mi = h.find_correspondences(p, p, X)
            got = h.classify_correspondences(p, p, X, fi, mi)
            want = []
            for a, b in zip(fi, mi):
                e, _ = oracle.error_and_jacobian(prm, X, f[a], m[b])
                if factor == 0:
                    chi = np.float32(np.float32(np.float32(e[0] * e[0]) + np.float32(e[1] * e[1])) + np.float32(e[2] * e[2]))
                else:
                    T = oracle.v2t(*[float(v) for v in X])
                    px = np.float32(np.float32(np.float32(T.c * m[b][0]) + np.float32(np.float32(-T.s) * m[b][1])) + np.float32(T.tx))
                    py = np.float32(np.float32(np.float32(T.s * m[b][0]) + np.float32(T.c * m[b][1])) + np.float32(T.ty))
                    dx, dy = np.float32(px - f[a][0]), np.float32(py - f[a][1])
                    chi = np.float32(np.float32(dx * dx) + np.float32(dy * dy))
                want.append(chi < np.float32(0.01))
            assert np.array_equal(got, np.array(want))
            n_in += int(got.sum())
            n_out += int((~got).sum())
        assert n_in > 100 and n_out > 10
