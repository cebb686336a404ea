"""Poses at the boundary (CPU part): the reference hands Isometry2f objects to its modules
(correspondence_finder_projective_2d.cpp:40,47, scene_clipper_projective_2d.cpp:22-32, merger_projective_2d.cpp:19-22);
a caller's accumulated isometry is NOT v2t(t2v(T)) in binary32, so the C ABI takes the matrix form (tx, ty, c, s)
verbatim (LS2D_POSE_ISO).  Here: the reference's own compiled sources driven with accumulated isometries agree with the
oracle bit for bit, and the (x, y, theta) round trip measurably does not -- which is why the format exists."""
import ctypes as C

import numpy as np

from srrg2_laser_slam_2d_b200.synthetic import make_scan_pairs


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def accumulated(oracle, seed, xyt, steps=20):
    """an Isometry2f near v2t(xyt) built the way a tracker builds its pose: a product of `steps` increments"""
    rng = np.random.default_rng(seed)
    inc = np.tile(np.asarray(xyt, np.float64) / steps, (steps, 1)) + rng.normal(0.0, 2e-3, (steps, 3))
    return oracle.accumulate(inc.astype(np.float32))


def roundtrip(oracle, T):
    xyt = np.zeros(3, np.float32)
    oracle.lib().orc_t2v(T, xyt.ctypes.data)
    return oracle.v2t(*[float(v) for v in xyt])


def test_xyt_round_trip_is_not_the_identity_on_accumulated_isometries(oracle):
    changed, depth_changed, n_depth = 0, 0, 0
    sp = make_scan_pairs(32, n_beams=1081, seed=5)
    prm = oracle.default_params(canvas_cols=1081)
    for k in range(32):
        T = accumulated(oracle, k, (0.3, -0.2, 0.4))
        R = roundtrip(oracle, T)
        changed += (T.c, T.s) != (R.c, R.s)
        cloud = sp.moving_pts[sp.moving_off[k]:sp.moving_off[k + 1]]
        a, b = oracle.project(prm, T, cloud), oracle.project(prm, R, cloud)
        hit = (a["source_idx"] >= 0) & (b["source_idx"] >= 0)
        depth_changed += int((_bits(a["depth"][hit]) != _bits(b["depth"][hit])).sum())
        n_depth += int(hit.sum())
    assert changed >= 16             # most accumulated rotations are not representable as (cosf, sinf) of an angle
    assert depth_changed > n_depth // 4   # and the projector sees it: depth bits move in a large share of the columns


def test_reference_sources_with_accumulated_isometries_match_the_oracle(oracle, ref):
    """finder, clipper and merger of the reference's own compiled sources, fed Isometry2f CONTENT (tx, ty, c, s)"""
    prm = oracle.default_params(canvas_cols=1081, normal_cos=0.9)
    sp = make_scan_pairs(12, n_beams=1081, seed=21, motion_xy=0.3, motion_theta=0.15)
    for p in range(12):
        f = sp.fixed_pts[sp.fixed_off[p]:sp.fixed_off[p + 1]]
        m = sp.moving_pts[sp.moving_off[p]:sp.moving_off[p + 1]]
        T = accumulated(oracle, 100 + p, sp.gt_xyt[p])
        iso4 = oracle.iso_array([T])
        fi, mi = np.full(1081, -7, np.int32), np.full(1081, -7, np.int32)
        k = ref.ref_find_correspondences_iso(C.byref(prm), _p(f), len(f), _p(m), len(m), _p(iso4), 1, _p(fi), _p(mi))
        ofi, omi, _, _ = oracle.find_correspondences(prm, f, m, T)
        assert k == len(ofi) > 100 and np.array_equal(fi[:k], ofi) and np.array_equal(mi[:k], omi)
        # clipper: robot_in_local_map and sensor_in_robot both accumulated
        S = accumulated(oracle, 200 + p, (0.2, 0.2, 0.1), steps=5)
        s4 = oracle.iso_array([S])
        scene = np.concatenate([f, m])
        want = oracle.clip_scene(prm, scene, T, S)
        out = np.zeros((1081, 4), np.float32)
        k = ref.ref_clip_iso(C.byref(prm), _p(scene), len(scene), _p(iso4), _p(s4), 0.0, _p(out))
        assert k == len(want) > 100 and np.array_equal(_bits(out[:k]), _bits(want))
        # merger
        want, _ = oracle.merge(prm, 0.2, f, m, T)
        buf = np.zeros((len(f) + 1081, 4), np.float32)
        buf[:len(f)] = f
        k = ref.ref_merge_iso(C.byref(prm), 0.2, _p(buf), len(f), _p(m), len(m), _p(iso4))
        assert k == len(want) and np.array_equal(_bits(buf[:k]), _bits(want))
