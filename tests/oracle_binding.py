"""ctypes binding of the CPU oracle (oracle/ls2d_oracle.h).  Test infrastructure: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "libls2d_oracle.so")


class OrcParams(C.Structure):
    _fields_ = [("canvas_cols", C.c_int32), ("angle_col_min", C.c_float), ("angle_col_max", C.c_float),
                ("range_min", C.c_float), ("range_max", C.c_float), ("point_distance", C.c_float),
                ("normal_cos", C.c_float), ("cauchy_chi_threshold", C.c_float), ("damping", C.c_float),
                ("max_iterations", C.c_int32), ("min_num_correspondences", C.c_int32),
                ("min_num_inliers", C.c_int32), ("with_sensor", C.c_int32),
                ("sensor_in_robot", C.c_float * 3), ("sensor_in_robot_cs", C.c_float * 2),
                ("factor", C.c_int32), ("algorithm", C.c_int32), ("lm_user_lambda_init", C.c_float),
                ("lm_tau", C.c_float), ("lm_step_low", C.c_float), ("lm_step_high", C.c_float),
                ("lm_iterations_max", C.c_int32), ("lm_variable_damping", C.c_int32),
                ("single_rounding_accumulation", C.c_int32), ("enable_inlier_only_runs", C.c_int32),
                ("keep_only_inlier_correspondences", C.c_int32), ("termination_epsilon", C.c_float)]


class OrcPrior(C.Structure):
    _fields_ = [("z", C.c_float * 4), ("z_is_iso", C.c_int32), ("information", C.c_float * 6),
                ("cauchy_chi_threshold", C.c_float)]


class OrcScanParams(C.Structure):
    _fields_ = [("angle_min", C.c_float), ("angle_max", C.c_float), ("msg_range_min", C.c_float),
                ("msg_range_max", C.c_float), ("range_min", C.c_float), ("range_max", C.c_float),
                ("voxelize_resolution", C.c_float), ("normal_point_distance", C.c_float),
                ("normal_min_points", C.c_int32)]


class OrcIso(C.Structure):
    _fields_ = [("tx", C.c_float), ("ty", C.c_float), ("c", C.c_float), ("s", C.c_float)]


class OrcPoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("nx", C.c_float), ("ny", C.c_float)]


CELL_DTYPE = np.dtype([("source_idx", "<i4"), ("depth", "<f4"), ("px", "<f4"), ("py", "<f4"),
                       ("nx", "<f4"), ("ny", "<f4")])
RESULT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("theta", "<f4"), ("chi_inliers", "<f4"),
                         ("chi_kernelized", "<f4"), ("n_inliers", "<i4"), ("n_kernelized", "<i4"),
                         ("n_corr", "<i4"), ("status", "<i4"), ("iterations", "<i4"), ("H", "<f4", (6,)),
                         ("c", "<f4"), ("s", "<f4"), ("lm_rejected", "<i4"), ("reserved", "<i4")])
ITER_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("theta", "<f4"), ("chi_inliers", "<f4"),
                       ("chi_kernelized", "<f4"), ("n_inliers", "<i4"), ("n_kernelized", "<i4"),
                       ("n_corr", "<i4"), ("c", "<f4"), ("s", "<f4")])
assert RESULT_DTYPE.itemsize == 80 and ITER_DTYPE.itemsize == 40 and CELL_DTYPE.itemsize == 24

SUM_SEQUENTIAL, SUM_TREE = 0, 1
_lib = None


def build() -> str:
    subprocess.run(["make", "-s", "-C", os.path.join(_ROOT, "oracle")], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
        L.orc_default_params.argtypes = [C.POINTER(OrcParams)]
        L.orc_v2t.argtypes, L.orc_v2t.restype = [f32, f32, f32], OrcIso
        L.orc_inverse.argtypes, L.orc_inverse.restype = [OrcIso], OrcIso
        L.orc_compose.argtypes, L.orc_compose.restype = [OrcIso, OrcIso], OrcIso
        L.orc_t2v.argtypes = [OrcIso, vp]
        L.orc_project.argtypes = [C.POINTER(OrcParams), OrcIso, vp, i32, vp]
        L.orc_find_correspondences.argtypes = [C.POINTER(OrcParams), vp, vp, i32, OrcIso, vp, vp, vp]
        L.orc_find_correspondences.restype = i32
        L.orc_error_and_jacobian.argtypes = [C.POINTER(OrcParams), OrcIso, OrcPoint, OrcPoint, vp, vp]
        L.orc_align.argtypes = [C.POINTER(OrcParams), vp, i32, vp, i32, vp, i32, i32, vp, vp]
        L.orc_align_iso.argtypes = [C.POINTER(OrcParams), vp, i32, vp, i32, OrcIso, i32, i32, vp, vp]
        L.orc_align_batch.argtypes = [C.POINTER(OrcParams), vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]
        L.orc_prior_error_and_jacobian.argtypes = [C.POINTER(OrcPrior), OrcIso, vp, vp]
        L.orc_align_multi_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]
        L.orc_best_of.argtypes = [vp, i32, i32, f32, f32]
        L.orc_best_of.restype = i32
        L.orc_accept.argtypes = [vp, i32, f32, f32]
        L.orc_accept.restype = i32
        L.orc_max_threads.restype = i32
        L.orc_clip_scene.argtypes, L.orc_clip_scene.restype = [C.POINTER(OrcParams), vp, i32, OrcIso, OrcIso, vp], i32
        L.orc_clip_scene_voxelized.argtypes = [C.POINTER(OrcParams), vp, i32, OrcIso, OrcIso, f32, vp]
        L.orc_clip_scene_voxelized.restype = i32
        L.orc_merge.argtypes = [C.POINTER(OrcParams), f32, vp, i32, vp, i32, OrcIso, vp]
        L.orc_merge.restype = i32
        L.orc_default_scan_params.argtypes = [C.POINTER(OrcScanParams)]
        L.orc_preprocess_scan.argtypes, L.orc_preprocess_scan.restype = [C.POINTER(OrcScanParams), vp, i32, vp], i32
        L.orc_preprocess_scans.argtypes = [C.POINTER(OrcScanParams), vp, i32, i32, i32, vp, vp]
        L.orc_libm_atan2f_n.argtypes = [vp, vp, vp, C.c_long]
        L.orc_libm_sincosf_n.argtypes = [vp, vp, vp, C.c_long]
        L.orc_libm_logf_n.argtypes = [vp, vp, C.c_long]
        L.orc_column_n.argtypes = [C.POINTER(OrcParams), vp, vp, vp, C.c_long]
        _lib = L
    return _lib


def default_params(**kw) -> OrcParams:
    p = OrcParams()
    lib().orc_default_params(C.byref(p))
    for k, v in kw.items():
        if k == "sensor_in_robot":
            set_sensor(p, v)
        elif k == "sensor_in_robot_cs":
            p.sensor_in_robot_cs = (C.c_float * 2)(*[float(x) for x in v])
        else:
            setattr(p, k, v)
    return p


def set_sensor(p, v):
    """sensor_in_robot of a params record: 3 values = (x, y, theta), 4 values = the isometry (tx, ty, c, s)"""
    v = [float(x) for x in v]
    if len(v) == 4:
        p.sensor_in_robot = (C.c_float * 3)(v[0], v[1], 0.0)
        p.sensor_in_robot_cs = (C.c_float * 2)(v[2], v[3])
        if p.with_sensor:
            p.with_sensor = 2
    else:
        p.sensor_in_robot = (C.c_float * 3)(*v)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def v2t(x, y, theta) -> OrcIso:
    return lib().orc_v2t(x, y, theta)


def as_iso(pose) -> OrcIso:
    """a pose in either wire format: 3 values = (x, y, theta) through geometry2d::v2t, 4 values = the Isometry2f
    content (tx, ty, c, s) used verbatim"""
    if isinstance(pose, OrcIso):
        return pose
    pose = [float(v) for v in np.asarray(pose, np.float32).ravel()]
    if len(pose) == 4:
        return OrcIso(*[np.float32(v) for v in pose])
    return v2t(*pose)


def iso_array(isos) -> np.ndarray:
    """[n, 4] float32 (tx, ty, c, s) from OrcIso records"""
    return np.array([[T.tx, T.ty, T.c, T.s] for T in isos], np.float32).reshape(-1, 4)


def compose(a, b) -> OrcIso:
    return lib().orc_compose(as_iso(a), as_iso(b))


def inverse(a) -> OrcIso:
    return lib().orc_inverse(as_iso(a))


def accumulate(steps_xyt) -> OrcIso:
    """product of v2t(step) over the rows of steps_xyt, the way a tracker accumulates its pose: an Isometry2f whose
    (c, s) are NOT the cosf / sinf of any angle"""
    T = v2t(0.0, 0.0, 0.0)
    for st in np.asarray(steps_xyt, np.float32).reshape(-1, 3):
        T = lib().orc_compose(T, v2t(*[float(v) for v in st]))
    return T


def project(prm: OrcParams, camera_pose_xyt, pts: np.ndarray) -> np.ndarray:
    """PointNormal2fProjectorPolar with setCameraPose(v2t(camera_pose_xyt)); returns canvas_cols cells."""
    pts = _f32(pts)
    img = np.zeros(prm.canvas_cols, CELL_DTYPE)
    lib().orc_project(C.byref(prm), as_iso(camera_pose_xyt), _ptr(pts), len(pts), _ptr(img))
    return img


def find_correspondences(prm: OrcParams, fixed: np.ndarray, moving: np.ndarray, local_map_in_sensor_xyt):
    """CorrespondenceFinderProjective2f::compute(); returns (fixed_idx, moving_idx, fixed_img, moving_img)."""
    fixed, moving = _f32(fixed), _f32(moving)
    fimg = project(prm, (0.0, 0.0, 0.0), fixed)
    mimg = np.zeros(prm.canvas_cols, CELL_DTYPE)
    fi = np.zeros(prm.canvas_cols, np.int32)
    mi = np.zeros(prm.canvas_cols, np.int32)
    k = lib().orc_find_correspondences(C.byref(prm), _ptr(fimg), _ptr(moving), len(moving),
                                       as_iso(local_map_in_sensor_xyt), _ptr(mimg), _ptr(fi), _ptr(mi))
    return fi[:k].copy(), mi[:k].copy(), fimg, mimg


def error_and_jacobian(prm: OrcParams, X_xyt, fixed_pt, moving_pt):
    e = np.zeros(3, np.float32)
    J = np.zeros(9, np.float32)
    lib().orc_error_and_jacobian(C.byref(prm), as_iso(X_xyt), OrcPoint(*[float(v) for v in fixed_pt]),
                                 OrcPoint(*[float(v) for v in moving_pt]), _ptr(e), _ptr(J))
    return e, J.reshape(3, 3)


def align(prm: OrcParams, fixed, moving, init_xyt, sum_mode=SUM_SEQUENTIAL, tree_threads=256):
    fixed, moving = _f32(fixed), _f32(moving)
    out = np.zeros(1, RESULT_DTYPE)
    its = np.zeros(prm.max_iterations * (2 if prm.enable_inlier_only_runs else 1), ITER_DTYPE)
    lib().orc_align_iso(C.byref(prm), _ptr(fixed), len(fixed), _ptr(moving), len(moving), as_iso(init_xyt),
                        sum_mode, tree_threads, _ptr(out), _ptr(its))
    return out[0], its


def align_batch(prm: OrcParams, fixed_pts, fixed_off, moving_pts, moving_off, init_xyt, fixed_id=None,
                moving_id=None, sum_mode=SUM_SEQUENTIAL, tree_threads=256, n_threads=1, want_iters=True):
    fixed_pts, moving_pts, init = _f32(fixed_pts), _f32(moving_pts), _f32(init_xyt)
    fixed_off, moving_off = _i32(fixed_off), _i32(moving_off)
    fixed_id, moving_id = _i32(fixed_id), _i32(moving_id)
    init = init.reshape(len(init), -1)  # [n, 3] (x, y, theta) or [n, 4] (tx, ty, c, s)
    n = len(init)
    out = np.zeros(n, RESULT_DTYPE)
    n_it = prm.max_iterations * (2 if prm.enable_inlier_only_runs else 1)
    its = np.zeros((n, n_it), ITER_DTYPE) if want_iters else None
    lib().orc_align_batch(C.byref(prm), _ptr(fixed_pts), _ptr(fixed_off), _ptr(moving_pts), _ptr(moving_off),
                          _ptr(fixed_id), _ptr(moving_id), _ptr(init), init.shape[1], n, sum_mode, tree_threads,
                          n_threads, _ptr(out), _ptr(its))
    return out, its


def make_prior(information, cauchy_chi_threshold=-1.0, z=(0.0, 0.0, 0.0)) -> OrcPrior:
    """information: 6 floats (upper triangle O00 O01 O02 O11 O12 O22) or a 3x3 symmetric matrix"""
    info = np.asarray(information, np.float32)
    if info.shape == (3, 3):
        info = info[np.triu_indices(3)]
    pr = OrcPrior()
    z = [float(v) for v in z]
    pr.z = (C.c_float * 4)(*(z + [0.0] * (4 - len(z))))
    pr.z_is_iso = 1 if len(z) == 4 else 0
    pr.information = (C.c_float * 6)(*[float(v) for v in info])
    pr.cauchy_chi_threshold = cauchy_chi_threshold
    return pr


def prior_error_and_jacobian(prior: OrcPrior, X_xyt):
    e, J = np.zeros(3, np.float32), np.zeros(9, np.float32)
    lib().orc_prior_error_and_jacobian(C.byref(prior), as_iso(X_xyt), _ptr(e), _ptr(J))
    return e, J.reshape(3, 3)


def align_multi_batch(slices, fixed_sets, moving_sets, init_xyt, prior=None, prior_z=None, fixed_id=None,
                      moving_id=None, sum_mode=SUM_SEQUENTIAL, tree_threads=256, n_threads=1, want_iters=True):
    """MultiAligner2D with several laser slices (+ optional odometry prior).  slices: list of OrcParams;
    fixed_sets / moving_sets: per slice a (points [n, 4], offsets) CSR pair; prior_z: [n_pairs, 3]."""
    n_s = len(slices)
    arr = (OrcParams * n_s)(*slices)
    keep = []

    def ptr_array(items):
        a = (C.c_void_p * n_s)(*[it.ctypes.data for it in items])
        keep.append(items)
        return a

    fp = ptr_array([_f32(f[0]) for f in fixed_sets])
    fo = ptr_array([_i32(f[1]) for f in fixed_sets])
    mp = ptr_array([_f32(m[0]) for m in moving_sets])
    mo = ptr_array([_i32(m[1]) for m in moving_sets])
    init = _f32(init_xyt)
    init = init.reshape(len(init), -1)
    fixed_id, moving_id = _i32(fixed_id), _i32(moving_id)
    pz = None if prior_z is None else _f32(prior_z).reshape(len(init), -1)
    assert pz is None or pz.shape[1] == init.shape[1], "init and prior_z share one pose format"
    n = len(init)
    out = np.zeros(n, RESULT_DTYPE)
    its = np.zeros((n, slices[0].max_iterations), ITER_DTYPE) if want_iters else None
    lib().orc_align_multi_batch(C.cast(arr, C.c_void_p), n_s, C.cast(fp, C.c_void_p), C.cast(fo, C.c_void_p),
                                C.cast(mp, C.c_void_p), C.cast(mo, C.c_void_p), _ptr(fixed_id), _ptr(moving_id),
                                C.byref(prior) if prior is not None and pz is not None else None, _ptr(pz),
                                _ptr(init), init.shape[1], n, sum_mode, tree_threads, n_threads, _ptr(out), _ptr(its))
    return out, its


def best_of(results: np.ndarray, min_inliers: int, max_chi_per_inlier: float, min_inlier_ratio: float) -> int:
    results = np.ascontiguousarray(results)
    return lib().orc_best_of(_ptr(results), len(results), min_inliers, max_chi_per_inlier, min_inlier_ratio)


def max_threads() -> int:
    return lib().orc_max_threads()


def libm_atan2f(y: np.ndarray, x: np.ndarray) -> np.ndarray:
    y, x = _f32(y), _f32(x)
    out = np.zeros(len(x), np.float32)
    lib().orc_libm_atan2f_n(_ptr(y), _ptr(x), _ptr(out), len(x))
    return out


def libm_logf(x: np.ndarray) -> np.ndarray:
    x = _f32(x)
    out = np.zeros(len(x), np.float32)
    lib().orc_libm_logf_n(_ptr(x), _ptr(out), len(x))
    return out


def libm_sincosf(x: np.ndarray):
    x = _f32(x)
    s, c = np.zeros(len(x), np.float32), np.zeros(len(x), np.float32)
    lib().orc_libm_sincosf_n(_ptr(x), _ptr(s), _ptr(c), len(x))
    return s, c


def column(prm: OrcParams, y: np.ndarray, x: np.ndarray) -> np.ndarray:
    y, x = _f32(y), _f32(x)
    col = np.zeros(len(x), np.int32)
    lib().orc_column_n(C.byref(prm), _ptr(y), _ptr(x), _ptr(col), len(x))
    return col


def clip_scene(prm: OrcParams, scene: np.ndarray, robot_in_local_map_xyt, sensor_in_robot_xyt,
               voxelize_resolution: float = 0.0) -> np.ndarray:
    """SceneClipperProjective2D::compute: returns the clipped cloud [k, 4] in the robot frame."""
    scene = _f32(scene)
    out = np.zeros((prm.canvas_cols, 4), np.float32)
    k = lib().orc_clip_scene_voxelized(C.byref(prm), _ptr(scene), len(scene), as_iso(robot_in_local_map_xyt),
                                       as_iso(sensor_in_robot_xyt), voxelize_resolution, _ptr(out))
    return out[:k].copy()


def merge(prm: OrcParams, merge_threshold: float, scene: np.ndarray, measurement: np.ndarray, measurement_in_scene_xyt):
    """MergerProjective2D::compute: returns (new scene [n, 4], counters [new, merged, replaced])."""
    scene, measurement = _f32(scene), _f32(measurement)
    buf = np.zeros((len(scene) + prm.canvas_cols, 4), np.float32)
    buf[:len(scene)] = scene
    counters = np.zeros(3, np.int32)
    n = lib().orc_merge(C.byref(prm), merge_threshold, _ptr(buf), len(scene), _ptr(measurement), len(measurement),
                        as_iso(measurement_in_scene_xyt), _ptr(counters))
    return buf[:n].copy(), counters


def default_scan_params(**kw) -> OrcScanParams:
    p = OrcScanParams()
    lib().orc_default_scan_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def preprocess_scan(sp: OrcScanParams, ranges: np.ndarray) -> np.ndarray:
    """RawDataPreprocessorProjective2D::compute for one LaserMessage: returns the cloud [k, 4]."""
    ranges = _f32(ranges)
    out = np.zeros((max(len(ranges), 1), 4), np.float32)
    k = lib().orc_preprocess_scan(C.byref(sp), _ptr(ranges), len(ranges), _ptr(out))
    return out[:k].copy()


def preprocess_scans(sp: OrcScanParams, ranges: np.ndarray, n_threads=1):
    """batch of scans [n_scans, n_beams]: returns (points [n_scans, n_beams, 4], counts [n_scans])."""
    ranges = _f32(ranges)
    n_scans, n_beams = ranges.shape
    out = np.zeros((n_scans, n_beams, 4), np.float32)
    counts = np.zeros(n_scans, np.int32)
    lib().orc_preprocess_scans(C.byref(sp), _ptr(ranges), n_beams, n_scans, n_threads, _ptr(out), _ptr(counts))
    return out, counts
