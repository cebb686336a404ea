"""ctypes binding of the C ABI in include/ls2d.h (libls2d.so).

Thin plumbing for the Python test-suite and bench harness: every call goes straight to the CUDA
library.  There is no Python or CPU implementation of the path here -- if the library is missing or
no CUDA device is present the calls fail loudly (Ls2dError)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LS2D_LIB") or os.path.join(_HERE, "libls2d.so")  # LS2D_LIB: A/B builds (bench plumbing)
MATHCHECK_PATH = os.path.join(_HERE, "libls2d_mathcheck.so")

LS2D_FIXED, LS2D_MOVING = 0, 1
STATUS_SUCCESS, STATUS_NOT_ENOUGH_CORRESPONDENCES, STATUS_NOT_ENOUGH_INLIERS, STATUS_SINGULAR = 0, 1, 2, 3


class Ls2dError(RuntimeError):
    pass


class Params(C.Structure):
    """ls2d_params -- the reference's parameter names (include/ls2d.h)."""
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


FACTOR_PLANE2PLANE, FACTOR_POINT2POINT = 0, 1
ALGORITHM_GN, ALGORITHM_LM = 0, 1
POSE_XYT, POSE_ISO = 0, 1


class Prior(C.Structure):
    """ls2d_prior -- information matrix (upper triangle) and robustifier threshold of the odometry prior slice."""
    _fields_ = [("information", C.c_float * 6), ("cauchy_chi_threshold", C.c_float)]


def make_prior(information, cauchy_chi_threshold: float = -1.0) -> Prior:
    info = np.asarray(information, np.float32)
    if info.shape == (3, 3):
        info = info[np.triu_indices(3)]
    pr = Prior()
    pr.information = (C.c_float * 6)(*[float(v) for v in info])
    pr.cauchy_chi_threshold = cauchy_chi_threshold
    return pr


class ScanParams(C.Structure):
    """ls2d_scan_params -- LaserMessage fields + RawDataPreprocessorProjective2D / normal computator PARAMs."""
    _fields_ = [("angle_min", C.c_float), ("angle_max", C.c_float), ("msg_range_min", C.c_float),
                ("msg_range_max", C.c_float), ("range_min", C.c_float), ("range_max", C.c_float),
                ("voxelize_resolution", C.c_float), ("normal_point_distance", C.c_float),
                ("normal_min_points", C.c_int32)]


class Gates(C.Structure):
    _fields_ = [("min_inliers", C.c_int32), ("max_chi_per_inlier", C.c_float), ("min_inlier_ratio", C.c_float)]


RESULT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("theta", "<f4"), ("chi_inliers", "<f4"),
                         ("chi_kernelized", "<f4"), ("n_inliers", "<i4"), ("n_kernelized", "<i4"),
                         ("n_corr", "<i4"), ("status", "<i4"), ("iterations", "<i4"), ("H", "<f4", (6,)),
                         ("c", "<f4"), ("s", "<f4"), ("lm_rejected", "<i4"), ("reserved", "<i4")])
ITER_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("theta", "<f4"), ("chi_inliers", "<f4"),
                       ("chi_kernelized", "<f4"), ("n_inliers", "<i4"), ("n_kernelized", "<i4"),
                       ("n_corr", "<i4"), ("c", "<f4"), ("s", "<f4")])
BEST_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("theta", "<f4"), ("chi_inliers", "<f4"),
                       ("n_inliers", "<i4"), ("n_corr", "<i4"), ("candidate", "<i4"), ("guess", "<i4"),
                       ("c", "<f4"), ("s", "<f4"), ("iterations", "<i4"), ("reserved", "<i4")])
assert RESULT_DTYPE.itemsize == 80 and ITER_DTYPE.itemsize == 40 and BEST_DTYPE.itemsize == 48

# every symbol include/ls2d.h declares (tests/test_abi_symbols.py checks the library exports them all)
EXPORTS = [
    "ls2d_create", "ls2d_destroy", "ls2d_set_stream", "ls2d_sync", "ls2d_strerror", "ls2d_version",
    "ls2d_default_params", "ls2d_set_params", "ls2d_get_params", "ls2d_set_pose_format", "ls2d_get_pose_format",
    "ls2d_upload_clouds", "ls2d_set_clouds_dev",
    "ls2d_align_batch", "ls2d_align_batch_dev", "ls2d_align_pairs_host", "ls2d_score_batch",
    "ls2d_score_batch_dev", "ls2d_find_correspondences", "ls2d_project", "ls2d_verify", "ls2d_verify_dev",
    "ls2d_reduce_best", "ls2d_verify_sharded_nccl", "ls2d_reduction_shape", "ls2d_launch_count",
    "ls2d_clip_scenes", "ls2d_merge_scene", "ls2d_merge_scene_dev", "ls2d_align_multi", "ls2d_align_multi_dev",
    "ls2d_find_correspondences_in", "ls2d_default_scan_params", "ls2d_preprocess_scans",
    "ls2d_preprocess_scans_to_set", "ls2d_preprocess_scans_to_set_dev", "ls2d_download_clouds",
    "ls2d_clip_scenes_to_set", "ls2d_track_batch", "ls2d_verify_pairs", "ls2d_verify_pairs_dev",
    "ls2d_clip_scenes_voxelized", "ls2d_multi_reduction_threads", "ls2d_classify_correspondences", "ls2d_score_reduction_shape", "ls2d_selftest_gated_sqrt", "ls2d_multi_reduction_shape",
]

_lib = None


def load():
    """dlopen libls2d.so and declare the prototypes.  Raises Ls2dError if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Ls2dError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback for this path)")
    L = C.CDLL(LIB_PATH)
    vp, i32, f32, i64 = C.c_void_p, C.c_int32, C.c_float, C.c_int64
    PP, GP = C.POINTER(Params), C.POINTER(Gates)
    L.ls2d_create.argtypes = [C.POINTER(vp), C.c_int]
    L.ls2d_destroy.argtypes = [vp]
    L.ls2d_set_stream.argtypes = [vp, vp]
    L.ls2d_sync.argtypes = [vp]
    L.ls2d_strerror.argtypes, L.ls2d_strerror.restype = [C.c_int], C.c_char_p
    L.ls2d_default_params.argtypes, L.ls2d_default_params.restype = [PP], None
    L.ls2d_set_params.argtypes = [vp, PP]
    L.ls2d_get_params.argtypes = [vp, PP]
    L.ls2d_set_pose_format.argtypes = [vp, C.c_int]
    L.ls2d_get_pose_format.argtypes = [vp]
    L.ls2d_upload_clouds.argtypes = [vp, C.c_int, vp, vp, i32]
    L.ls2d_set_clouds_dev.argtypes = [vp, C.c_int, vp, vp, i32, i32]
    L.ls2d_align_batch.argtypes = [vp, vp, vp, vp, i32, vp, vp]
    L.ls2d_align_batch_dev.argtypes = [vp, vp, vp, vp, i32, vp, vp]
    L.ls2d_align_pairs_host.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp]
    L.ls2d_score_batch.argtypes = [vp, vp, vp, vp, i32, vp]
    L.ls2d_score_batch_dev.argtypes = [vp, vp, vp, vp, i32, vp]
    L.ls2d_find_correspondences.argtypes = [vp, i32, i32, vp, vp, vp, C.POINTER(i32)]
    L.ls2d_find_correspondences_in.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, C.POINTER(i32)]
    L.ls2d_project.argtypes = [vp, C.c_int, i32, vp, vp, vp]
    L.ls2d_classify_correspondences.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, i32, vp]
    L.ls2d_verify.argtypes = [vp, i32, vp, i32, vp, i32, GP, i32, vp, vp]
    L.ls2d_verify_dev.argtypes = [vp, i32, vp, i32, vp, i32, GP, i32, vp, vp]
    L.ls2d_reduce_best.argtypes = [vp, i32, vp]
    L.ls2d_verify_pairs.argtypes = [vp, vp, vp, vp, i32, vp, i32, GP, vp, vp]
    L.ls2d_verify_pairs_dev.argtypes = [vp, vp, vp, vp, i32, vp, i32, GP, vp, vp]
    L.ls2d_verify_sharded_nccl.argtypes = [vp, i32, vp, i32, vp, i32, GP, i32, vp, i32, vp]
    L.ls2d_clip_scenes.argtypes = [vp, C.c_int, vp, vp, vp, i32, vp, vp]
    L.ls2d_merge_scene.argtypes = [vp, vp, C.POINTER(i32), i32, vp, i32, vp, f32, vp]
    L.ls2d_merge_scene_dev.argtypes = [vp, vp, vp, i32, vp, i32, vp, f32, vp]
    L.ls2d_align_multi.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, vp, vp]
    L.ls2d_align_multi_dev.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, vp, vp]
    SP = C.POINTER(ScanParams)
    L.ls2d_default_scan_params.argtypes, L.ls2d_default_scan_params.restype = [SP], None
    L.ls2d_preprocess_scans.argtypes = [vp, SP, vp, i32, i32, vp, vp]
    L.ls2d_preprocess_scans_to_set.argtypes = [vp, C.c_int, SP, vp, i32, i32]
    L.ls2d_preprocess_scans_to_set_dev.argtypes = [vp, C.c_int, SP, vp, i32, i32]
    L.ls2d_download_clouds.argtypes = [vp, C.c_int, vp, vp, i32, i64]
    L.ls2d_clip_scenes_voxelized.argtypes = [vp, C.c_int, vp, vp, vp, i32, f32, vp, vp]
    L.ls2d_clip_scenes_to_set.argtypes = [vp, C.c_int, vp, vp, vp, i32, C.c_int]
    L.ls2d_track_batch.argtypes = [vp, SP, vp, i32, i32, C.c_int, vp, vp, vp, vp]
    L.ls2d_reduction_shape.argtypes = [PP, i32]
    L.ls2d_score_reduction_shape.argtypes = [PP, i32]
    L.ls2d_multi_reduction_shape.argtypes = [vp, i32, i32, i32, i32]
    L.ls2d_selftest_gated_sqrt.argtypes = [vp, f32, f32, C.POINTER(i64), C.POINTER(i64)]
    L.ls2d_launch_count.argtypes, L.ls2d_launch_count.restype = [vp], i64
    _lib = L
    return L


def default_params(**kw) -> Params:
    p = Params()
    load().ls2d_default_params(C.byref(p))
    for k, v in kw.items():
        if k == "sensor_in_robot":
            set_sensor(p, v)
        elif k == "sensor_in_robot_cs":
            p.sensor_in_robot_cs = (C.c_float * 2)(*[float(x) for x in v])
        else:
            setattr(p, k, v)
    return p


def set_sensor(p: Params, v):
    """sensor_in_robot of a params record: 3 values = (x, y, theta), 4 values = the isometry (tx, ty, c, s), which
    switches a WithSensor slice to with_sensor = 2"""
    v = [float(x) for x in v]
    if len(v) == 4:
        p.sensor_in_robot = (C.c_float * 3)(v[0], v[1], 0.0)
        p.sensor_in_robot_cs = (C.c_float * 2)(v[2], v[3])
        if p.with_sensor:
            p.with_sensor = 2
    else:
        p.sensor_in_robot = (C.c_float * 3)(*v)


def default_scan_params(**kw) -> ScanParams:
    p = ScanParams()
    load().ls2d_default_scan_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def reduction_threads(max_points: int, canvas_cols: int = 1081, params: Params | None = None) -> int:
    """shape of the H/b reduction the aligner runs for these parameters and cloud size (threads per pair | warp-combine
    flag << 16 | fused-accumulation flag << 17): the value the oracle's SUM_TREE mode takes as tree_threads"""
    p = params if params is not None else default_params(canvas_cols=canvas_cols)
    shape = load().ls2d_reduction_shape(C.byref(p), max_points)
    if shape < 0:
        raise Ls2dError(f"ls2d_reduction_shape: {shape}")
    return shape


def score_reduction_threads(max_points: int, canvas_cols: int = 1081, params: Params | None = None) -> int:
    """the same for the scoring pass (ls2d_score_batch)"""
    p = params if params is not None else default_params(canvas_cols=canvas_cols)
    shape = load().ls2d_score_reduction_shape(C.byref(p), max_points)
    if shape < 0:
        raise Ls2dError(f"ls2d_score_reduction_shape: {shape}")
    return shape


def multi_reduction_threads(slices=None, max_fixed_points: int = 0, max_moving_points: int = 0,
                            shared_moving: bool = True) -> int:
    """shape of the multi-slice aligner's per-slice reduction: the general kernel's (no arguments), or the one
    ls2d_align_multi runs for these slices (list of Params), cloud sizes and moving-set sharing"""
    if slices is None:
        return load().ls2d_multi_reduction_threads()
    arr = (Params * len(slices))(*slices)
    shape = load().ls2d_multi_reduction_shape(C.cast(arr, C.c_void_p), len(slices), max_fixed_points,
                                              max_moving_points, 1 if shared_moving else 0)
    if shape < 0:
        raise Ls2dError(f"ls2d_multi_reduction_shape: {shape}")
    return shape


def reduce_best(records: np.ndarray) -> np.ndarray:
    records = np.ascontiguousarray(records, dtype=BEST_DTYPE)
    out = np.zeros(1, BEST_DTYPE)
    rc = load().ls2d_reduce_best(_ptr(records), len(records), _ptr(out))
    if rc:
        raise Ls2dError(load().ls2d_strerror(rc).decode())
    return out[0]


class Handle:
    """One ls2d_handle: one device, one stream.  Host-array methods copy in/out; *_dev methods take raw
    device addresses (ints, e.g. torch.Tensor.data_ptr()) and are asynchronous on the handle's stream."""

    def __init__(self, device: int = 0, params: Params | None = None):
        self._L = load()
        self._h = C.c_void_p()
        self._check(self._L.ls2d_create(C.byref(self._h), device))
        self.params = params if params is not None else default_params()
        self.set_params(self.params)

    def _check(self, rc: int):
        if rc != 0:
            raise Ls2dError(f"ls2d error {rc}: {self._L.ls2d_strerror(rc).decode()}")

    def close(self):
        if self._h:
            self._L.ls2d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- configuration
    def set_params(self, p: Params):
        self._check(self._L.ls2d_set_params(self._h, C.byref(p)))
        self.params = p

    def set_stream(self, cuda_stream: int | None):
        self._check(self._L.ls2d_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def sync(self):
        self._check(self._L.ls2d_sync(self._h))

    def set_pose_format(self, fmt: int):
        self._check(self._L.ls2d_set_pose_format(self._h, fmt))

    @property
    def pose_format(self) -> int:
        return int(self._L.ls2d_get_pose_format(self._h))

    def _poses(self, a, lead=None):
        """host poses as a contiguous float32 array [..., 3] (x, y, theta) or [..., 4] (tx, ty, c, s); the handle's
        pose format follows the array (every pose argument of ONE call must use the same format)"""
        a = np.ascontiguousarray(a, dtype=np.float32)
        k = a.shape[-1]
        assert k in (3, 4), "a pose is (x, y, theta) or (tx, ty, c, s)"
        want = POSE_ISO if k == 4 else POSE_XYT
        if self.pose_format != want:
            self.set_pose_format(want)
        return a if lead is None else a.reshape(lead + (k,))

    @property
    def iters_per_pair(self) -> int:
        return self.params.max_iterations * (2 if self.params.enable_inlier_only_runs else 1)

    @property
    def launch_count(self) -> int:
        return int(self._L.ls2d_launch_count(self._h))

    # ---- clouds
    def upload_clouds(self, which: int, points: np.ndarray, offsets: np.ndarray):
        points, offsets = _f32(points), _i32(offsets)
        assert points.ndim == 2 and points.shape[1] == 4 and offsets[-1] == len(points)
        self._check(self._L.ls2d_upload_clouds(self._h, which, _ptr(points), _ptr(offsets), len(offsets) - 1))
        self.sync()  # the numpy temporaries may die after this call

    def set_clouds_dev(self, which: int, points_ptr: int, offsets_ptr: int, n_clouds: int, max_points: int):
        self._check(self._L.ls2d_set_clouds_dev(self._h, which, C.c_void_p(points_ptr), C.c_void_p(offsets_ptr),
                                                n_clouds, max_points))

    # ---- registration
    def align_batch(self, init_xyt, fixed_id=None, moving_id=None, want_iters: bool = False):
        init = self._poses(init_xyt)
        init = init.reshape(-1, init.shape[-1])
        fid, mid = _i32(fixed_id), _i32(moving_id)
        n = len(init)
        out = np.zeros(n, RESULT_DTYPE)
        its = np.zeros((n, self.iters_per_pair), ITER_DTYPE) if want_iters else None
        self._check(self._L.ls2d_align_batch(self._h, _ptr(fid), _ptr(mid), _ptr(init), n, _ptr(out), _ptr(its)))
        return (out, its) if want_iters else out

    def align_batch_dev(self, fixed_id_ptr, moving_id_ptr, init_ptr: int, n_pairs: int, out_ptr: int,
                        iters_ptr: int | None = None):
        self._check(self._L.ls2d_align_batch_dev(self._h, C.c_void_p(fixed_id_ptr or 0), C.c_void_p(moving_id_ptr or 0),
                                                 C.c_void_p(init_ptr), n_pairs, C.c_void_p(out_ptr),
                                                 C.c_void_p(iters_ptr or 0)))

    def align_pairs_host(self, fixed_pts, fixed_off, moving_pts, moving_off, init_xyt, out=None):
        """One call from host buffers (upload + align + download) -- the end-to-end path bench.py times."""
        n = len(fixed_off) - 1
        if out is None:
            out = np.zeros(n, RESULT_DTYPE)
        init_xyt = self._poses(init_xyt)  # no copy when the caller hands a contiguous float32 (e.g. pinned) array
        self._check(self._L.ls2d_align_pairs_host(self._h, _ptr(fixed_pts), _ptr(fixed_off), _ptr(moving_pts),
                                                  _ptr(moving_off), _ptr(init_xyt), n, _ptr(out)))
        return out

    def align_multi(self, slices, fixed_sets, moving_sets, init_xyt, prior: Prior | None = None, prior_z=None,
                    fixed_id=None, moving_id=None, want_iters: bool = False):
        """MultiAligner2D with several laser slices (+ odometry prior): slices = list of Params, fixed_sets /
        moving_sets = cloud-set ids per slice (uploaded with upload_clouds), prior_z = [n_pairs, 3]."""
        n_s = len(slices)
        arr = (Params * n_s)(*slices)
        fs, ms = _i32(fixed_sets), _i32(moving_sets)
        init = self._poses(init_xyt)
        init = init.reshape(-1, init.shape[-1])
        pz = None if prior_z is None else self._poses(prior_z).reshape(len(init), -1)
        assert pz is None or pz.shape[1] == init.shape[1], "init and prior_z share one pose format"
        fid, mid = _i32(fixed_id), _i32(moving_id)
        n = len(init)
        out = np.zeros(n, RESULT_DTYPE)
        its = np.zeros((n, slices[0].max_iterations), ITER_DTYPE) if want_iters else None
        self._check(self._L.ls2d_align_multi(self._h, C.cast(arr, C.c_void_p), _ptr(fs), _ptr(ms), n_s,
                                             C.cast(C.pointer(prior), C.c_void_p) if prior is not None else None,
                                             _ptr(pz), _ptr(fid), _ptr(mid), _ptr(init), n, _ptr(out), _ptr(its)))
        return (out, its) if want_iters else out

    def align_multi_dev(self, slices, fixed_sets, moving_sets, init_ptr: int, n_pairs: int, out_ptr: int,
                        prior: Prior | None = None, prior_z_ptr: int | None = None, iters_ptr: int | None = None):
        n_s = len(slices)
        arr = (Params * n_s)(*slices)
        fs, ms = _i32(fixed_sets), _i32(moving_sets)
        self._check(self._L.ls2d_align_multi_dev(self._h, C.cast(arr, C.c_void_p), _ptr(fs), _ptr(ms), n_s,
                                                 C.cast(C.pointer(prior), C.c_void_p) if prior is not None else None,
                                                 C.c_void_p(prior_z_ptr or 0), None, None, C.c_void_p(init_ptr),
                                                 n_pairs, C.c_void_p(out_ptr), C.c_void_p(iters_ptr or 0)))

    def score_batch(self, xyt, fixed_id=None, moving_id=None):
        xyt = self._poses(xyt)
        xyt = xyt.reshape(-1, xyt.shape[-1])
        fid, mid = _i32(fixed_id), _i32(moving_id)
        out = np.zeros(len(xyt), RESULT_DTYPE)
        self._check(self._L.ls2d_score_batch(self._h, _ptr(fid), _ptr(mid), _ptr(xyt), len(xyt), _ptr(out)))
        return out

    def score_batch_dev(self, fixed_id_ptr, moving_id_ptr, xyt_ptr: int, n_pairs: int, out_ptr: int):
        self._check(self._L.ls2d_score_batch_dev(self._h, C.c_void_p(fixed_id_ptr or 0), C.c_void_p(moving_id_ptr or 0),
                                                 C.c_void_p(xyt_ptr), n_pairs, C.c_void_p(out_ptr)))

    # ---- finder / projector
    def find_correspondences(self, fixed_id: int, moving_id: int, local_map_in_sensor_xyt):
        xyt = self._poses(local_map_in_sensor_xyt)
        cols = self.params.canvas_cols
        fi, mi = np.zeros(cols, np.int32), np.zeros(cols, np.int32)
        n = C.c_int32(0)
        self._check(self._L.ls2d_find_correspondences(self._h, fixed_id, moving_id, _ptr(xyt), _ptr(fi), _ptr(mi),
                                                      C.byref(n)))
        return fi[:n.value].copy(), mi[:n.value].copy()

    def selftest_gated_sqrt(self, lo: float, hi: float):
        """(values checked, mismatches) of the kernels' gated square root against __fsqrt_rn over [lo, hi]"""
        n, bad = C.c_int64(0), C.c_int64(0)
        self._check(self._L.ls2d_selftest_gated_sqrt(self._h, lo, hi, C.byref(n), C.byref(bad)))
        return n.value, bad.value

    def classify_correspondences(self, fixed_id: int, moving_id: int, moving_in_fixed, fixed_idx, moving_idx,
                                 fixed_set: int = LS2D_FIXED, moving_set: int = LS2D_MOVING):
        """inlier flags (chi < cauchy_chi_threshold at the estimate) of given correspondences"""
        X = self._poses(moving_in_fixed)
        fi, mi = _i32(fixed_idx), _i32(moving_idx)
        out = np.zeros(len(fi), np.uint8)
        self._check(self._L.ls2d_classify_correspondences(self._h, fixed_set, moving_set, fixed_id, moving_id, _ptr(X),
                                                          _ptr(fi), _ptr(mi), len(fi), _ptr(out)))
        return out.astype(bool)

    def project(self, which: int, cloud_id: int, camera_pose_xyt):
        xyt = self._poses(camera_pose_xyt)
        cols = self.params.canvas_cols
        idx, depth = np.zeros(cols, np.int32), np.zeros(cols, np.float32)
        self._check(self._L.ls2d_project(self._h, which, cloud_id, _ptr(xyt), _ptr(idx), _ptr(depth)))
        return idx, depth

    # ---- local-map maintenance
    def clip_scenes(self, which: int, cloud_ids, robot_in_local_map_xyt, sensor_in_robot_xyt=(0.0, 0.0, 0.0),
                    voxelize_resolution: float = 0.0):
        """SceneClipperProjective2D: returns a list of clipped clouds [k_r, 4] in the robot frame."""
        ids = _i32(cloud_ids)
        sen = self._poses(sensor_in_robot_xyt)
        rob = self._poses(robot_in_local_map_xyt).reshape(-1, len(sen))
        n, cols = len(ids), self.params.canvas_cols
        out = np.zeros((n, cols, 4), np.float32)
        cnt = np.zeros(n, np.int32)
        if voxelize_resolution > 0:
            self._check(self._L.ls2d_clip_scenes_voxelized(self._h, which, _ptr(ids), _ptr(rob), _ptr(sen), n,
                                                           voxelize_resolution, _ptr(out), _ptr(cnt)))
        else:
            self._check(self._L.ls2d_clip_scenes(self._h, which, _ptr(ids), _ptr(rob), _ptr(sen), n, _ptr(out), _ptr(cnt)))
        return [out[r, :cnt[r]].copy() for r in range(n)]

    def merge_scene(self, scene: np.ndarray, measurement: np.ndarray, measurement_in_scene_xyt, merge_threshold: float = 0.2):
        """MergerProjective2D: returns (new scene [n, 4], counters [new, merged, replaced])."""
        scene, measurement = _f32(scene), _f32(measurement)
        cap = len(scene) + self.params.canvas_cols
        buf = np.zeros((cap, 4), np.float32)
        buf[:len(scene)] = scene
        size = C.c_int32(len(scene))
        counters = np.zeros(3, np.int32)
        xyt = self._poses(measurement_in_scene_xyt)
        self._check(self._L.ls2d_merge_scene(self._h, _ptr(buf), C.byref(size), cap, _ptr(measurement), len(measurement),
                                             _ptr(xyt), merge_threshold, _ptr(counters)))
        return buf[:size.value].copy(), counters

    # ---- raw scans (RawDataPreprocessorProjective2D)
    def preprocess_scans(self, sp: ScanParams, ranges: np.ndarray):
        """ranges [n_scans, n_beams] -> (points [n_scans, n_beams, 4], counts [n_scans])"""
        ranges = _f32(ranges)
        n_scans, n_beams = ranges.shape
        out = np.zeros((n_scans, n_beams, 4), np.float32)
        cnt = np.zeros(n_scans, np.int32)
        self._check(self._L.ls2d_preprocess_scans(self._h, C.byref(sp), _ptr(ranges), n_beams, n_scans, _ptr(out),
                                                  _ptr(cnt)))
        return out, cnt

    def preprocess_scans_to_set(self, which: int, sp: ScanParams, ranges: np.ndarray):
        """raw ranges [n_scans, n_beams] become the resident cloud set `which`"""
        ranges = _f32(ranges)
        n_scans, n_beams = ranges.shape
        self._check(self._L.ls2d_preprocess_scans_to_set(self._h, which, C.byref(sp), _ptr(ranges), n_beams, n_scans))
        self.sync()

    def preprocess_scans_to_set_dev(self, which: int, sp: ScanParams, ranges_ptr: int, n_beams: int, n_scans: int):
        self._check(self._L.ls2d_preprocess_scans_to_set_dev(self._h, which, C.byref(sp), C.c_void_p(ranges_ptr),
                                                             n_beams, n_scans))

    def download_clouds(self, which: int, n_clouds: int, capacity_points: int):
        pts = np.zeros((capacity_points, 4), np.float32)
        off = np.zeros(n_clouds + 1, np.int32)
        self._check(self._L.ls2d_download_clouds(self._h, which, _ptr(pts), _ptr(off), n_clouds, capacity_points))
        return pts[:off[-1]].copy(), off

    def clip_scenes_to_set(self, scene_set: int, cloud_ids, robot_in_local_map_xyt, sensor_in_robot_xyt, out_set: int):
        ids = _i32(cloud_ids)
        sen = self._poses(sensor_in_robot_xyt)
        rob = self._poses(robot_in_local_map_xyt).reshape(-1, len(sen))
        self._check(self._L.ls2d_clip_scenes_to_set(self._h, scene_set, _ptr(ids), _ptr(rob), _ptr(sen), len(ids),
                                                    out_set))
        self.sync()

    def track_batch(self, sp: ScanParams, ranges, scene_set: int, scene_ids, robot_in_local_map_xyt, init_xyt=None,
                    out=None):
        """MultiTracker2D frame step (pre-process -> clip -> align) for a batch of frames, from host buffers."""
        ranges = _f32(ranges)
        n, n_beams = ranges.shape
        ids = _i32(scene_ids)
        rob = self._poses(robot_in_local_map_xyt)
        rob = rob.reshape(-1, rob.shape[-1])
        init = None if init_xyt is None else self._poses(init_xyt).reshape(len(rob), -1)
        assert init is None or init.shape[1] == rob.shape[1], "robot poses and initial guesses share one pose format"
        if out is None:
            out = np.zeros(n, RESULT_DTYPE)
        self._check(self._L.ls2d_track_batch(self._h, C.byref(sp), _ptr(ranges), n_beams, n, scene_set, _ptr(ids),
                                             _ptr(rob), _ptr(init), _ptr(out)))
        return out

    # ---- loop-closure verification
    def verify(self, query_id: int, candidate_ids, guesses_xyt, gates: Gates, candidate_base: int = 0,
               want_all: bool = False):
        g = self._poses(guesses_xyt)
        n_cand, n_guess = g.shape[0], g.shape[1]
        cand = _i32(candidate_ids)
        best = np.zeros(1, BEST_DTYPE)
        allr = np.zeros(n_cand * n_guess, RESULT_DTYPE) if want_all else None
        self._check(self._L.ls2d_verify(self._h, query_id, _ptr(cand), n_cand, _ptr(g), n_guess, C.byref(gates),
                                        candidate_base, _ptr(best), _ptr(allr)))
        return (best[0], allr) if want_all else best[0]

    def verify_sharded_nccl(self, query_id: int, cand_ptr, n_cand: int, guesses_ptr: int, n_guess: int, gates: Gates,
                            candidate_base: int, nccl_comm_ptr: int, n_ranks: int):
        """ls2d_verify_dev on this rank's shard + all-gather of the 48-byte records over the caller's ncclComm_t +
        the deterministic best-of: every rank returns the same global winner (device-resident inputs)"""
        best = np.zeros(1, BEST_DTYPE)
        self._check(self._L.ls2d_verify_sharded_nccl(self._h, query_id, C.c_void_p(cand_ptr or 0), n_cand,
                                                     C.c_void_p(guesses_ptr), n_guess, C.byref(gates), candidate_base,
                                                     C.c_void_p(nccl_comm_ptr), n_ranks, _ptr(best)))
        return best[0]

    def verify_pairs(self, fixed_ids, moving_ids, guesses_xyt, group_offsets, gates: Gates, want_all: bool = False):
        """all-pairs search: one ls2d_best per group of consecutive pairs (group_offsets: CSR over the pairs)"""
        fid, mid, off = _i32(fixed_ids), _i32(moving_ids), _i32(group_offsets)
        g = self._poses(guesses_xyt)
        g = g.reshape(-1, g.shape[-1])
        n, n_groups = len(g), len(off) - 1
        best = np.zeros(n_groups, BEST_DTYPE)
        allr = np.zeros(n, RESULT_DTYPE) if want_all else None
        self._check(self._L.ls2d_verify_pairs(self._h, _ptr(fid), _ptr(mid), _ptr(g), n, _ptr(off), n_groups,
                                              C.byref(gates), _ptr(best), _ptr(allr)))
        return (best, allr) if want_all else best

    def verify_pairs_dev(self, fid_ptr, mid_ptr, guesses_ptr: int, n_pairs: int, group_off_ptr: int, n_groups: int,
                         gates: Gates, best_ptr: int, all_ptr: int | None = None):
        self._check(self._L.ls2d_verify_pairs_dev(self._h, C.c_void_p(fid_ptr or 0), C.c_void_p(mid_ptr or 0),
                                                  C.c_void_p(guesses_ptr), n_pairs, C.c_void_p(group_off_ptr), n_groups,
                                                  C.byref(gates), C.c_void_p(best_ptr), C.c_void_p(all_ptr or 0)))

    def verify_dev(self, query_id: int, cand_ptr, n_cand: int, guesses_ptr: int, n_guess: int, gates: Gates,
                   candidate_base: int, best_ptr: int, all_ptr: int | None = None):
        self._check(self._L.ls2d_verify_dev(self._h, query_id, C.c_void_p(cand_ptr or 0), n_cand,
                                            C.c_void_p(guesses_ptr), n_guess, C.byref(gates), candidate_base,
                                            C.c_void_p(best_ptr), C.c_void_p(all_ptr or 0)))
