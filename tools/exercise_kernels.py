"""Runs every kernel of libls2d.so that bench.py's default line does not reach, at batch sizes that fill the GPU, so
that one ncu run can capture them (tools/profile_round2.sh).  GPU box only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srrg2_laser_slam_2d_b200 import Handle, default_params  # noqa: E402
from srrg2_laser_slam_2d_b200 import _abi  # noqa: E402
from srrg2_laser_slam_2d_b200._abi import default_scan_params  # noqa: E402
from srrg2_laser_slam_2d_b200.synthetic import make_raw_scans, make_scan_pairs  # noqa: E402

N = 2048
sp = make_scan_pairs(N, n_beams=1081, seed=77)
init = np.ascontiguousarray(sp.init_xyt, np.float32)

# icp_general_kernel: Levenberg-Marquardt + inlier-only runs
h = Handle(0, default_params(canvas_cols=1081, normal_cos=0.9, max_iterations=10, algorithm=_abi.ALGORITHM_LM,
                             enable_inlier_only_runs=1))
h.upload_clouds(_abi.LS2D_FIXED, sp.fixed_pts, sp.fixed_off)
h.upload_clouds(_abi.LS2D_MOVING, sp.moving_pts, sp.moving_off)
for _ in range(3):
    h.align_batch(init)
# project / correspond service kernels
h.set_params(default_params(canvas_cols=1081, normal_cos=0.9, max_iterations=10))
for k in range(3):
    h.find_correspondences(k, k, init[k])
    h.project(_abi.LS2D_MOVING, k, init[k])
h.close()

# icp_stream_kernel: local maps of 8192 points on a 2048-column canvas
big = make_scan_pairs(256, n_beams=8192, seed=78)
h = Handle(0, default_params(canvas_cols=2048, normal_cos=0.8, max_iterations=10))
h.upload_clouds(_abi.LS2D_FIXED, big.fixed_pts, big.fixed_off)
h.upload_clouds(_abi.LS2D_MOVING, big.moving_pts, big.moving_off)
for _ in range(3):
    h.align_batch(np.ascontiguousarray(big.init_xyt, np.float32))
h.close()

# clipper (plain and voxelized), merger, pre-processor without voxelisation
raw = make_raw_scans(N, seed=79)
kw = dict(angle_min=raw.angle_min, angle_max=raw.angle_max)
h = Handle(0, default_params(canvas_cols=1081))
h.preprocess_scans_to_set(2, default_scan_params(voxelize_resolution=0.0, **kw), raw.moving_ranges)
ids = np.arange(N, dtype=np.int32)
robots = np.zeros((N, 3), np.float32)
for _ in range(3):
    h.clip_scenes_to_set(2, ids, robots, np.zeros(3, np.float32), _abi.LS2D_MOVING)
    h.clip_scenes(2, ids, robots, voxelize_resolution=0.05)
scene = sp.fixed_pts[sp.fixed_off[0]:sp.fixed_off[1]]
meas = sp.moving_pts[sp.moving_off[0]:sp.moving_off[1]]
for _ in range(3):
    h.merge_scene(scene, meas, sp.gt_xyt[0], 0.2)
h.close()
print("exercised")
