"""wall-clock time of ls2d_align_batch on the options only icp_general_kernel serves (GPU box)
usage: python tools/general_time.py [n_pairs]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srrg2_laser_slam_2d_b200 import Handle, default_params, _abi  # noqa: E402
from srrg2_laser_slam_2d_b200.synthetic import make_scan_pairs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sp = make_scan_pairs(n, n_beams=1081, seed=77)
init = np.ascontiguousarray(sp.init_xyt, np.float32)
base = dict(canvas_cols=1081, normal_cos=0.9, max_iterations=10)
cases = {"GN plane2plane (icp_fused2_kernel)": {},
         "GN point2point": dict(factor=_abi.FACTOR_POINT2POINT),
         "GN + termination 1e-4": dict(termination_epsilon=1e-4),
         "LM": dict(algorithm=_abi.ALGORITHM_LM),
         "LM + inlier-only runs": dict(algorithm=_abi.ALGORITHM_LM, enable_inlier_only_runs=1)}
for name, kw in cases.items():
    h = Handle(0, default_params(**base, **kw))
    h.upload_clouds(_abi.LS2D_FIXED, sp.fixed_pts, sp.fixed_off)
    h.upload_clouds(_abi.LS2D_MOVING, sp.moving_pts, sp.moving_off)
    for _ in range(3):
        out = h.align_batch(init)
    t0 = time.perf_counter()
    for _ in range(10):
        out = h.align_batch(init)
    dt = (time.perf_counter() - t0) / 10
    print("%-36s %8.3f ms per %d pairs  %7.2f M pairs/s  mean iterations %.1f  success %.3f" %
          (name, 1e3 * dt, n, n / dt / 1e6, out["iterations"].mean(), (out["status"] == 0).mean()))
    h.close()
