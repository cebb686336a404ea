import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from srrg2_laser_slam_2d_b200 import Handle, default_params
from srrg2_laser_slam_2d_b200._abi import default_scan_params
from srrg2_laser_slam_2d_b200.synthetic import make_raw_scans
n = 4096
raw = make_raw_scans(n, seed=0xC0FFEE, device="cuda:0")
sp = default_scan_params(angle_min=raw.angle_min, angle_max=raw.angle_max, voxelize_resolution=0.02)
h = Handle(0, default_params(canvas_cols=1081))
r = torch.from_numpy(np.ascontiguousarray(raw.fixed_ranges)).cuda()
torch.cuda.synchronize()
for _ in range(5): h.preprocess_scans_to_set_dev(0, sp, r.data_ptr(), raw.fixed_ranges.shape[1], n)
h.sync()
t0 = time.perf_counter()
for _ in range(50): h.preprocess_scans_to_set_dev(0, sp, r.data_ptr(), raw.fixed_ranges.shape[1], n)
h.sync()
print("preprocess + pack: %.1f us per %d scans" % (1e6 * (time.perf_counter() - t0) / 50, n))
