"""times ls2d_align_batch_dev on the config-3 batch (device-resident), prints ms per launch; LS2D_LIB selects the build"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from srrg2_laser_slam_2d_b200 import Handle, default_params
from srrg2_laser_slam_2d_b200._abi import LS2D_FIXED, LS2D_MOVING, RESULT_DTYPE
from srrg2_laser_slam_2d_b200.synthetic import make_scan_pairs
n = 4096
beams = int(sys.argv[1]) if len(sys.argv) > 1 else 1081
sp = make_scan_pairs(n, n_beams=beams, seed=0xC0FFEE, device="cuda:0")
dev = torch.device("cuda:0")
h = Handle(0, default_params(canvas_cols=beams, point_distance=0.5, normal_cos=0.9, cauchy_chi_threshold=0.01, max_iterations=10))
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st); h.set_stream(st.cuda_stream)
fp, fo = torch.from_numpy(sp.fixed_pts).to(dev), torch.from_numpy(sp.fixed_off).to(dev)
mp, mo = torch.from_numpy(sp.moving_pts).to(dev), torch.from_numpy(sp.moving_off).to(dev)
init = torch.from_numpy(sp.init_xyt).to(dev)
out = torch.zeros(n * 20, dtype=torch.int32, device=dev)
h.set_clouds_dev(LS2D_FIXED, fp.data_ptr(), fo.data_ptr(), n, beams); h.set_clouds_dev(LS2D_MOVING, mp.data_ptr(), mo.data_ptr(), n, beams)
for _ in range(5): h.align_batch_dev(None, None, init.data_ptr(), n, out.data_ptr())
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(30): h.align_batch_dev(None, None, init.data_ptr(), n, out.data_ptr())
    e1.record(st); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 30)
r = np.frombuffer(out.cpu().numpy().tobytes(), dtype=RESULT_DTYPE)
print(os.environ.get("LS2D_LIB", "default"), beams, "align ms/launch %.4f" % best, "checksum", int(r["n_corr"].sum()), float(r["chi_inliers"].sum()), float(r["theta"].sum()))
