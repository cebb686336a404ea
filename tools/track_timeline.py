"""GPU timeline of one ls2d_track_batch step (CUPTI through torch.profiler): start / duration of every kernel and copy.
usage (GPU box): python tools/track_timeline.py [n_frames]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srrg2_laser_slam_2d_b200 import Handle, default_params  # noqa: E402
from srrg2_laser_slam_2d_b200._abi import RESULT_DTYPE, default_scan_params  # noqa: E402
from srrg2_laser_slam_2d_b200.synthetic import make_raw_scans  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
raw = make_raw_scans(n, seed=0xC0FFEE, device="cuda:0")
sp_map = default_scan_params(angle_min=raw.angle_min, angle_max=raw.angle_max, voxelize_resolution=0.0)
sp = default_scan_params(angle_min=raw.angle_min, angle_max=raw.angle_max, voxelize_resolution=0.02)
h = Handle(0, default_params(**bench.TRACK))
h.preprocess_scans_to_set(2, sp_map, raw.moving_ranges)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
ranges, ids = pin(raw.fixed_ranges), pin(np.arange(n, dtype=np.int32))
robots, init = pin(np.zeros((n, 3), np.float32)), pin(np.zeros((n, 3), np.float32))
out = torch.zeros(n * RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(RESULT_DTYPE)
for _ in range(5):
    h.track_batch(sp, ranges, 2, ids, robots, init, out)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        h.track_batch(sp, ranges, 2, ids, robots, init, out)
    torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/track_trace.json")
ev = [e for e in json.load(open("/tmp/track_trace.json"))["traceEvents"]
      if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
for e in ev:
    print("%9.1f us  +%8.1f us  stream %-3s %s" % (e["ts"] - t0, e["dur"], e["args"].get("stream", "?"), e["name"][:70]))
