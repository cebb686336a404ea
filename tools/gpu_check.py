"""Developer diagnostic (run under gpurun): GPU path vs oracle on a seeded batch, printed as counts."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402
from srrg2_laser_slam_2d_b200 import Handle, default_params  # noqa: E402
from srrg2_laser_slam_2d_b200._abi import LS2D_FIXED, LS2D_MOVING, reduction_threads  # noqa: E402
from srrg2_laser_slam_2d_b200.synthetic import make_scan_pairs  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    sp = make_scan_pairs(n, seed=7)
    kw = dict(canvas_cols=1081, normal_cos=0.9, max_iterations=10)
    op = ob.default_params(**kw)
    gp = default_params(**kw)
    T = reduction_threads(1081)
    print("threads per pair:", T)
    h = Handle(0, gp)
    h.upload_clouds(LS2D_FIXED, sp.fixed_pts, sp.fixed_off)
    h.upload_clouds(LS2D_MOVING, sp.moving_pts, sp.moving_off)
    t = time.time()
    g, gi = h.align_batch(sp.init_xyt, want_iters=True)
    print("gpu align_batch (cold) %.1f ms" % ((time.time() - t) * 1e3))
    t = time.time()
    g, gi = h.align_batch(sp.init_xyt, want_iters=True)
    print("gpu align_batch (warm) %.1f ms" % ((time.time() - t) * 1e3))
    t = time.time()
    o_seq, oi_seq = ob.align_batch(op, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, sp.init_xyt,
                                   n_threads=ob.max_threads())
    print("oracle seq %.1f ms (%d threads)" % ((time.time() - t) * 1e3, ob.max_threads()))
    o_tree, oi_tree = ob.align_batch(op, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, sp.init_xyt,
                                     sum_mode=ob.SUM_TREE, tree_threads=T, n_threads=ob.max_threads())
    for name, o, oi in (("tree", o_tree, oi_tree), ("seq", o_seq, oi_seq)):
        print("---- GPU vs oracle[%s]" % name)
        for f in ("status", "iterations", "n_corr", "n_inliers", "n_kernelized"):
            print("  %-14s equal: %d / %d" % (f, int((g[f] == o[f]).sum()), n))
        for f in ("x", "y", "theta", "chi_inliers", "chi_kernelized"):
            bit = int((g[f].view(np.uint32) == o[f].view(np.uint32)).sum())
            print("  %-14s bit-equal: %d / %d   max|d| = %.3e" % (f, bit, n, np.abs(g[f] - o[f]).max()))
        print("  H bit-equal rows: %d / %d" % (int((g["H"].view(np.uint32) == o["H"].view(np.uint32)).all(1).sum()), n))
        it_ncorr = (gi["n_corr"] == oi["n_corr"]).all(1).sum()
        it_pose = ((gi["x"].view(np.uint32) == oi["x"].view(np.uint32)) & (gi["y"].view(np.uint32) == oi["y"].view(np.uint32)) &
                   (gi["theta"].view(np.uint32) == oi["theta"].view(np.uint32))).all(1).sum()
        print("  all-iteration n_corr equal: %d / %d ; all-iteration pose bit-equal: %d / %d" % (it_ncorr, n, it_pose, n))
        pose_ok = (np.abs(g["x"] - o["x"]) <= 1e-5) & (np.abs(g["y"] - o["y"]) <= 1e-5) & (np.abs(g["theta"] - o["theta"]) <= 1e-6)
        chi_ok = np.abs(g["chi_inliers"] - o["chi_inliers"]) <= 1e-4 * np.abs(o["chi_inliers"])
        print("  pose within 1e-5 m / 1e-6 rad: %d / %d ; chi within 1e-4 rel: %d / %d" % (pose_ok.sum(), n, chi_ok.sum(), n))
    # projector + finder parity
    bad_idx = bad_corr = 0
    for p in range(min(n, 64)):
        f = sp.fixed_pts[sp.fixed_off[p]:sp.fixed_off[p + 1]]
        m = sp.moving_pts[sp.moving_off[p]:sp.moving_off[p + 1]]
        xyt = sp.gt_xyt[p] * 0.7
        fi, mi, fimg, mimg = ob.find_correspondences(op, f, m, xyt)
        gfi, gmi = h.find_correspondences(p, p, xyt)
        bad_corr += not (np.array_equal(fi, gfi) and np.array_equal(mi, gmi))
        idx, depth = h.project(LS2D_FIXED, p, (0.0, 0.0, 0.0))
        bad_idx += not (np.array_equal(idx, fimg["source_idx"]) and np.array_equal(depth.view(np.uint32), fimg["depth"].view(np.uint32)))
        cam = np.array([0.3, -0.2, 0.4], np.float32)
        idx, depth = h.project(LS2D_MOVING, p, cam)
        img = ob.project(op, cam, m)
        bad_idx += not (np.array_equal(idx, img["source_idx"]) and np.array_equal(depth.view(np.uint32), img["depth"].view(np.uint32)))
    print("finder mismatching pairs: %d / %d ; projector mismatching images: %d / %d" % (bad_corr, min(n, 64), bad_idx, 2 * min(n, 64)))
    print("launches:", h.launch_count)


if __name__ == "__main__":
    main()
