"""Classifies every pair of the timed config-3 batch (bench.py: 4096 pairs x 1081 beams, seed 0xC0FFEE) on which the
kernel's result and the reference-order result differ in an integer outcome or leave the north_star tolerances.

The kernel is bit-identical to the oracle run in the kernel's summation shape (ORC_SUM_TREE with the shape
ls2d_reduction_shape reports: tests/test_gpu_parity.py), so the comparison needs no GPU: oracle(TREE, kernel shape) vs
oracle(SEQUENTIAL, the reference's order), both with per-iteration records.  Per differing pair: the first iteration
whose integer statistics differ, the first whose pose differs at all, the condition number of the final H, and which
arithmetic (fused D18 / single-rounding) produced it.   usage: python tools/classify_parity.py [out.md]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402
from srrg2_laser_slam_2d_b200.synthetic import make_scan_pairs  # noqa: E402

TRACK = dict(canvas_cols=1081, point_distance=0.5, normal_cos=0.9, cauchy_chi_threshold=0.01, max_iterations=10)
SHAPES = {"fused (default kernel, D18)": 288 | 1 << 16 | 1 << 17, "single-rounding kernel": 288 | 1 << 16}
INT = ("status", "iterations", "n_corr", "n_inliers", "n_kernelized")


def cond3(H6):
    H = np.array([[H6[0], H6[1], H6[2]], [H6[1], H6[3], H6[4]], [H6[2], H6[4], H6[5]]], np.float64)
    w = np.linalg.eigvalsh(H)
    return float(w[-1] / max(w[0], 1e-300))


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02", "parity_classification.md")
    n = 4096
    sp = make_scan_pairs(n, n_beams=1081, seed=0xC0FFEE)
    prm = ob.default_params(**TRACK)
    nt = ob.max_threads()
    seq, seq_it = ob.align_batch(prm, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, sp.init_xyt, n_threads=nt)
    lines = ["# Parity of the timed batch, pair by pair (round 2)", "",
             "Config 3 of BASELINE.json as bench.py times it: 4096 pairs x 1081 beams, 10 iterations, tracking parameters, seed",
             "0xC0FFEE.  Kernel result = oracle in the kernel's summation shape (bit-identical, tests/test_gpu_parity.py);",
             "reference-order result = oracle summing sequentially in correspondence order.  Tolerances of north_star:",
             "1e-5 m, 1e-6 rad, 1e-4 relative chi2.", ""]
    for name, shape in SHAPES.items():
        tree, tree_it = ob.align_batch(prm, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, sp.init_xyt,
                                       sum_mode=ob.SUM_TREE, tree_threads=shape, n_threads=nt)
        ints = np.ones(n, bool)
        for f in INT:
            ints &= tree[f] == seq[f]
        pose = (np.abs(tree["x"] - seq["x"]) <= 1e-5) & (np.abs(tree["y"] - seq["y"]) <= 1e-5) & \
               (np.abs(tree["theta"] - seq["theta"]) <= 1e-6)
        chi = np.abs(tree["chi_inliers"] - seq["chi_inliers"]) <= 1e-4 * np.abs(seq["chi_inliers"]) + 1e-12
        bad = np.flatnonzero(~(ints & pose & chi))
        lines += ["## %s (ORC_SUM_TREE shape 0x%x)" % (name, shape), "",
                  "%d of %d pairs differ in an integer outcome, %d leave a tolerance, %d either (%.2f %% inside everything)."
                  % ((~ints).sum(), n, (~(pose & chi)).sum(), len(bad), 100.0 * (1 - len(bad) / n)), ""]
        cats = {"integer flip (a gate decision of one correspondence changed)": 0,
                "counts equal in every iteration (pose drift: summation order, possibly a swapped z-buffer winner or a moved Cauchy weight)": 0}
        rows = []
        for p in bad:
            first_int = next((k for k in range(prm.max_iterations)
                              if any(tree_it[f][p, k] != seq_it[f][p, k] for f in ("n_corr", "n_inliers", "n_kernelized"))), None)
            first_pose = next((k for k in range(prm.max_iterations)
                               if any(tree_it[f][p, k] != seq_it[f][p, k] for f in ("x", "y", "theta"))), None)
            c = cond3(seq["H"][p])
            flip = first_int is not None
            cats[list(cats)[0 if flip else 1]] += 1
            rows.append((int(p), first_pose, first_int, c, float(np.abs(tree["theta"][p] - seq["theta"][p])),
                         float(max(np.abs(tree["x"][p] - seq["x"][p]), np.abs(tree["y"][p] - seq["y"][p]))),
                         int(tree["n_corr"][p]) - int(seq["n_corr"][p]), int(tree["n_inliers"][p]) - int(seq["n_inliers"][p])))
        conds_all = np.array([cond3(h) for h in seq["H"]])
        lines += ["* %s: %d" % kv for kv in cats.items()]
        lines += ["* condition number of H (lambda_max / lambda_min): median over ALL pairs %.0f; median over the differing pairs %.0f"
                  % (np.median(conds_all), np.median([r[3] for r in rows]) if rows else 0.0), "",
                  "| pair | first iteration with a different pose bit | first iteration with different counts | cond(H) | "
                  "abs d theta [rad] | max abs d x,y [m] | d n_corr | d n_inliers |", "|---|---|---|---|---|---|---|---|"]
        for r in rows:
            lines.append("| %d | %s | %s | %.0f | %.2e | %.2e | %+d | %+d |" % (r[0], r[1], "-" if r[2] is None else r[2], *r[3:]))
        lines.append("")
        print(name, "differing pairs:", len(bad), cats)
    lines += ["Reading: the pose bits part ways in the very first iteration on almost every pair of the batch (the summation order",
              "of H and b moves the solution by a few ulp) -- that alone stays far inside the tolerances: the differing pairs are",
              "NOT ill-conditioned (their cond(H) is the batch's median).  What takes a pair out is a DISCRETE event downstream of",
              "those ulp: a column's gate (|rho_fixed - rho_moving| <= point_distance, normal cosine >= normal_cos), its z-buffer",
              "winner or a factor's Cauchy branch (chi < 0.01) sits within the ulp of its threshold and goes the other way in one",
              "of the ten iterations; the estimate then differs by the weight of one correspondence (1e-6 .. 1e-4 rad / m).  The",
              "fused arithmetic (D18) moves more such events than the single-rounding one because its per-correspondence terms",
              "differ from the reference order's by one rounding each, on top of the summation shape.", ""]
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    open(out_path, "w").write("\n".join(lines))


if __name__ == "__main__":
    main()
