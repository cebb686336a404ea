"""Generates tests/golden/*.npz: small seeded inputs plus the ORACLE's outputs for them.

The reference holds no golden vectors for this path and cannot be built here, so these
fixtures freeze the oracle restatement (oracle/ls2d_oracle.c) instead: the CPU suite checks that the
oracle still reproduces them, the GPU suite checks the CUDA path against them.  Re-run only when a
decision point of the oracle changes on purpose:   python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402
from srrg2_laser_slam_2d_b200.synthetic import (make_multi_sensor_pairs, make_raw_scans, make_scan_pairs,  # noqa: E402
                                                reference_demo_scene)

OUT = os.path.join(ROOT, "tests", "golden")

# name -> (generator kwargs, oracle params)
CASES = {
    "track_1081": (dict(n_pairs=4, n_beams=1081, seed=101),
                   dict(canvas_cols=1081, normal_cos=0.9, max_iterations=10)),
    "track_721_l0": (dict(n_pairs=6, n_beams=721, seed=102),   # LASER_0.json tracking aligner values
                     dict(canvas_cols=721, normal_cos=0.9, point_distance=0.5, cauchy_chi_threshold=0.01,
                          max_iterations=10)),
    "loop_721_l0": (dict(n_pairs=6, n_beams=721, seed=103, motion_xy=0.3, motion_theta=0.15,
                         init_noise_xy=0.15, init_noise_theta=0.05),  # LASER_0.json loop-closure aligner values
                    dict(canvas_cols=721, normal_cos=0.8, point_distance=1.414, cauchy_chi_threshold=0.05,
                         max_iterations=30)),
    "sensor_361": (dict(n_pairs=6, n_beams=361, seed=104),
                   dict(canvas_cols=361, normal_cos=0.9, max_iterations=8, with_sensor=1,
                        sensor_in_robot=(0.2, 0.2, 0.1))),          # synthetic_scene_generator.cpp:77
    "norobust_361": (dict(n_pairs=6, n_beams=361, seed=105),
                     dict(canvas_cols=361, normal_cos=0.8, cauchy_chi_threshold=-1.0, max_iterations=6,
                          min_num_correspondences=5)),              # MULTI.json laser_1 slice: no robustifier
    # poses handed over as Isometry2f content (tx, ty, c, s): initial guesses and sensor_in_robot are PRODUCTS of
    # increments, the way a tracker accumulates them -- their (c, s) are not the cosf / sinf of any angle
    "iso_721": (dict(n_pairs=6, n_beams=721, seed=108),
                dict(canvas_cols=721, normal_cos=0.9, max_iterations=10, with_sensor=2, sensor_in_robot="accumulated")),
    # options north_star names and the shipped configurations leave off (oracle decisions D19, L1-L8, I1, T1)
    "p2p_721": (dict(n_pairs=6, n_beams=721, seed=109),
                dict(canvas_cols=721, normal_cos=0.9, max_iterations=10, factor=1)),
    "lm_721": (dict(n_pairs=6, n_beams=721, seed=110, motion_xy=0.3, motion_theta=0.15, init_noise_xy=0.15,
                    init_noise_theta=0.05),
               dict(canvas_cols=721, normal_cos=0.8, point_distance=1.414, cauchy_chi_threshold=0.05, max_iterations=15,
                    algorithm=1)),
    "lm_p2p_sensor_361": (dict(n_pairs=6, n_beams=361, seed=111),
                          dict(canvas_cols=361, normal_cos=0.9, max_iterations=8, with_sensor=1,
                               sensor_in_robot=(0.2, 0.2, 0.1), algorithm=1, factor=1, lm_variable_damping=0,
                               lm_user_lambda_init=0.5)),
    "options_361": (dict(n_pairs=6, n_beams=361, seed=112),
                    dict(canvas_cols=361, normal_cos=0.9, max_iterations=8, enable_inlier_only_runs=1,
                         termination_epsilon=1e-3)),
}


def params_dict(p):
    d = {}
    for k, _ in p._fields_:
        v = getattr(p, k)
        d[k] = list(v) if hasattr(v, "__len__") else v
    return d


def accumulated_pose(seed, xyt, steps=20):
    """an Isometry2f close to v2t(xyt) built as a product of `steps` increments (tests/oracle_binding.accumulate)"""
    rng = np.random.default_rng(seed)
    inc = np.tile(np.asarray(xyt, np.float64) / steps, (steps, 1)) + rng.normal(0.0, 2e-3, (steps, 3))
    return ob.accumulate(inc.astype(np.float32))


# MULTI.json tracking aligner (:700-730): al_sl_laser_0 (Cauchy 0.01, finder 0.5 / 0.9), ad_sl_odom, al_sl_laser_1
# (no robustifier, finder 0.5 / 0.8), min_num_correspondences 5, 721 columns, 10 iterations
MULTI_SENSORS = ((0.2, 0.05, 0.1), (-0.2, 0.0, 3.1415927))
MULTI_PRIOR_INFO = (100.0, 0.0, 0.0, 100.0, 0.0, 400.0)


def multi_slices(factory, sensors):
    base = dict(canvas_cols=721, max_iterations=10, min_num_correspondences=5, with_sensor=1, point_distance=0.5)
    return [factory(normal_cos=0.9, cauchy_chi_threshold=0.01, sensor_in_robot=tuple(float(v) for v in sensors[0]), **base),
            factory(normal_cos=0.8, cauchy_chi_threshold=-1.0, sensor_in_robot=tuple(float(v) for v in sensors[1]), **base)]


def make_multi():
    msp = make_multi_sensor_pairs(4, sensors=MULTI_SENSORS, n_beams=721, seed=106)
    sl = multi_slices(ob.default_params, msp.sensors)
    fixed = [(msp.fixed_pts[s], msp.fixed_off[s]) for s in range(2)]
    moving = [(msp.moving_pts, msp.moving_off)] * 2
    res, its = ob.align_multi_batch(sl, fixed, moving, msp.init_xyt, prior=ob.make_prior(MULTI_PRIOR_INFO),
                                    prior_z=msp.odom_xyt)
    res_np, its_np = ob.align_multi_batch(sl, fixed, moving, msp.init_xyt)
    np.savez_compressed(os.path.join(OUT, "multi_721_mu.npz"), fixed_pts_0=msp.fixed_pts[0], fixed_pts_1=msp.fixed_pts[1],
                        fixed_off=msp.fixed_off[0], moving_pts=msp.moving_pts, moving_off=msp.moving_off,
                        sensors=msp.sensors, init_xyt=msp.init_xyt, gt_xyt=msp.gt_xyt, odom_xyt=msp.odom_xyt,
                        prior_info=np.array(MULTI_PRIOR_INFO, np.float32), results=res, iters=its,
                        results_no_prior=res_np, iters_no_prior=its_np)
    print("multi_721_mu status", res["status"], "n_corr", res["n_corr"], "n_inl", res["n_inliers"],
          "err", np.abs(np.stack([res["x"], res["y"], res["theta"]], 1) - msp.gt_xyt).max())


def make_mapping():
    """tracker_721_l0.npz: the rows around the aligner with the LASER_0.json values -- raw 721-beam
    scans -> RawDataPreprocessorProjective2D (voxelize 0.02 and off) -> SceneClipperProjective2D of a local map ->
    MergerProjective2D of the measurement into it; plus the reference's own Synthetic fixture (100 points)."""
    raw = make_raw_scans(3, n_beams=721, seed=107)
    kw = dict(angle_min=raw.angle_min, angle_max=raw.angle_max)
    sp_vox, sp_full = ob.default_scan_params(**kw), ob.default_scan_params(voxelize_resolution=0.0, **kw)
    prm = ob.default_params(canvas_cols=721)
    out = dict(fixed_ranges=raw.fixed_ranges, moving_ranges=raw.moving_ranges,
               angles=np.array([raw.angle_min, raw.angle_max], np.float32), gt_xyt=raw.gt_xyt,
               robot_in_local_map=np.array([[0.01, -0.02, 0.015]] * 3, np.float32),
               sensor_in_robot=np.array([0.2, 0.2, 0.1], np.float32))   # synthetic_scene_generator.cpp:77
    for k in range(3):
        meas = ob.preprocess_scan(sp_vox, raw.fixed_ranges[k])
        scene = ob.preprocess_scan(sp_full, raw.moving_ranges[k])
        clip = ob.clip_scene(prm, scene, out["robot_in_local_map"][k], out["sensor_in_robot"])
        merged, counters = ob.merge(prm, 0.2, scene, meas, raw.gt_xyt[k])
        out.update({f"meas_{k}": meas, f"scene_{k}": scene, f"clip_{k}": clip, f"merged_{k}": merged,
                    f"merge_counters_{k}": counters})
        print("tracker_721_l0[%d]: %d beams -> %d (voxel 0.02) / %d (full); clip %d; merge %s -> %d" %
              (k, 721, len(meas), len(scene), len(clip), counters, len(merged)))
    fx = ob.default_scan_params(angle_min=-1.0, angle_max=1.0, msg_range_min=0.0, msg_range_max=1000.0, range_min=0.0,
                                range_max=1000.0, voxelize_resolution=0.01)  # tests/fixtures.hpp:38-47
    out["synthetic_fixture_cloud"] = ob.preprocess_scan(fx, np.full(100, 1.0, np.float32))
    assert len(out["synthetic_fixture_cloud"]) == 100                      # tests/test_measurement_adaptor.cpp:36
    np.savez_compressed(os.path.join(OUT, "tracker_721_l0.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--mapping-only" in sys.argv:
        return make_mapping()
    if "--multi-only" in sys.argv:
        return make_multi()
    make_multi()
    make_mapping()
    for name, (gen, prm_kw) in CASES.items():
        sp = make_scan_pairs(**gen)
        prm_kw = dict(prm_kw)
        iso_case = prm_kw.get("sensor_in_robot") == "accumulated"
        if iso_case:
            prm_kw["sensor_in_robot"] = ob.iso_array([accumulated_pose(7, (0.2, 0.2, 0.1))])[0]
        prm = ob.default_params(**prm_kw)
        init = sp.init_xyt
        if iso_case:
            init = ob.iso_array([accumulated_pose(100 + p, sp.init_xyt[p] + np.float32([0.02, -0.01, 0.01]))
                                 for p in range(sp.n_pairs)])
        res, its = ob.align_batch(prm, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, init)
        fidx, midx, ncorr, f_src, f_depth, m_src, m_depth = [], [], [], [], [], [], []
        for p in range(sp.n_pairs):
            f = sp.fixed_pts[sp.fixed_off[p]:sp.fixed_off[p + 1]]
            m = sp.moving_pts[sp.moving_off[p]:sp.moving_off[p + 1]]
            lmis = ob.as_iso(init[p])
            if prm.with_sensor == 2:
                S = ob.OrcIso(prm.sensor_in_robot[0], prm.sensor_in_robot[1], *prm.sensor_in_robot_cs)
                lmis = ob.compose(ob.inverse(S), lmis)
            elif prm.with_sensor:
                lmis = ob.compose(ob.inverse(ob.v2t(*prm.sensor_in_robot)), lmis)
            fi, mi, fimg, mimg = ob.find_correspondences(prm, f, m, lmis)
            pad = np.full(prm.canvas_cols, -1, np.int32)
            a, b = pad.copy(), pad.copy()
            a[:len(fi)], b[:len(mi)] = fi, mi
            fidx.append(a), midx.append(b), ncorr.append(len(fi))
            f_src.append(fimg["source_idx"]), f_depth.append(fimg["depth"])
            m_src.append(mimg["source_idx"]), m_depth.append(mimg["depth"])
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"), fixed_pts=sp.fixed_pts, fixed_off=sp.fixed_off,
            moving_pts=sp.moving_pts, moving_off=sp.moving_off, init_xyt=init, gt_xyt=sp.gt_xyt,
            params=np.array(repr(params_dict(prm))), results=res, iters=its,
            corr_fixed_idx=np.stack(fidx), corr_moving_idx=np.stack(midx), corr_n=np.array(ncorr, np.int32),
            fixed_source_idx=np.stack(f_src), fixed_depth=np.stack(f_depth),
            moving_source_idx=np.stack(m_src), moving_depth=np.stack(m_depth), lmis_is_init=np.array(1))
        print(name, "status", res["status"], "n_corr", res["n_corr"], "n_inl", res["n_inliers"], "it", res["iterations"],
              "lm_rej", res["lm_rejected"], "err",
              np.abs(np.stack([res["x"], res["y"], res["theta"]], 1) - sp.gt_xyt).max(0))

    # the reference's own deterministic demo world (apps/synthetic_scene_generator.cpp:36-88):
    # 1024 bins over +-0.4 pi, range_min 0.01, projector at (0.2, 0.2, 0.1) in the robot, robot at identity
    scene = reference_demo_scene()
    prm = ob.default_params(canvas_cols=1024, angle_col_min=np.float32(-np.pi * 0.4), angle_col_max=np.float32(np.pi * 0.4),
                            range_min=0.01)
    img = ob.project(prm, (0.2, 0.2, 0.1), scene)
    np.savez_compressed(os.path.join(OUT, "demo_scene_projection.npz"), scene=scene,
                        params=np.array(repr(params_dict(prm))), camera_pose=np.array([0.2, 0.2, 0.1], np.float32),
                        source_idx=img["source_idx"], depth=img["depth"])
    print("demo scene: %d / 1024 bins hit, depth range %.3f..%.3f" % ((img["source_idx"] >= 0).sum(),
          img["depth"][img["source_idx"] >= 0].min(), img["depth"][img["source_idx"] >= 0].max()))


if __name__ == "__main__":
    main()
