"""C++ host shim (srrg2_laser_slam_2d_b200/plugin): the reference's config-driven plugin surface.
CPU part: class registry, BOSS-text reader/writer, parameter resolution, mis-wiring exceptions.
GPU part: the reference's own call sequences (apps/visual_test_aligner_2d.cpp:108-156,
apps/visual_test_correspondence_finder_projective_2d.cpp:73-79) through the plugin classes, compared
with the C ABI called directly and with the oracle."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "srrg2_laser_slam_2d_b200", "plugin_test")
OWN_CONFIG = os.path.join(ROOT, "configs", "laser_aligner_b200.json")
MULTI_CONFIG = os.path.join(ROOT, "configs", "multi_laser_aligner_b200.json")
REF_CONFIGS = "/root/reference/configurations"


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "srrg2_laser_slam_2d_b200", "plugin")], check=True)
    return EXE


def run(exe, *args):
    r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return r.stdout


def test_selftest(exe):
    assert "SELFTEST OK" in run(exe, "selftest")


def test_own_config_resolves_to_the_shipped_parameter_values(exe):
    d = json.loads(run(exe, "parse", OWN_CONFIG))
    a = {x["name"]: x for x in d["aligners"]}
    t, l = a["aligner_tracking"], a["aligner_loop"]
    assert (t["max_iterations"], t["point_distance"], t["normal_cos"], t["cauchy_chi_threshold"], t["with_sensor"]) == \
           (10, 0.5, 0.9, 0.01, 1)
    assert (l["max_iterations"], l["point_distance"], l["normal_cos"], l["cauchy_chi_threshold"], l["with_sensor"]) == \
           (30, 1.414, 0.8, 0.05, 0)
    assert t["canvas_cols"] == l["canvas_cols"] == 1081 and d["projectors"] == 1          # shared projector
    assert d["loop_detectors"] == [{"name": "loop_detector", "min_inliers": 300, "max_chi": 0.1, "min_ratio": 0.8,
                                    "aligner": 12}]


@pytest.mark.skipif(not os.path.isdir(REF_CONFIGS), reason="reference checkout not present on this machine")
def test_reference_configurations_load_unchanged(exe):
    """both shipped files instantiate the CUDA-backed classes by their reference names."""
    d = json.loads(run(exe, "parse", os.path.join(REF_CONFIGS, "stage_segway_double_config_LASER_0.json")))
    assert d["objects"] == 45
    a = {x["name"]: x for x in d["aligners"]}
    ld, tr = a["multi_aligner_ld"], a["multi_aligner"]
    assert (ld["max_iterations"], ld["min_num_inliers"], ld["point_distance"], ld["normal_cos"],
            ld["cauchy_chi_threshold"], ld["canvas_cols"], ld["with_sensor"], ld["min_num_correspondences"]) == \
           (30, 10, 1.414, 0.8, 0.05, 721, 0, 0)
    assert (tr["max_iterations"], tr["point_distance"], tr["normal_cos"], tr["cauchy_chi_threshold"],
            tr["with_sensor"], tr["slices"]) == (10, 0.5, 0.9, 0.01, 1, 2)
    assert (tr["range_min"], tr["range_max"], tr["damping"]) == (0.3, 20.0, 0.0)
    assert d["loop_detectors"][0]["min_inliers"] == 300 and d["loop_detectors"][0]["aligner"] == ld["id"]
    pre = {x["name"]: x for x in d["preprocessors"]}          # RawDataPreprocessorProjective2D, L0.json:781-806
    assert pre["ad_scan_0"] == {"name": "ad_scan_0", "scan_topic": "/diago_0/scan_0_0", "voxelize_resolution": 0.02,
                                "range_min": 0.0, "range_max": 1000.0, "normal_point_distance": 0.3,
                                "normal_min_points": 5, "num_ranges": 721}
    m = json.loads(run(exe, "parse", os.path.join(REF_CONFIGS, "stage_segway_double_config_MULTI.json")))
    assert m["objects"] == 56
    am = {x["name"]: x for x in m["aligners"]}
    assert am["multi_aligner_ld"]["min_num_correspondences"] == 10 and am["multi_aligner_ld"]["max_iterations"] == 30
    tr = am["multi_aligner"]                              # two rangefinders + odometry prior in one aligner (:700-730)
    assert (tr["slices"], tr["prior_slices"], len(tr["laser_slices"])) == (3, 1, 2)
    assert [(x["normal_cos"], x["cauchy_chi_threshold"], x["min_num_correspondences"], x["with_sensor"])
            for x in tr["laser_slices"]] == [(0.9, 0.01, 5, 1), (0.8, -1.0, 5, 1)]
    assert m["loop_detectors"][0]["min_inliers"] == 500


def test_own_multi_config_mirrors_the_reference_multi_aligner(exe):
    d = json.loads(run(exe, "parse", MULTI_CONFIG))
    tr = d["aligners"][0]
    assert (tr["name"], tr["slices"], tr["prior_slices"], tr["max_iterations"], tr["canvas_cols"]) == \
           ("multi_aligner", 3, 1, 10, 721)
    assert [(x["point_distance"], x["normal_cos"], x["cauchy_chi_threshold"], x["min_num_correspondences"])
            for x in tr["laser_slices"]] == [(0.5, 0.9, 0.01, 5), (0.5, 0.8, -1.0, 5)]


# ---------------------------------------------------------------------------------------------- GPU
def write_pairs(path, sp, guesses, sensor=(0.0, 0.0, 0.0)):
    n, g = len(guesses), guesses.shape[1]
    with open(path, "wb") as f:
        f.write(struct.pack("<ii3f", n, g, *sensor))
        for p in range(n):
            fx = sp.fixed_pts[sp.fixed_off[p]:sp.fixed_off[p + 1]]
            mv = sp.moving_pts[sp.moving_off[p]:sp.moving_off[p + 1]]
            f.write(struct.pack("<ii", len(fx), len(mv)))
            f.write(np.ascontiguousarray(fx, np.float32).tobytes())
            f.write(np.ascontiguousarray(mv, np.float32).tobytes())
            f.write(np.ascontiguousarray(guesses[p], np.float32).tobytes())


REC = np.dtype([("x", "<f4"), ("y", "<f4"), ("theta", "<f4"), ("chi_inliers", "<f4"), ("chi_kernelized", "<f4"),
                ("status", "<i4"), ("n_inliers", "<i4"), ("n_kernelized", "<i4"), ("n_corr", "<i4"),
                ("iterations", "<i4")])


def same(rec, res):
    for f in rec.dtype.names:
        if not np.array_equal(rec[f], res[f]):
            return f
    return None


@pytest.mark.gpu
@pytest.mark.parametrize("aligner,sensor", [("aligner_tracking", (0.2, 0.2, 0.1)), ("aligner_loop", (0.0, 0.0, 0.0))])
def test_plugin_aligner_matches_abi_and_oracle(exe, tmp_path, handle_factory, oracle, aligner, sensor):
    from srrg2_laser_slam_2d_b200 import default_params
    from srrg2_laser_slam_2d_b200._abi import LS2D_FIXED, LS2D_MOVING, reduction_threads
    from srrg2_laser_slam_2d_b200.synthetic import make_scan_pairs
    n = 6
    sp = make_scan_pairs(n, seed=55)
    guesses = sp.init_xyt[:, None, :].copy()
    inp, out = str(tmp_path / "pairs.bin"), str(tmp_path / "out.bin")
    write_pairs(inp, sp, guesses, sensor)
    assert "ALIGN OK" in run(exe, "align", OWN_CONFIG, aligner, inp, out)
    raw = open(out, "rb").read()
    pos = 0
    single = []
    for p in range(n):
        single.append(np.frombuffer(raw, REC, 1, pos)[0])
        pos += REC.itemsize
        if p == 0:
            k = struct.unpack_from("<i", raw, pos)[0]
            last_corr = np.frombuffer(raw, np.int32, 2 * k, pos + 4).reshape(k, 2)
            pos += 4 + 8 * k
    single = np.array(single, REC)
    batch = np.frombuffer(raw, REC, n, pos)
    pos += n * REC.itemsize
    k = struct.unpack_from("<i", raw, pos)[0]
    finder_corr = np.frombuffer(raw, np.int32, 2 * k, pos + 4).reshape(k, 2)
    track = aligner == "aligner_tracking"
    # the slice reads sensor_in_robot from the tf tree as an isometry (v2t of this triple) and hands it over verbatim
    kw = dict(canvas_cols=1081, point_distance=0.5 if track else 1.414, normal_cos=0.9 if track else 0.8,
              cauchy_chi_threshold=0.01 if track else 0.05, max_iterations=10 if track else 30,
              with_sensor=1 if track else 0, sensor_in_robot=sensor)
    # the plugin is a thin layer: identical to the C ABI called directly ...
    h = handle_factory(default_params(**kw))
    h.upload_clouds(LS2D_FIXED, sp.fixed_pts, sp.fixed_off)
    h.upload_clouds(LS2D_MOVING, sp.moving_pts, sp.moving_off)
    g, gi = h.align_batch(sp.init_xyt, want_iters=True)
    assert same(batch, g) is None
    # compute() hands the estimate back as the Isometry2f the kernel holds (movingInFixed()): no round trip anywhere
    assert same(single, g) is None
    # ... and therefore bit-identical to the oracle in the kernel's summation order
    o, oi = oracle.align_batch(oracle.default_params(**kw), sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off,
                               sp.init_xyt, sum_mode=oracle.SUM_TREE, tree_threads=reduction_threads(1081))
    for f in ("x", "y", "theta", "chi_inliers", "status", "n_inliers", "n_corr", "iterations"):
        assert np.array_equal(batch[f], o[f]), f
    # slice->correspondences() after compute(): the last iteration's list
    assert len(last_corr) == g["n_corr"][0]
    # the finder on its own == oracle finder at the same local_map_in_sensor
    fi, mi, _, _ = oracle.find_correspondences(oracle.default_params(**kw), sp.fixed_pts[:1081], sp.moving_pts[:1081],
                                               sp.init_xyt[0])
    assert np.array_equal(finder_corr[:, 0], fi) and np.array_equal(finder_corr[:, 1], mi)


@pytest.mark.gpu
def test_plugin_loop_detector_verification(exe, tmp_path, handle_factory, oracle):
    from srrg2_laser_slam_2d_b200 import Gates, default_params
    from srrg2_laser_slam_2d_b200._abi import LS2D_FIXED, LS2D_MOVING
    from srrg2_laser_slam_2d_b200.synthetic import make_scan_pairs
    n_cand, n_guess = 10, 3
    sp = make_scan_pairs(n_cand, seed=91, motion_xy=0.3, motion_theta=0.15)
    rng = np.random.default_rng(2)
    guesses = (sp.gt_xyt[0][None, None, :] + rng.uniform(-0.1, 0.1, (n_cand, n_guess, 3))).astype(np.float32)
    inp, out = str(tmp_path / "cands.bin"), str(tmp_path / "out.bin")
    write_pairs(inp, sp, guesses)
    assert "VERIFY OK" in run(exe, "verify", OWN_CONFIG, "loop_detector", inp, out)
    raw = open(out, "rb").read()
    cand, guess, n_inl, n_corr = struct.unpack_from("<4i", raw, 0)
    allr = np.frombuffer(raw, REC, n_cand * n_guess, 16 + 12)
    kw = dict(canvas_cols=1081, point_distance=1.414, normal_cos=0.8, cauchy_chi_threshold=0.05, max_iterations=30)
    h = handle_factory(default_params(**kw))
    h.upload_clouds(LS2D_FIXED, sp.fixed_pts, sp.fixed_off)
    h.upload_clouds(LS2D_MOVING, sp.moving_pts, sp.moving_off)
    # the detector takes its initial guesses as Isometry2f (here v2t of the file's triples) and hands them to the C ABI
    # verbatim (tx, ty, c, s): the same poses as the triples themselves
    best, ref = h.verify(0, None, guesses, Gates(300, 0.1, 0.8), want_all=True)
    assert same(allr, ref) is None
    assert (cand, guess, n_inl) == (int(best["candidate"]), int(best["guess"]), int(best["n_inliers"]))
    assert cand == 0                                       # the true match


@pytest.mark.gpu
def test_plugin_clipper_and_merger_match_the_oracle(exe, tmp_path, oracle):
    """SceneClipperProjective2D / MergerProjective2D through the plugin classes (the reference's call sequence,
    apps/visual_test_merger_projective_2d.cpp:103-123) vs the oracle, bit for bit."""
    from srrg2_laser_slam_2d_b200.synthetic import make_scan_pairs
    n, cols, sensor = 4, 721, (0.2, 0.1, 0.3)
    sp = make_scan_pairs(n, n_beams=721, seed=33)
    poses = (sp.gt_xyt + np.float32(0.02)).astype(np.float32)
    inp, out = str(tmp_path / "map.bin"), str(tmp_path / "map_out.bin")
    write_pairs(inp, sp, poses[:, None, :], sensor)
    assert "MAP OK" in run(exe, "map", str(cols), inp, out)
    raw = open(out, "rb").read()
    pos = 0

    def cloud():
        nonlocal pos
        k = struct.unpack_from("<i", raw, pos)[0]
        c = np.frombuffer(raw, np.float32, 4 * k, pos + 4).reshape(k, 4)
        pos += 4 + 16 * k
        return c

    prm = oracle.default_params(canvas_cols=cols)
    for p in range(n):
        scene = sp.fixed_pts[sp.fixed_off[p]:sp.fixed_off[p + 1]]
        meas = sp.moving_pts[sp.moving_off[p]:sp.moving_off[p + 1]]
        pose, sen = poses[p], np.float32(sensor)      # the plugin holds v2t of these and hands the isometries over verbatim
        clipped, merged = cloud(), cloud()
        ref_c = oracle.clip_scene(prm, scene, pose, sen)
        ref_m, _ = oracle.merge(prm, 0.2, scene, meas, pose)
        assert clipped.shape == ref_c.shape and np.array_equal(clipped.view(np.uint32), ref_c.view(np.uint32))
        assert merged.shape == ref_m.shape and np.array_equal(merged.view(np.uint32), ref_m.view(np.uint32))


@pytest.mark.gpu
def test_plugin_multi_slice_aligner_with_odometry_prior(exe, tmp_path, handle_factory, oracle):
    """MULTI.json-shaped MultiAligner2D (two WithSensor laser slices + AlignerSliceOdom2DPrior) through the plugin
    classes: scenes carry "points_0" / "points_1" / "points" clouds and "odom" poses; compared with the C ABI called
    directly and with the oracle (kernel summation order), bit for bit."""
    from srrg2_laser_slam_2d_b200 import default_params
    from srrg2_laser_slam_2d_b200._abi import make_prior, multi_reduction_threads
    from srrg2_laser_slam_2d_b200.synthetic import make_multi_sensor_pairs
    n = 4
    sensors = ((0.2, 0.05, 0.1), (-0.2, 0.0, 3.0))
    msp = make_multi_sensor_pairs(n, sensors=sensors, n_beams=721, seed=77)
    info = np.array([80.0, 2.0, 0.5, 90.0, -1.0, 300.0], np.float32)
    rng = np.random.default_rng(5)
    odom_fixed = rng.uniform(-2, 2, (n, 3)).astype(np.float32)       # robot pose in the odometry frame now ...
    inp, out = str(tmp_path / "multi.bin"), str(tmp_path / "multi_out.bin")
    odom_moving = np.zeros((n, 3), np.float32)
    z = np.zeros((n, 4), np.float32)
    L = oracle.lib()
    for p in range(n):                                               # ... and the moving scene's: fixed * odometry delta
        M = L.orc_compose(oracle.v2t(*odom_fixed[p]), oracle.v2t(*msp.odom_xyt[p]))
        L.orc_t2v(M, odom_moving[p:p + 1].ctypes.data)
        # what the slice computes: Z = fixed^-1 * moving, handed to the C ABI as the isometry itself
        Z = L.orc_compose(L.orc_inverse(oracle.v2t(*odom_fixed[p])), oracle.v2t(*odom_moving[p]))
        z[p] = oracle.iso_array([Z])[0]
    with open(inp, "wb") as f:
        f.write(struct.pack("<i6f6f", n, *sensors[0], *sensors[1], *info))
        for p in range(n):
            f0 = msp.fixed_pts[0][msp.fixed_off[0][p]:msp.fixed_off[0][p + 1]]
            f1 = msp.fixed_pts[1][msp.fixed_off[1][p]:msp.fixed_off[1][p + 1]]
            mv = msp.moving_pts[msp.moving_off[p]:msp.moving_off[p + 1]]
            f.write(struct.pack("<3i", len(f0), len(f1), len(mv)))
            for c in (f0, f1, mv):
                f.write(np.ascontiguousarray(c, np.float32).tobytes())
            f.write(np.concatenate([msp.init_xyt[p], odom_fixed[p], odom_moving[p]]).astype(np.float32).tobytes())
    assert "MULTI OK" in run(exe, "multi", MULTI_CONFIG, "multi_aligner", inp, out)
    item = np.dtype([("rec", REC), ("nc", "<i4", (2,)), ("H", "<f4", (6,))])
    got = np.frombuffer(open(out, "rb").read(), item).reshape(n, 2)
    # reference run: the C ABI directly; every pose is the isometry the plugin holds (v2t of the file's triples)
    base = dict(canvas_cols=721, max_iterations=10, min_num_correspondences=5, with_sensor=1, point_distance=0.5)
    mk = lambda fac: [fac(normal_cos=0.9, cauchy_chi_threshold=0.01, sensor_in_robot=tuple(map(float, sensors[0])), **base),
                      fac(normal_cos=0.8, cauchy_chi_threshold=-1.0, sensor_in_robot=tuple(map(float, sensors[1])), **base)]
    init_iso = oracle.iso_array([oracle.v2t(*[float(v) for v in x]) for x in msp.init_xyt])
    h = handle_factory()
    h.upload_clouds(0, msp.fixed_pts[0], msp.fixed_off[0])
    h.upload_clouds(2, msp.fixed_pts[1], msp.fixed_off[1])
    h.upload_clouds(1, msp.moving_pts, msp.moving_off)
    fixed = [(msp.fixed_pts[s], msp.fixed_off[s]) for s in range(2)]
    moving = [(msp.moving_pts, msp.moving_off)] * 2
    exact = [f for f in REC.names if f != "theta"]
    for pass_, kw, okw in ((0, dict(prior=make_prior(info), prior_z=z), dict(prior=oracle.make_prior(info), prior_z=z)),
                           (1, {}, {})):
        g = h.align_multi(mk(default_params), [0, 2], [1, 1], init_iso, **kw)
        o, _ = oracle.align_multi_batch(mk(oracle.default_params), fixed, moving, init_iso, sum_mode=oracle.SUM_TREE,
                                        tree_threads=multi_reduction_threads(mk(default_params), 721, 1442, True), **okw)
        rec = got[:, pass_]["rec"]
        assert same(rec[exact], g[exact]) is None
        assert np.array_equal(rec["theta"], g["theta"])              # movingInFixed() is the kernel's isometry itself
        assert np.array_equal(got[:, pass_]["H"].view(np.uint32), g["H"].view(np.uint32))   # informationMatrix()
        for f in ("x", "y", "theta", "chi_inliers", "status", "n_inliers", "n_corr", "iterations"):
            assert np.array_equal(g[f], o[f]), f
        assert (g["status"] == 0).all()
        # slice->correspondences() of both slices: the last iteration's lists
        assert np.array_equal(got[:, pass_]["nc"].sum(1), g["n_corr"])
    # the bound prior changes the answer and is one more inlier factor
    assert not np.array_equal(got[:, 0]["rec"]["x"], got[:, 1]["rec"]["x"])
    assert np.array_equal(got[:, 0]["rec"]["n_inliers"] + got[:, 0]["rec"]["n_kernelized"], got[:, 0]["rec"]["n_corr"] + 1)


@pytest.mark.gpu
@pytest.mark.parametrize("res", [0.02, 0.0])
def test_plugin_raw_data_preprocessor_matches_the_oracle(exe, tmp_path, oracle, res):
    """RawDataPreprocessorProjective2D through the plugin class (the reference's call sequence,
    tests/test_measurement_adaptor.cpp:12-33) vs the oracle, bit for bit; plus the reference's own fixture count."""
    from srrg2_laser_slam_2d_b200.synthetic import make_raw_scans
    raw = make_raw_scans(3, n_beams=721, seed=17)
    scans = [raw.fixed_ranges[0], raw.moving_ranges[1], raw.fixed_ranges[2]]
    inp, out = str(tmp_path / "scan.bin"), str(tmp_path / "scan_out.bin")
    with open(inp, "wb") as f:
        f.write(struct.pack("<iiff", len(scans), 721, raw.angle_min, raw.angle_max))
        for r in scans:
            f.write(np.ascontiguousarray(r, np.float32).tobytes())
    assert "SCAN OK" in run(exe, "scan", str(res), inp, out)
    blob = open(out, "rb").read()
    pos = 0
    osp = oracle.default_scan_params(angle_min=raw.angle_min, angle_max=raw.angle_max, voxelize_resolution=res)
    for r in scans:
        k = struct.unpack_from("<i", blob, pos)[0]
        c = np.frombuffer(blob, np.float32, 4 * k, pos + 4).reshape(k, 4)
        pos += 4 + 16 * k
        ref = oracle.preprocess_scan(osp, r)
        assert c.shape == ref.shape and k > 50 and np.array_equal(c.view(np.uint32), ref.view(np.uint32))
