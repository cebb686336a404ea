"""Loader of tests/golden/*.npz (written by tools/make_golden.py from the oracle)."""
import ast
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALIGN_CASES = ["track_1081", "track_721_l0", "loop_721_l0", "sensor_361", "norobust_361", "iso_721", "p2p_721", "lm_721",
               "lm_p2p_sensor_361", "options_361"]


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["params"] = ast.literal_eval(str(d["params"]))
    return d


def load_raw(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def make_params(factory, d):
    """factory = oracle_binding.default_params or srrg2_laser_slam_2d_b200.default_params"""
    return factory(**d["params"])


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
