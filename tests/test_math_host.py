"""The device math of csrc/ls2d_math.cuh, compiled for the host, against glibc (CPU only).

The kernels reproduce the reference's discrete outcomes by carrying operation-for-operation copies of
the libm routines the reference calls (atan2f in the projector, sinf/cosf in geometry2d::v2t).  The
same header builds as plain C++ (libls2d_mathcheck.so); here it must agree with the host libm bit for
bit, and the fast column path must never disagree with the exact one."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from srrg2_laser_slam_2d_b200._abi import MATHCHECK_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mc():
    if not os.path.exists(MATHCHECK_PATH):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "srrg2_laser_slam_2d_b200", "csrc"),
                        "../libls2d_mathcheck.so"], check=True)
    L = C.CDLL(MATHCHECK_PATH)
    vp = C.c_void_p
    L.ls2d_host_atan2f_n.argtypes = [vp, vp, vp, C.c_long]
    L.ls2d_host_sincosf_n.argtypes = [vp, vp, vp, C.c_long]
    L.ls2d_host_polar_column_n.argtypes = [C.c_int, C.c_float, C.c_float, vp, vp, vp, vp, C.c_long]
    L.ls2d_host_polar_column_tiered_n.argtypes = [C.c_int, C.c_float, C.c_float, vp, vp, vp, vp, C.c_long]
    L.ls2d_host_edge_tol.argtypes, L.ls2d_host_edge_tol.restype = [C.c_int, C.c_float, C.c_float], C.c_float
    L.ls2d_host_margin.argtypes, L.ls2d_host_margin.restype = [C.c_int, C.c_float, C.c_float], C.c_float
    L.ls2d_host_atan2f_fast.argtypes, L.ls2d_host_atan2f_fast.restype = [C.c_float, C.c_float], C.c_float
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _bits(a):
    return a.view(np.uint32)


def test_atan2f_copy_matches_libm_bit_for_bit(mc, oracle):
    rng = np.random.default_rng(0)
    n = 4_000_000
    x = rng.uniform(-25, 25, n).astype(np.float32)
    y = rng.uniform(-25, 25, n).astype(np.float32)
    # special operands: axes, zeros, tiny and huge ratios
    x[:8] = [0, 0, 1, -1, 1e-30, 1e30, -0.0, 3]
    y[:8] = [1, -1, 0, 0, 1, 1, 0.0, 3e-38]
    out = np.zeros(n, np.float32)
    mc.ls2d_host_atan2f_n(_p(y), _p(x), _p(out), n)
    ref = oracle.libm_atan2f(y, x)  # the host libm itself (numpy's float32 ufuncs may use SIMD variants)
    assert np.array_equal(_bits(out), _bits(ref))


def test_sincosf_copy_matches_libm_bit_for_bit(mc, oracle):
    rng = np.random.default_rng(1)
    n = 4_000_000
    x = np.concatenate([rng.uniform(-0.8, 0.8, n // 2), rng.uniform(-7, 7, n // 4),
                        rng.uniform(-1e-3, 1e-3, n // 4)]).astype(np.float32)
    s, c = np.zeros(n, np.float32), np.zeros(n, np.float32)
    mc.ls2d_host_sincosf_n(_p(x), _p(s), _p(c), n)
    rs, rc = oracle.libm_sincosf(x)
    assert np.array_equal(_bits(s), _bits(rs))
    assert np.array_equal(_bits(c), _bits(rc))


@pytest.mark.parametrize("cols,amin,amax", [(1081, -3.14159, 3.14159), (721, -3.14159, 3.14159),
                                            (1081, -2.35619, 2.35619), (1024, -1.2566371, 1.2566371),
                                            (4096, -3.14159, 3.14159), (64, -3.14159, 3.14159)])
def test_fast_column_never_disagrees_with_exact(mc, oracle, cols, amin, amax):
    rng = np.random.default_rng(cols)
    n = 3_000_000
    r = rng.uniform(0.3, 20, n)
    a = rng.uniform(-np.pi, np.pi, n)
    # a third of the samples sit right on rounding edges of u = K00*theta + K01
    k00 = cols / (amax - amin)
    edge = (rng.integers(0, cols, n // 3) + 0.5 - cols / 2) / k00 + rng.normal(0, 2e-6, n // 3)
    a[:n // 3] = edge
    x, y = (r * np.cos(a)).astype(np.float32), (r * np.sin(a)).astype(np.float32)
    fast, exact = np.zeros(n, np.int32), np.zeros(n, np.int32)
    mc.ls2d_host_polar_column_n(cols, amin, amax, _p(y), _p(x), _p(fast), _p(exact), n)
    assert np.array_equal(fast, exact)
    # and the exact path is lrintf(K00 * atan2f + K01) as the oracle computes it
    prm = oracle.default_params(canvas_cols=cols, angle_col_min=amin, angle_col_max=amax)
    assert np.array_equal(exact, oracle.column(prm, y, x))


def test_fast_atan2_error_budget(mc):
    """make_polar_cam's margin assumes |atan2f_fast - atan2| <= 1.5e-6 rad."""
    rng = np.random.default_rng(5)
    worst = 0.0
    for _ in range(200000):
        x, y = rng.uniform(-20, 20, 2)
        worst = max(worst, abs(mc.ls2d_host_atan2f_fast(y, x) - np.arctan2(np.float32(y).astype(np.float64),
                                                                            np.float32(x).astype(np.float64))))
    assert worst < 1.2e-6
    assert mc.ls2d_host_margin(1081, -3.14159, 3.14159) < 1e-3


@pytest.mark.parametrize("cols,amin,amax", [(1081, -3.14159, 3.14159), (721, -3.14159, 3.14159),
                                            (1081, -2.35619, 2.35619), (1024, -1.2566371, 1.2566371),
                                            (361, -3.14159, 3.14159)])
def test_tiered_column_never_disagrees_with_exact(mc, cols, amin, amax):
    """icp_fused2_kernel's three tiers (fast proposal -> side of the rounding edge's ray in binary64 -> exact
    atan2f) against the exact path alone, with most samples packed tightly around the rounding edges"""
    rng = np.random.default_rng(cols + 17)
    n = 4_000_000
    r = rng.uniform(0.05, 25, n)
    a = rng.uniform(-np.pi, np.pi, n)
    k00 = np.float32(cols) / (np.float32(amax) - np.float32(amin))
    k01 = np.float32(cols) * np.float32(0.5)
    m = 3 * n // 4
    spread = np.repeat([3e-6, 1e-6, 3e-7, 5e-8], m // 4)          # down to well inside edge_tol
    edge = (rng.integers(-2, cols + 2, m) + 0.5 - float(k01)) / float(k00) + rng.normal(0, 1, m) * spread
    a[:m] = np.clip(edge, -np.pi, np.pi)
    x, y = (r * np.cos(a)).astype(np.float32), (r * np.sin(a)).astype(np.float32)
    col, tier = np.zeros(n, np.int32), np.zeros(n, np.int32)
    fast, exact = np.zeros(n, np.int32), np.zeros(n, np.int32)
    mc.ls2d_host_polar_column_tiered_n(cols, amin, amax, _p(y), _p(x), _p(col), _p(tier), n)
    mc.ls2d_host_polar_column_n(cols, amin, amax, _p(y), _p(x), _p(fast), _p(exact), n)
    assert np.array_equal(col, exact)
    # the second tier does take work off the exact path, even on this edge-packed sample
    n2, n3 = int((tier == 2).sum()), int((tier == 3).sum())
    assert n2 > 0 and n3 > 0
    assert 5e-7 < mc.ls2d_host_edge_tol(cols, amin, amax) < 5e-6
    print("cols %d: tier 2 decided %d, tier 3 (exact) %d of %d near points" % (cols, n2, n3, n2 + n3))
