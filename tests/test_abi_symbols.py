"""libls2d.so loads on a machine without a GPU and exports every symbol include/ls2d.h declares;
without a device the compute entry points fail loudly instead of falling back (CPU only)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from srrg2_laser_slam_2d_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "ls2d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ls2d_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_list_the_same_entry_points():
    assert declared_functions() == sorted(_abi.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = _abi.load()
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_struct_layouts_match_the_header():
    assert C.sizeof(_abi.Params) == 4 * 30
    assert C.sizeof(_abi.Gates) == 12
    assert _abi.RESULT_DTYPE.itemsize == 80 and _abi.ITER_DTYPE.itemsize == 40 and _abi.BEST_DTYPE.itemsize == 48
    p = _abi.default_params()
    assert (p.canvas_cols, p.max_iterations, p.min_num_inliers) == (721, 10, 10)
    assert abs(p.point_distance - 0.5) < 1e-7 and abs(p.normal_cos - 0.8) < 1e-7
    assert abs(p.range_min - 0.3) < 1e-7 and p.range_max == 20.0 and abs(p.cauchy_chi_threshold - 0.01) < 1e-9


def test_host_side_helpers_need_no_gpu():
    lib = _abi.load()
    assert lib.ls2d_version() == 200
    assert lib.ls2d_strerror(0) == b"ok"
    # the 1081-beam shape: 288 threads, two-half warp combine, fused accumulation -- unless the caller asks for
    # single-rounding sums, the point-to-point factor or an option only the general kernel has
    assert _abi.reduction_threads(1081) == 288 | 1 << 16 | 1 << 17
    assert _abi.reduction_threads(1081, params=_abi.default_params(canvas_cols=1081, single_rounding_accumulation=1)) == 288 | 1 << 16
    # the point-to-point factor is a template switch of the same kernel: its shape, single-rounding sums
    assert _abi.reduction_threads(1081, params=_abi.default_params(canvas_cols=1081, factor=_abi.FACTOR_POINT2POINT)) == 288 | 1 << 16
    assert _abi.reduction_threads(721, 721) == 256 | 1 << 16 | 1 << 17
    assert _abi.reduction_threads(1081, params=_abi.default_params(canvas_cols=1081, algorithm=_abi.ALGORITHM_LM)) == 512
    assert _abi.reduction_threads(1081, 4000) == 288                      # canvas wider than the compile-time stride
    assert _abi.reduction_threads(4096, 7680) == 512                      # 247 KB in the register kernel: falls back to streaming
    with pytest.raises(_abi.Ls2dError):
        _abi.reduction_threads(100000)
    rec = np.zeros(4, _abi.BEST_DTYPE)
    rec["candidate"] = [-1, 7, 3, 9]
    rec["guess"] = [-1, 0, 2, 1]
    rec["n_inliers"] = [0, 500, 500, 400]
    rec["chi_inliers"] = [0, 5.0, 5.0, 1.0]
    best = _abi.reduce_best(rec)
    assert best["candidate"] == 3                      # tie on inliers and chi -> lowest candidate id
    assert _abi.reduce_best(rec[:1])["candidate"] == -1


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present")
def test_no_cpu_fallback_without_a_device():
    with pytest.raises(_abi.Ls2dError):
        _abi.Handle(0)


def test_preprocessor_sums_are_not_contracted():
    """ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (a single rounding) when the product has no other use;
    the pre-processor's window walk relies on the product ALSO feeding the distance test (ls2d_scan.cuh).  Its SASS
    must hold packed multiplies and adds but no packed fused multiply-add."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _abi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    body, inside = [], False
    for line in sass.splitlines():
        if "Function : " in line:
            inside = "preprocess_kernel" in line
        elif inside:
            body.append(line)
    text = "\n".join(body)
    assert "FMUL2" in text and "FADD2" in text
    assert "FFMA2" not in text
