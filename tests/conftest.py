import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_binding

    oracle_binding.lib()
    return oracle_binding


REF_SO = os.path.join(ROOT, "oracle", "_ref", "libls2d_ref.so")
REF_SRC = "/root/reference/srrg2_laser_slam_2d/src"


@pytest.fixture(scope="session")
def ref(oracle):
    """oracle/_ref/libls2d_ref.so: the reference's OWN finder / merger / clipper / pre-processor sources, compiled
    where they lie under /root/reference against the stand-in headers of oracle/ref_shim (oracle/Makefile).  Rebuilt
    here when the reference checkout is present; the GPU box uses the prebuilt file that travels with the repo."""
    import ctypes as C
    import subprocess

    if os.path.isdir(REF_SRC):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libls2d_ref.so not built and /root/reference absent")
    oracle.lib()  # libls2d_oracle.so first: the reference library resolves the upstream stand-ins from it
    L = C.CDLL(REF_SO)
    vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
    for name in ("ref_find_correspondences", "ref_find_correspondences_iso"):
        getattr(L, name).argtypes = [vp, vp, i32, vp, i32, vp, i32, vp, vp]
        getattr(L, name).restype = i32
    for name in ("ref_merge", "ref_merge_iso"):
        getattr(L, name).argtypes = [vp, f32, vp, i32, vp, i32, vp]
        getattr(L, name).restype = i32
    for name in ("ref_clip", "ref_clip_iso"):
        getattr(L, name).argtypes = [vp, vp, i32, vp, vp, f32, vp]
        getattr(L, name).restype = i32
    L.ref_preprocess_scan.argtypes = [vp, vp, i32, vp]
    L.ref_preprocess_scan.restype = i32
    L.ref_finder_throws_without_inputs.restype = i32
    return L


@pytest.fixture(scope="session")
def handle_factory():
    """Creates ls2d handles on cuda:0; fails loudly (no skip, no fallback) when CUDA is unavailable."""
    from srrg2_laser_slam_2d_b200 import Handle

    made = []

    def make(params=None):
        h = Handle(0, params)
        made.append(h)
        return h

    yield make
    for h in made:
        h.close()
