"""srrg2_laser_slam_2d_b200 -- B200-native projective 2D scan-to-local-map registration.

Product code lives in csrc/ (CUDA kernels + the C ABI of include/ls2d.h) and plugin/ (the C++ host shim
that mirrors the reference's config-driven plugin surface).  The Python modules are plumbing for tests
and benchmarks: `_abi` (ctypes binding), `synthetic` (seeded workloads), `sharding` (multi-GPU split)."""
from . import _abi  # noqa: F401
from ._abi import Gates, Handle, Ls2dError, Params, default_params  # noqa: F401

__all__ = ["Handle", "Params", "Gates", "Ls2dError", "default_params"]
