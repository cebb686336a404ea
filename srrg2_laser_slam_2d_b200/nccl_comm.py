"""A raw ncclComm_t for ls2d_verify_sharded_nccl (include/ls2d.h), from Python.

The C ABI takes an existing NCCL communicator as void*.  torch.distributed does not hand its communicators out, so the
tests and bench.py create one here with the NCCL library the process already runs (torch's bundled copy: dlopen by
soname returns the loaded object): rank 0 draws an ncclUniqueId, torch.distributed broadcasts the 128 bytes, every
rank calls ncclCommInitRank.  Plumbing for tests / bench only -- a C++ caller passes the communicator it owns."""
from __future__ import annotations

import ctypes as C


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


class NcclComm:
    def __init__(self, rank: int, world: int, device: int, group=None):
        import torch
        import torch.distributed as dist

        self._lib = C.CDLL("libnccl.so.2")
        self._lib.ncclGetUniqueId.argtypes = [C.POINTER(_UniqueId)]
        self._lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        self._lib.ncclCommDestroy.argtypes = [C.c_void_p]
        self._lib.ncclGetErrorString.argtypes, self._lib.ncclGetErrorString.restype = [C.c_int], C.c_char_p
        uid = _UniqueId()
        if rank == 0:
            self._check(self._lib.ncclGetUniqueId(C.byref(uid)))
        if world > 1 or dist.is_initialized():
            dev = torch.device("cuda", device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            buf = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).to(dev)
            dist.broadcast(buf, src=0, group=group)
            C.memmove(C.byref(uid), bytes(buf.cpu().numpy().tobytes()), 128)
        torch.cuda.set_device(device)
        self.comm = C.c_void_p()
        self._check(self._lib.ncclCommInitRank(C.byref(self.comm), world, uid, rank))
        self.world = world

    def _check(self, rc: int):
        if rc != 0:
            raise RuntimeError("NCCL: " + self._lib.ncclGetErrorString(rc).decode())

    @property
    def ptr(self) -> int:
        return self.comm.value

    def close(self):
        if self.comm:
            self._lib.ncclCommDestroy(self.comm)
            self.comm = C.c_void_p()
