"""Pinned (page-locked) host staging buffers for the async checkpoint path
(north star: "the storage/WAL layer checkpoints shards through pinned async copies")."""
from __future__ import annotations

import ctypes as C

import numpy as np

from quantum_simulations_b200 import _lib as L


class PinnedBuffer:
    def __init__(self, nbytes: int):
        self.lib = L.load()
        self.nbytes = nbytes
        self._ptr = C.c_void_p()
        rc = self.lib.qsv_host_alloc(C.byref(self._ptr), nbytes)
        if rc:
            raise MemoryError(f"cudaHostAlloc({nbytes}) failed ({rc})")
        self._raw = (C.c_uint8 * nbytes).from_address(self._ptr.value)

    @property
    def ptr(self) -> int:
        return self._ptr.value

    def array(self, dtype, count: int | None = None) -> np.ndarray:
        a = np.frombuffer(self._raw, dtype=dtype)
        return a if count is None else a[:count]

    def free(self) -> None:
        if self._ptr.value:
            self._raw = None
            self.lib.qsv_host_free(self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
