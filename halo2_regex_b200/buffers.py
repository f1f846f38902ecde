"""Host buffers of the product package: the neutral layouts of b2r_layout plus the page-locked allocator of the library."""
import ctypes as C

import numpy as np

from b2r_layout.buffers import HostOutputs, aligned_empty, compare_outputs, round_up  # noqa: F401


class PinnedAllocator:
    """Page-locked host memory from the library (b2r_host_alloc / b2r_host_free, include/b2r.h) as numpy arrays — what a
    host without PyTorch uses to make the copies of the host-pointer entry points asynchronous."""

    def __init__(self):
        self._ptrs = []

    def __call__(self, nbytes):
        from ._ffi import last_error, lib
        p = C.c_void_p()
        rc = lib.b2r_host_alloc(max(int(nbytes), 1), C.byref(p))
        if rc != 0:
            raise MemoryError(f"b2r_host_alloc({nbytes}) failed: {last_error()}")
        self._ptrs.append(p.value)
        return np.ctypeslib.as_array((C.c_uint8 * max(int(nbytes), 1)).from_address(p.value))[:nbytes]

    def free(self):
        from ._ffi import lib
        for p in self._ptrs:
            lib.b2r_host_free(p)
        self._ptrs = []
