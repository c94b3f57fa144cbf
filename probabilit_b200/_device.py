"""Device buffers owned through the C ABI (no torch needed): fp64 columns in HBM."""
import ctypes as C

import numpy as np

from . import _lib


class DeviceColumns:
    """``k`` contiguous fp64 columns of ``n`` rows each, laid out ``[k][n]`` in HBM -- the
    column-major (n, k) matrix the graph path hands to a correlator (reference
    src/probabilit/modeling.py:580).  Exposes ``__cuda_array_interface__`` (shape (k, n)), so
    ``torch.as_tensor(cols, device="cuda")`` / CuPy view it without a copy."""

    def __init__(self, n, k):
        self.lib = _lib.require_gpu()
        self.n, self.k = int(n), int(k)
        self.nbytes = self.n * self.k * 8
        self._ptr = C.c_void_p()
        _lib.check(self.lib.pbl_device_malloc(C.byref(self._ptr), max(self.nbytes, 16)), "pbl_device_malloc")

    @property
    def ptr(self):
        if not self._ptr.value:
            raise _lib.PblError("device buffer already released")
        return self._ptr.value

    def column_ptr(self, j):
        if not 0 <= j < self.k:
            raise IndexError(j)
        return self.ptr + j * self.n * 8

    @classmethod
    def from_host(cls, arr):
        """(n, k) host array (any layout / real dtype) -> device columns."""
        arr = np.asarray(arr)
        if arr.ndim != 2:
            raise ValueError("expected a 2-D array")
        cols = np.asfortranarray(arr, dtype=np.float64)
        out = cls(arr.shape[0], arr.shape[1])
        if cols.nbytes:
            _lib.check(out.lib.pbl_memcpy_h2d(out._ptr, cols.ctypes.data, cols.nbytes, None), "pbl_memcpy_h2d")
            _lib.check(out.lib.pbl_stream_synchronize(None))
        return out

    def to_host(self):
        """-> (n, k) float64 array, column-major."""
        out = np.empty((self.n, self.k), dtype=np.float64, order="F")
        if out.nbytes:
            _lib.check(self.lib.pbl_memcpy_d2h(out.ctypes.data, C.c_void_p(self.ptr), out.nbytes, None), "pbl_memcpy_d2h")
            _lib.check(self.lib.pbl_stream_synchronize(None))
        return out

    def column_to_host(self, j):
        out = np.empty(self.n, dtype=np.float64)
        if out.nbytes:
            _lib.check(self.lib.pbl_memcpy_d2h(out.ctypes.data, C.c_void_p(self.column_ptr(j)), out.nbytes, None),
                       "pbl_memcpy_d2h")
            _lib.check(self.lib.pbl_stream_synchronize(None))
        return out

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.k, self.n), "typestr": "<f8", "data": (self.ptr, False), "version": 3,
                "strides": None}

    def free(self):
        if getattr(self, "_ptr", None) is not None and self._ptr.value:
            self.lib.pbl_device_free(self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def as_device_columns(q):
    """Accept what a caller may hold for an (n, k) matrix: DeviceColumns, a CUDA torch tensor /
    any ``__cuda_array_interface__`` object laid out column-major, or a host array (uploaded).
    Returns (columns, keepalive)."""
    if isinstance(q, DeviceColumns):
        return q, q
    cai = getattr(q, "__cuda_array_interface__", None)
    if cai is not None and not isinstance(q, np.ndarray):
        shape, strides = tuple(cai["shape"]), cai.get("strides")
        if cai["typestr"] != "<f8" or len(shape) != 2:
            raise TypeError("device quantiles must be a 2-D float64 array")
        n, k = shape
        if strides is None:
            strides = (k * 8, 8)
        if strides != (8, n * 8) and not (k == 1 and strides[0] == 8):
            # not column-major: let the owner library transpose (torch only)
            if hasattr(q, "t") and hasattr(q, "contiguous"):
                qc = q.t().contiguous().t()
                return as_device_columns(qc)
            raise TypeError("device quantiles must be column-major (strides (1, n) elements)")
        # the library's graph / correlator entry points run on the default stream: whatever produced this
        # memory on the caller's current stream (torch: possibly a non-default one) must have finished
        if type(q).__module__.startswith("torch"):
            import torch

            torch.cuda.current_stream(q.device).synchronize()
        view = DeviceColumns.__new__(DeviceColumns)
        view.lib = _lib.require_gpu()
        view.n, view.k, view.nbytes = int(n), int(k), int(n) * int(k) * 8
        view._ptr = C.c_void_p(int(cai["data"][0]))
        view.free = lambda: None  # borrowed memory
        return view, q
    cols = DeviceColumns.from_host(q)
    return cols, cols
