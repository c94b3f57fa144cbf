"""Helpers for the -m gpu tests: drive the C ABI stage by stage and read intermediates back."""
import ctypes as C

import numpy as np

from probabilit_b200 import _lib


def has_gpu():
    try:
        return _lib.load().pbl_device_count() > 0
    except Exception:
        return False


class DeviceArray:
    def __init__(self, host):
        self.lib = _lib.require_gpu()
        self.host_shape = host.shape
        self.nbytes = host.nbytes
        self.ptr = C.c_void_p()
        _lib.check(self.lib.pbl_device_malloc(C.byref(self.ptr), self.nbytes))
        if host.flags.f_contiguous or host.flags.c_contiguous:
            src = host
        else:
            raise ValueError("contiguous arrays only")
        _lib.check(self.lib.pbl_memcpy_h2d(self.ptr, src.ctypes.data, self.nbytes, None))
        _lib.check(self.lib.pbl_stream_synchronize(None))

    def free(self):
        if self.ptr.value:
            self.lib.pbl_device_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def read_device(ptr, shape, dtype=np.float64, order="C"):
    lib = _lib.load()
    out = np.empty(shape, dtype=dtype, order=order)
    _lib.check(lib.pbl_memcpy_d2h(out.ctypes.data, C.c_void_p(ptr), out.nbytes, None))
    _lib.check(lib.pbl_stream_synchronize(None))
    return out


def run_stages(X, C_target, col_batch=0):
    """Run the device pipeline stage by stage; returns dict of intermediates + final status."""
    from probabilit_b200.correlation import _IcPlan

    lib = _lib.require_gpu()
    X = np.asarray(X, dtype=np.float64)
    if not (X.flags.f_contiguous or X.flags.c_contiguous):
        X = np.asfortranarray(X)
    N, K = X.shape
    P = np.linalg.cholesky(C_target)
    plan = _IcPlan(N, K, 0, col_batch)
    plan.set_target(P)
    dX = DeviceArray(X)
    Y = np.empty_like(X)
    dY = DeviceArray(Y)
    rs, cs = (s // 8 for s in X.strides)
    h = plan.handle
    chk = _lib.check
    for attempt in range(2):
        out = _run_stages_once(lib, chk, plan, h, dX, dY, rs, cs, N, K)
        if out["status"] != 6:  # PBL_RETRY: the plan switched to the 64-bit sort, run again
            break
    out["attempts"] = attempt + 1
    _lib.check(lib.pbl_memcpy_d2h(Y.ctypes.data, dY.ptr, Y.nbytes, None))
    _lib.check(lib.pbl_stream_synchronize(None))
    out["result"] = Y
    dX.free()
    dY.free()
    plan.close()
    return out


def _run_stages_once(lib, chk, plan, h, dX, dY, rs, cs, N, K):
    out = {}
    chk(lib.pbl_ic_stage_begin(h, None))
    chk(lib.pbl_ic_stage_rank_scores(h, dX.ptr, rs, cs, 0, K, None))
    out["scores"] = read_device(plan.buffer(0)[0], (N, K), order="F")
    out["sortedX"] = read_device(plan.buffer(1)[0], (N, K), order="F")
    chk(lib.pbl_ic_stage_gram(h, None))
    out["gram"] = read_device(plan.buffer(2)[0], (K, K))
    out["colsum"] = read_device(plan.buffer(3)[0], (K,))
    chk(lib.pbl_ic_stage_solve(h, N, None))
    out["T"] = read_device(plan.buffer(4)[0], (K, K))
    out["Q"] = np.tril(read_device(plan.buffer(5)[0], (K, K)))
    chk(lib.pbl_ic_stage_transform(h, None))
    out["correlated"] = read_device(plan.buffer(0)[0], (N, K), order="F")
    chk(lib.pbl_ic_stage_rank_gather(h, dY.ptr, rs, cs, 0, K, None))
    out["status"] = chk(lib.pbl_ic_stage_status(h, None))
    return out


def ulp_diff(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.spacing(np.maximum(np.abs(a), np.abs(b)))
