"""Cycle budget of one tile of the radix digit pass by phase (developer tool).  Needs the library
built with instrumentation:  PBL_EXTRA_NVCC_FLAGS=-DPBL_PASS_PROFILE python -m probabilit_b200.build --force
    python tools/pass_phases.py [N] [K]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilit_b200 import _lib  # noqa: E402
from probabilit_b200.correlation import _IcPlan  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
lib = _lib.require_gpu()
X = torch.randn((k, n), device="cuda", dtype=torch.float64)
plan = _IcPlan(n, k, 0)
plan.set_target(np.eye(k))
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for rep in range(2):
    lib.pbl_ic_stage_begin(plan.handle, s)
    lib.pbl_ic_stage_rank_scores(plan.handle, C.c_void_p(X.data_ptr()), 1, n, 0, k, s)
    torch.cuda.synchronize()
    cap = 8192
    buf = np.zeros((8, cap), dtype=np.int64)
    fn = lib.pbl_debug_pass_stamps
    fn.restype = C.c_int
    take = fn(buf.ctypes.data_as(C.POINTER(C.c_longlong)), cap)
buf = buf[:, :take]
names = ["entry->ticket+hist0", "->keys loaded+counted (sync1)", "->scan+publish", "->ballot ranking (sync3)",
         "->smem scatter (+payload load)", "->look-back (sync4)", "->write-out"]
d = np.diff(buf, axis=0)
out = {"tiles_sampled": int(take), "cycles_total_mean": float((buf[7] - buf[0]).mean())}
for i, nm in enumerate(names):
    out[nm] = {"mean": float(d[i].mean()), "p50": float(np.median(d[i])), "p90": float(np.percentile(d[i], 90))}
print(json.dumps(out, indent=1))
