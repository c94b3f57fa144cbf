"""Per-stage device times of the Iman-Conover pipeline (CUDA events on the launching stream).

    python tools/stage_times.py [N] [K] [reps] [col_batch]
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilit_b200 import _lib  # noqa: E402
from probabilit_b200.correlation import _IcPlan  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    col_batch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    lib = _lib.require_gpu()
    torch.cuda.init()
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn((k, n), generator=g, device="cuda", dtype=torch.float64)  # column-major (n,k)
    Y = torch.empty_like(X)
    rng = np.random.default_rng(0)
    A = rng.normal(size=(2 * k, k))
    Ct = 0.9 * np.corrcoef(A, rowvar=False) + 0.1 * np.eye(k)
    plan = _IcPlan(n, k, 0, col_batch)
    plan.set_target(np.linalg.cholesky(Ct))
    h = plan.handle
    s = torch.cuda.current_stream().cuda_stream
    sp = C.c_void_p(s)
    stages = [
        ("rank_scores", lambda: lib.pbl_ic_stage_rank_scores(h, X.data_ptr(), 1, n, 0, k, sp)),
        ("gram", lambda: lib.pbl_ic_stage_gram(h, sp)),
        ("solve", lambda: lib.pbl_ic_stage_solve(h, n, sp)),
        ("transform", lambda: lib.pbl_ic_stage_transform(h, sp)),
        ("rank_gather", lambda: lib.pbl_ic_stage_rank_gather(h, Y.data_ptr(), 1, n, 0, k, sp)),
    ]
    out = {"n": n, "k": k, "plan_bytes": int(lib.pbl_ic_plan_bytes(h)), "reps": []}
    for rep in range(reps):
        lib.pbl_ic_stage_begin(h, sp)
        times = {}
        total = 0.0
        for name, fn in stages:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            st = fn()
            e1.record()
            e1.synchronize()
            assert st == 0, (name, st, _lib.last_error())
            times[name] = e0.elapsed_time(e1)
            total += times[name]
        st = lib.pbl_ic_stage_status(h, sp)
        times["total_ms"] = total
        times["status"] = st
        times["samples_vars_per_s"] = n * k / (total * 1e-3)
        times["frac_of_456B_roofline"] = 456.0 * n * k / (total * 1e-3) / 6550.1e9
        out["reps"].append(times)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
