"""Developer aid: look-back / tile-life statistics of the digit pass (build with
PBL_EXTRA_NVCC_FLAGS=-DPBL_TILE_STATS).   python tools/tile_stats.py [N] [K] [col_batch]"""
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
cb = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lib = _lib.require_gpu()
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn((k, n), generator=g, device="cuda", dtype=torch.float64)
plan = _IcPlan(n, k, 0, cb)
plan.set_target(np.eye(k))
h, sp = plan.handle, C.c_void_p(torch.cuda.current_stream().cuda_stream)
out = (C.c_ulonglong * 8)()
for rep in range(2):
    lib.pbl_ic_stage_begin(h, sp)
    lib.pbl_debug_tile_stats(out)
    assert lib.pbl_ic_stage_rank_scores(h, X.data_ptr(), 1, n, 0, k, sp) == 0
    lib.pbl_debug_tile_stats(out)
v = [int(x) for x in out]
t = max(v[0], 1)
t6 = max(v[6], 1)
print(json.dumps({"n": n, "k": k, "col_batch": cb, "tiles": v[0], "split_tiles": v[6], "lookback_words_per_tile": v[1] / t6,
                  "polls_unpublished_per_tile": v[2] / t6, "lookback_cycles_per_tile": v[3] / t6,
                  "tile_cycles": v[4] / t, "data_wait_cycles_per_tile": v[5] / t}))
