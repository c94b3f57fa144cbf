import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilit_b200 import _lib
from probabilit_b200.correlation import _IcPlan
n, k = int(float(sys.argv[1])), int(sys.argv[2])
lib = _lib.require_gpu()
X = torch.randn((k, n), device="cuda", dtype=torch.float64)
Y = torch.empty_like(X)
plan = _IcPlan(n, k, 0); plan.set_target(np.eye(k)); h = plan.handle
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def t(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); e1.synchronize(); return e0.elapsed_time(e1)
out = {}
for rep in range(2):
    lib.pbl_ic_stage_begin(h, sp)
    out["rank_scores_batch"] = t(lambda: lib.pbl_ic_stage_rank_scores(h, X.data_ptr(), 1, n, 0, k, sp))
    out["rank_scores_percol"] = t(lambda: [lib.pbl_ic_stage_rank_scores(h, X.data_ptr(), 1, n, c, 1, sp) for c in range(k)])
    out["rank_gather_batch"] = t(lambda: lib.pbl_ic_stage_rank_gather(h, Y.data_ptr(), 1, n, 0, k, sp))
    out["rank_gather_percol"] = t(lambda: [lib.pbl_ic_stage_rank_gather(h, Y.data_ptr(), 1, n, c, 1, sp) for c in range(k)])
print(json.dumps(out))
