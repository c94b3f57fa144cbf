"""Who is closer to the true gammaincinv(a, p): scipy or the device kernel?  (GPU box; mpmath truth)
    python tools/gamma_truth.py > gpurun_out/gamma_truth.json"""
import json
import os
import sys

import mpmath as mp
import numpy as np
import scipy.special as sc

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_graph_gpu import ppf_device  # noqa: E402
from probabilit_b200.modeling import OP  # noqa: E402

mp.mp.dps = 40
rng = np.random.default_rng(5)
out = {}
for a in (0.05, 0.5, 1.0, 2.5, 30.0):
    q = rng.random(150)
    dev = ppf_device(OP["PPF_GAMMA"], q, a, 0.0, 1.0)
    ref = sc.gammaincinv(a, q)
    truth = []
    for qi, x0 in zip(q, ref):
        f = lambda x: mp.gammainc(a, 0, x, regularized=True) - mp.mpf(float(qi))  # noqa: E731
        try:
            x = mp.findroot(f, (mp.mpf(float(x0)) * (1 - mp.mpf("1e-9")), mp.mpf(float(x0)) * (1 + mp.mpf("1e-9"))),
                            solver="illinois", tol=1e-34)
            x = mp.re(x)
        except Exception:
            x = mp.mpf("nan")
        truth.append(x)
    def ulps(vals):
        e = []
        for v, t in zip(vals, truth):
            if t != t:
                continue
            e.append(abs(float((mp.mpf(float(v)) - t) / mp.mpf(float(np.spacing(abs(float(t))))))))
        return np.array(e)
    ed, es = ulps(dev), ulps(ref)
    out[str(a)] = {"device_vs_truth": {"max": float(ed.max()), "median": float(np.median(ed)), "p95": float(np.percentile(ed, 95))},
                   "scipy_vs_truth": {"max": float(es.max()), "median": float(np.median(es)), "p95": float(np.percentile(es, 95))},
                   "points": int(len(ed))}
print(json.dumps(out, indent=1))
