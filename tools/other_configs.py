"""Timings of the BASELINE.json configurations that are not the bench line (parity-test cases):
config 1 (README ImanConover), 2 (mutual-fund graph), 4 (wide IC), 5 (PermutationCorrelator on composite
poisson -> binom columns).  GPU box:  python tools/other_configs.py > gpurun_out/other_configs.json"""
import json
import os
import sys
import time

import numpy as np
import scipy.stats as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from oracle import iman_conover as oic  # noqa: E402
from oracle import permutation as operm  # noqa: E402
from probabilit_b200 import ImanConover, PermutationCorrelator, _lib  # noqa: E402
import probabilit_b200.modeling as m  # noqa: E402


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts), out


out = {}
# config 1: README ImanConover, LatinHypercube d=2 n=100 seed=42 (reference README.md:112-129)
u = st.qmc.LatinHypercube(d=2, seed=42, scramble=True).random(n=100)
X = np.vstack((st.triang(0.5).ppf(u[:, 0]), st.gamma.ppf(u[:, 1], a=1))).T
C = np.array([[1, 0.3], [0.3, 1]])
ic = ImanConover().set_target(C)
t, Y = best(lambda: ic(X))
out["config1_readme_ic"] = {"n": 100, "d": 2, "ms": t * 1e3, "pearson": float(st.pearsonr(*Y.T).statistic),
                            "equals_oracle": bool(np.array_equal(Y, oic.iman_conover(X, C)))}

# config 4: wide Iman-Conover N=1e6 d=1024, device resident
n, d = 1_000_000, 1024
g = torch.Generator(device="cuda").manual_seed(0)
Xd = torch.randn((d, n), generator=g, device="cuda", dtype=torch.float64).T
A = np.random.default_rng(0).normal(size=(2 * d, d))
Cw = 0.9 * np.corrcoef(A, rowvar=False) + 0.1 * np.eye(d)
icw = ImanConover().set_target(Cw)
icw(Xd)
t, Yd = best(lambda: icw(Xd), reps=2)
out["config4_wide_ic"] = {"n": n, "d": d, "ms": t * 1e3, "samples_vars_per_s": n * d / t,
                          "col0_unchanged": bool(torch.equal(Xd[:, 0], Yd[:, 0]))}
del Xd, Yd, icw
torch.cuda.empty_cache()

# config 5: PermutationCorrelator N=1e7 d=8, composite poisson -> binom columns (README bird survival)
n, d = 10_000_000, 8
cols = []
for k in range(d):
    eggs = m.Distribution("poisson", mu=3 + k)
    surv = m.Distribution("binom", n=eggs, p=0.4)
    t0 = time.perf_counter()
    cols.append(surv.sample(n, random_state=k, gc_strategy=[]))
Xp = np.column_stack(cols)
Cp = np.full((d, d), 0.5)
np.fill_diagonal(Cp, 1.0)
pc = PermutationCorrelator(seed=0, iterations=1000, tol=1e-9).set_target(Cp)
t0 = time.perf_counter()
Yp = pc(Xp)
t_dev = time.perf_counter() - t0
t0 = time.perf_counter()
Yo = operm.permutation_correlator(Xp, Cp, iterations=1000, tol=1e-9, seed=0)
t_cpu = time.perf_counter() - t0
out["config5_permcorr"] = {"n": n, "d": d, "iterations": 1000, "device_s_incl_h2d_d2h_and_host_swap_stream": t_dev,
                           "cpu_port_s": t_cpu, "equals_cpu_port": bool(np.array_equal(Yp, Yo)),
                           "error_before": pc._error(np.corrcoef(Xp, rowvar=False), Cp),
                           "error_after": pc._error(np.corrcoef(Yp, rowvar=False), Cp)}
print(json.dumps(out, indent=1))
