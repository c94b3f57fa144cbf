"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/src/probabilit with empty ``cvxpy`` / ``seaborn`` modules stubbed into
sys.modules (neither is installed here and neither is on the hot path),
runs the reference's ImanConover on seeded inputs and stores inputs + outputs as small .npz
fixtures next to this script.  /root/reference does not exist on the GPU box, so the tests only
ever read the committed .npz files.
"""
import os
import sys
import types

import numpy as np
import scipy as sp
import scipy.stats

HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name in ("cvxpy", "seaborn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, "/root/reference/src")
    import probabilit.correlation as rc
    return rc


def random_target(rng, K):
    """Recipe of reference tests/test_iman_conover.py:154-155."""
    A = rng.normal(size=(2 * K, K))
    return 0.9 * np.corrcoef(A, rowvar=False) + 0.1 * np.eye(K)


def ic_cases():
    """name -> (X, C).  Small enough to commit, varied enough to cover the tie semantics."""
    cases = {}
    # config 1: README.md:112-129 / correlation.py:13-30
    sampler = sp.stats.qmc.LatinHypercube(d=2, seed=42, scramble=True)
    u = sampler.random(n=100)
    X = np.vstack((sp.stats.triang(0.5).ppf(u[:, 0]), sp.stats.gamma.ppf(u[:, 1], a=1))).T
    cases["readme_lhs"] = (X, np.array([[1, 0.3], [0.3, 1]]))
    # correlation.py:315-330 toy matrix (ties in both columns)
    X = np.array([[0, 0], [0, 0.5], [0, 1], [1, 0], [1, 0.5], [1, 1]], dtype=float)
    cases["toy_ties"] = (X, np.array([[1, 0.7], [0.7, 1]]))
    # correlation.py:347-361
    cases["normal_1000x2"] = (np.random.default_rng(42).normal(size=(1000, 2)),
                              np.array([[1, 0.7], [0.7, 1]]))
    cases["lognormal_1000x2"] = (np.random.default_rng(42).lognormal(size=(1000, 2)),
                                 np.array([[1, 0.7], [0.7, 1]]))
    # config 3 at reduced N: mixed norm / triang / gamma marginals on scrambled Sobol, d=16
    rng = np.random.default_rng(0)
    d, n = 16, 4096
    u = sp.stats.qmc.Sobol(d=d, seed=0, scramble=True).random(n)
    X = np.empty((n, d), order="F")
    for k in range(d):
        if k % 3 == 0:
            X[:, k] = sp.stats.norm(loc=1, scale=2).ppf(u[:, k])
        elif k % 3 == 1:
            X[:, k] = sp.stats.triang(0.5).ppf(u[:, k])
        else:
            X[:, k] = sp.stats.gamma(a=2).ppf(u[:, k])
    cases["sobol_mixed_4096x16"] = (X, random_target(rng, d))
    # discrete data: heavy ties in every column, ties in the correlated scores too
    rng = np.random.default_rng(1)
    X = rng.poisson(3.0, size=(2000, 3)).astype(float)
    cases["poisson_2000x3"] = (X, random_target(rng, 3))
    # +-0.0, +-inf, denormals, duplicates and negative numbers
    rng = np.random.default_rng(2)
    X = rng.normal(size=(500, 4))
    X[::7, 0] = 0.0
    X[3::7, 0] = -0.0
    X[5, 1] = np.inf
    X[6, 1] = -np.inf
    X[10:20, 2] = 5e-324 * np.arange(10)
    X[20:30, 2] = -5e-324 * np.arange(10)
    X[::5, 3] = np.round(X[::5, 3], 1)
    cases["specials_500x4"] = (X, random_target(rng, 4))
    # a wider one (config 4 at reduced size) and a C-ordered input
    rng = np.random.default_rng(3)
    X = np.ascontiguousarray(rng.normal(size=(700, 64)))
    cases["wide_700x64"] = (X, random_target(rng, 64))
    return cases


def main():
    rc = import_reference()
    out = {}
    for name, (X, C) in ic_cases().items():
        Y = rc.ImanConover().set_target(C)(X)
        out[name + "__X"] = X
        out[name + "__C"] = C
        out[name + "__Y"] = Y
        print(f"{name:24s} X{X.shape} -> pearson[0,1] = {np.corrcoef(Y, rowvar=False)[0, 1]:.8f}")
    np.savez_compressed(os.path.join(HERE, "ic_reference.npz"), **out)
    print("wrote ic_reference.npz", os.path.getsize(os.path.join(HERE, "ic_reference.npz")), "bytes")


if __name__ == "__main__":
    main()
