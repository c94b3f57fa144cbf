"""Golden vectors for PermutationCorrelator / CorrelationMatrix from the UNMODIFIED reference
(build container only):  python tests/golden/make_permcorr_golden.py"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    rng = np.random.default_rng(42)
    out = {}
    out["doctest_100x2"] = dict(X=rng.normal(size=(100, 2)), C=np.array([[1, 0.7], [0.7, 1]]), seed=0,
                                iterations=1000, tol=0.01, correlation_type="pearson", weights=None)
    K = 6
    A = rng.normal(size=(2 * K, K))
    C = 0.9 * np.corrcoef(A, rowvar=False) + 0.1 * np.eye(K)
    out["pearson_500x6"] = dict(X=rng.lognormal(size=(500, K)), C=C, seed=3, iterations=300, tol=1e-9,
                                correlation_type="pearson", weights=None)
    W = rng.uniform(0.5, 2.0, size=(K, K))
    W = W + W.T
    out["weighted_400x6"] = dict(X=rng.normal(size=(400, K)), C=C, seed=5, iterations=200, tol=1e-9,
                                 correlation_type="pearson", weights=W)
    Xp = np.column_stack([rng.poisson(3.0, 600).astype(float), rng.normal(size=600), rng.exponential(size=600)])
    out["spearman_ties_600x3"] = dict(X=Xp, C=np.array([[1, 0.5, 0.2], [0.5, 1, -0.3], [0.2, -0.3, 1]]), seed=11,
                                      iterations=150, tol=1e-9, correlation_type="spearman", weights=None)
    out["early_stop_300x3"] = dict(X=rng.normal(size=(300, 3)), C=np.array([[1, 0.3, 0], [0.3, 1, 0.2], [0, 0.2, 1]]),
                                   seed=1, iterations=2000, tol=0.02, correlation_type="pearson", weights=None)
    K = 20
    A = rng.normal(size=(2 * K, K))
    out["wide_60x20"] = dict(X=rng.normal(size=(60, K)), C=0.9 * np.corrcoef(A, rowvar=False) + 0.1 * np.eye(K),
                             seed=2, iterations=100, tol=1e-9, correlation_type="pearson", weights=None)
    return out


def main():
    for name in ("cvxpy", "seaborn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, "/root/reference/src")
    import probabilit.correlation as rc

    store = {}
    for name, c in cases().items():
        pc = rc.PermutationCorrelator(iterations=c["iterations"], tol=c["tol"], seed=c["seed"],
                                      correlation_type=c["correlation_type"])
        pc.set_target(c["C"], weights=c["weights"])
        Y = pc(c["X"])
        store[f"{name}__X"], store[f"{name}__C"], store[f"{name}__Y"] = c["X"], c["C"], Y
        store[f"{name}__params"] = np.array([c["seed"], c["iterations"], c["tol"],
                                             1.0 if c["correlation_type"] == "spearman" else 0.0])
        if c["weights"] is not None:
            store[f"{name}__W"] = c["weights"]
        moved = int(np.sum(Y != c["X"]))
        print(name, Y.shape, "entries moved:", moved, "corr01", np.corrcoef(Y, rowvar=False)[0, 1])
    np.savez_compressed(os.path.join(HERE, "permcorr_reference.npz"), **store)


if __name__ == "__main__":
    main()
