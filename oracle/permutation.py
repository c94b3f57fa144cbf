"""CPU restatement of the reference's PermutationCorrelator.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows src/probabilit/correlation.py: ``SwapIndexGenerator`` (:428-470), ``PermutationCorrelator``
(:473-703: subiters schedule :603-617, hill-climbing loop :619-703, weighted errors :597-601 and
:679-686) and ``CorrelationMatrix`` (:757-921: initial matrix :843-853, O(s K) delta :882-907,
commit :861-880).  Same NumPy calls, same order, so the accept / reject sequence is the reference's.

Pinned against: doctests correlation.py:435-446 (swap indices), :517-525 (pearsonr 0.6832...),
:782-817 (CorrelationMatrix values) in tests/test_oracle_permutation.py, and
tests/golden/permcorr_reference.npz (outputs of the unmodified reference,
tests/golden/make_permcorr_golden.py).
"""
import itertools

import numpy as np
import scipy.stats

from .iman_conover import validate_target, validate_X


class SwapIndexGenerator:
    """correlation.py:428-470"""

    def __init__(self, rng, n):
        assert n >= 2
        self.rng = rng
        self.indices = np.arange(n)
        self.permutation = self.rng.permutation(self.indices)

    def __call__(self, size):
        assert size >= 1
        size = min(size, len(self.indices) // 2)
        chosen, self.permutation = self.permutation[: 2 * size], self.permutation[2 * size:]
        if len(chosen) < 2 * size:
            self.permutation = self.rng.permutation(self.indices)
            return self(size)
        return chosen[:size], chosen[size:]


def subiters(n, i):
    """correlation.py:603-617"""
    C = np.log2(n) + 1
    return int(np.ceil(C ** (1 - (2 * i / n))))


class CorrelationMatrix:
    """correlation.py:757-921 (state + incremental update)."""

    def __init__(self, X, correlation_type="pearson"):
        assert correlation_type in ("pearson", "spearman")
        self.correlation_type = correlation_type
        self.X = X.copy()
        if correlation_type == "pearson":
            self.X_ = self.X
        else:
            self.X_ = np.apply_along_axis(scipy.stats.rankdata, axis=0, arr=self.X)
        self.m, self.n = self.X_.shape
        Xc = self.X_ - np.mean(self.X_, axis=0)
        self.numerator = (Xc.T @ Xc) / self.m
        self.denominator = np.std(Xc, axis=0)
        if np.any(np.isclose(self.denominator, 0)):
            raise ValueError("X has one or several constant columns")
        self.corr_mat = (self.numerator / self.denominator[None, :]) / self.denominator[:, None]

    def delta_numerator(self, col, i, j):
        row_i, row_j = self.X_[i, :], self.X_[j, :]
        d = np.sum((row_i - row_j) * (row_j[:, col] - row_i[:, col])[:, None], axis=0)
        d[col] = 0.0
        return d

    def update_column(self, col, i, j):
        delta = self.delta_numerator(col, i, j) / (self.m * self.denominator * self.denominator[col])
        return self.corr_mat[:, col] + delta

    def commit(self, col, i, j):
        dn = self.delta_numerator(col, i, j)
        dc = dn / (self.m * self.denominator * self.denominator[col])
        self.corr_mat[:, col] += dc
        self.corr_mat[col, :] += dc
        self.numerator[:, col] += dn
        self.numerator[col, :] += dn
        self.X_[i, col], self.X_[j, col] = self.X_[j, col], self.X_[i, col]
        if self.correlation_type == "spearman":
            self.X[i, col], self.X[j, col] = self.X[j, col], self.X[i, col]


def rmse(weights, triu, observed, target):
    """_error, correlation.py:597-601"""
    return float(np.sqrt(np.sum(weights[triu] * (observed[triu] - target[triu]) ** 2.0)))


def permutation_correlator(X, C, *, weights=None, iterations=1000, tol=0.01, correlation_type="pearson",
                           seed=None, trace=None):
    """PermutationCorrelator(...).set_target(C, weights=weights)(X) (correlation.py:619-703).
    ``trace`` (optional list) receives (iteration, k, accepted) per step."""
    C, P = validate_target(C)
    w = np.ones_like(C) if weights is None else weights
    w = w / np.sum(w)
    triu = np.triu_indices(C.shape[0], k=1)
    validate_X(X, P, check_rows_cols=False)
    num_obs, num_vars = X.shape
    rng = np.random.default_rng(seed)
    iter_gen = range(1, iterations + 1) if iterations else itertools.count(1)
    swaps = SwapIndexGenerator(rng=rng, n=num_obs)
    cm = CorrelationMatrix(X, correlation_type=correlation_type)
    for iteration in iter_gen:
        num_swaps = subiters(n=iterations if iterations else 10_000, i=iteration)
        for k in range(num_vars):
            i, j = swaps(num_swaps)
            new_col = cm.update_column(k, i, j)
            old_col = cm.corr_mat[k, :]
            target_col = C[k, :]
            old_error = np.average((target_col - old_col) ** 2, weights=w[k, :])
            new_error = np.average((target_col - new_col) ** 2, weights=w[k, :])
            accepted = bool(new_error < old_error)
            if accepted:
                cm.commit(k, i, j)
            if trace is not None:
                trace.append((iteration, k, accepted))
            if k == 0:
                if rmse(w, triu, cm.corr_mat, C) < tol:
                    return cm.X
    return cm.X
