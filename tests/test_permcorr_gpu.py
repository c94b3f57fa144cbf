"""Parity of the CUDA PermutationCorrelator against the unmodified reference's outputs (golden
vectors) and the oracle.  Bar: the result equals the reference's entry for entry (it is a
re-ordering of the input decided by an accept/reject sequence that must be reproduced exactly)."""
import numpy as np
import pytest
import scipy.stats

from oracle import permutation as op
from test_oracle_permutation import CASES, run_case

pytestmark = pytest.mark.gpu


def device_fn(X, Ct, *, weights, iterations, tol, seed, correlation_type):
    from probabilit_b200 import PermutationCorrelator

    pc = PermutationCorrelator(iterations=iterations, tol=tol, seed=seed, correlation_type=correlation_type)
    return pc.set_target(Ct, weights=weights)(X)


@pytest.mark.parametrize("name", CASES)
def test_output_equals_reference_golden(name):
    got, want = run_case(name, device_fn)
    assert got.shape == want.shape and got.dtype == want.dtype and got.flags.c_contiguous
    np.testing.assert_array_equal(got, want)


def test_chunking_and_rng_state_match_oracle():
    from probabilit_b200 import PermutationCorrelator

    rng = np.random.default_rng(9)
    X = rng.normal(size=(2000, 5))
    A = rng.normal(size=(10, 5))
    Ct = 0.9 * np.corrcoef(A, rowvar=False) + 0.1 * np.eye(5)
    want = op.permutation_correlator(X, Ct, iterations=400, tol=1e-9, seed=4)
    pc = PermutationCorrelator(iterations=400, tol=1e-9, seed=4).set_target(Ct)
    pc._CHUNK_STEPS = 333  # force many launches, chunk boundaries inside the schedule
    np.testing.assert_array_equal(pc(X), want)
    # columns are permutations of the input columns
    np.testing.assert_array_equal(np.sort(want, axis=0), np.sort(X, axis=0))


def test_large_latency_bound_run():
    """README-scale use: N = 1e6, d = 8, composite poisson -> binom columns (BASELINE configs[4] reduced)."""
    from probabilit_b200 import PermutationCorrelator

    rng = np.random.default_rng(0)
    N, K = 1_000_000, 8
    X = np.column_stack([scipy.stats.binom(rng.poisson(3 + k, N), 0.4).ppf(rng.random(N)) for k in range(K)])
    Ct = np.full((K, K), 0.5)
    np.fill_diagonal(Ct, 1.0)
    pc = PermutationCorrelator(seed=0, iterations=200, tol=1e-9).set_target(Ct)
    Y = pc(X)
    np.testing.assert_array_equal(np.sort(Y, axis=0), np.sort(X, axis=0))
    before = pc._error(np.corrcoef(X, rowvar=False), Ct)
    after = pc._error(np.corrcoef(Y, rowvar=False), Ct)
    assert after <= before


def test_errors():
    from probabilit_b200 import PermutationCorrelator

    X = np.random.default_rng(0).normal(size=(50, 2))
    X[:, 1] = 3.0
    with pytest.raises(ValueError, match="constant columns"):
        PermutationCorrelator(seed=0).set_target(np.array([[1, 0.5], [0.5, 1]]))(X)
    with pytest.raises(ValueError):
        PermutationCorrelator(iterations=-1)


def test_infinite_iterations_stop_on_tolerance():
    """iterations=0 means 'until tol' in the reference (correlation.py:644-646): the device loop runs
    chunk after chunk and stops at the same step as the oracle."""
    from probabilit_b200 import PermutationCorrelator

    rng = np.random.default_rng(17)
    X = rng.normal(size=(400, 3))
    Ct = np.array([[1, 0.4, 0.1], [0.4, 1, 0.3], [0.1, 0.3, 1]])
    want = op.permutation_correlator(X, Ct, iterations=0, tol=0.02, seed=2)
    pc = PermutationCorrelator(iterations=0, tol=0.02, seed=2).set_target(Ct)
    pc._CHUNK_STEPS = 300
    got = pc(X)
    np.testing.assert_array_equal(got, want)
    assert pc._error(np.corrcoef(got, rowvar=False), Ct) < 0.02
