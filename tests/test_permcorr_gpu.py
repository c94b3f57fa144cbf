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


@pytest.mark.parametrize("n,k", [(1, 1), (7, 3), (1000, 8), (4097, 33), (65, 70)])
@pytest.mark.parametrize("src_order,dst_order", [("F", "C"), ("C", "F"), ("C", "C"), ("F", "F")])
def test_strided_copy_converts_between_layouts(n, k, src_order, dst_order):
    """pbl_copy_strided_f64: the device-side layout conversion around the correlators that work column-major
    inside (the result of PermutationCorrelator goes back in C order like the reference's X.copy(),
    correlation.py:830-831), against NumPy on every pair of layouts, tile-edge shapes included."""
    import ctypes as C

    import gpu_util
    from probabilit_b200 import _lib

    lib = _lib.require_gpu()
    rng = np.random.default_rng(n * 100 + k)
    A = np.asarray(rng.normal(size=(n, k)), order=src_order)
    src = gpu_util.DeviceArray(A)
    dst = gpu_util.DeviceArray(np.zeros((n, k), order=dst_order))
    srs, scs = (A.strides[0] // 8, A.strides[1] // 8)
    drs, dcs = ((k, 1) if dst_order == "C" else (1, n))
    _lib.check(lib.pbl_copy_strided_f64(src.ptr, srs, scs, dst.ptr, drs, dcs, n, k, None), "pbl_copy_strided_f64")
    out = np.empty((n, k), order=dst_order)
    _lib.check(lib.pbl_memcpy_d2h(out.ctypes.data, dst.ptr, out.nbytes, None))
    _lib.check(lib.pbl_stream_synchronize(None))
    np.testing.assert_array_equal(out, A)
    src.free()
    dst.free()
