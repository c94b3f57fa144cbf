"""Pin the CPU oracle: reference doctest goldens + vectors produced by the unmodified reference."""
import numpy as np
import pytest
import scipy as sp
import scipy.special
import scipy.stats

from oracle import iman_conover as oic
from oracle.ndtri import ndtri, ndtri_scalar

CASES = ["readme_lhs", "toy_ties", "normal_1000x2", "lognormal_1000x2", "sobol_mixed_4096x16",
         "poisson_2000x3", "specials_500x4", "wide_700x64"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_equals_reference_output(ic_golden, name):
    X, C, Y = ic_golden[name]
    np.testing.assert_array_equal(oic.iman_conover(X, C), Y)


@pytest.mark.parametrize("name", CASES)
def test_explicit_stages_equal_reference_output(ic_golden, name):
    X, C, Y = ic_golden[name]
    st = oic.iman_conover_stages(X, C)
    np.testing.assert_array_equal(st["result"], Y)
    # T is upper triangular, and scores @ T reproduces the two-step trsm + gemm to rounding
    assert np.allclose(np.tril(st["T"], -1), 0.0)
    assert np.allclose(st["correlated_fused"], st["correlated"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(oic.corrcoef_restated(st["scores"]), st["R"], rtol=0, atol=1e-14)


def test_reference_doctest_goldens(ic_golden):
    """README.md:112-129, correlation.py:13-30, :315-338, :347-361."""
    X, C, _ = ic_golden["readme_lhs"]
    assert abs(sp.stats.pearsonr(*X.T).statistic - 0.06589800) < 5e-9
    Y = oic.iman_conover(X, C)
    assert abs(sp.stats.pearsonr(*Y.T).statistic - 0.27965287) < 5e-9
    X, C, _ = ic_golden["toy_ties"]
    Y = oic.iman_conover(X, C)
    np.testing.assert_array_equal(
        Y, np.array([[0, 0], [0, 0], [0, 0.5], [1, 0.5], [1, 1], [1, 1]], dtype=float))
    assert sp.stats.pearsonr(*Y.T).statistic.round(6) == 0.816497
    X, C, _ = ic_golden["normal_1000x2"]
    assert sp.stats.pearsonr(*oic.iman_conover(X, C).T).statistic.round(6) == 0.697701
    X, C, _ = ic_golden["lognormal_1000x2"]
    assert sp.stats.pearsonr(*oic.iman_conover(X, C).T).statistic.round(6) == 0.592541


def test_tie_semantics():
    """rankdata(...).astype(int) - 1 == run start + (run length - 1) // 2 (correlation.py:422)."""
    rng = np.random.default_rng(5)
    x = rng.integers(0, 7, size=1000).astype(float)
    x[::13] = -0.0
    x[5::13] = 0.0
    np.testing.assert_array_equal(oic.midpoint_index(x), sp.stats.rankdata(x).astype(int) - 1)
    r, _ = oic.average_ranks(x)
    np.testing.assert_array_equal(r, sp.stats.rankdata(x))


def test_errors_match_reference_conventions():
    rng = np.random.default_rng(0)
    X = rng.normal(size=(100, 3))
    with pytest.raises(ValueError):  # not symmetric (tests/test_iman_conover.py:49-61)
        oic.iman_conover(X, np.array([[1.0, 0.7, -0.3], [0.8, 1.0, 0.5], [-0.3, 0.5, 1.0]]))
    with pytest.raises(ValueError):  # not PD (:85-96)
        oic.iman_conover(X, np.array([[1.0, 2.0, 0.3], [2.0, 1.0, 0.2], [0.3, 0.2, 1.0]]))
    with pytest.raises(ValueError):  # wrong size (:98-109)
        oic.iman_conover(X, np.array([[1.0, 0.5], [0.5, 1.0]]))
    with pytest.raises(ValueError):  # unity rank correlation (:200-210)
        oic.iman_conover(np.array([[1.0, 1], [2.0, 1.1], [2.1, 3]]), np.identity(2))
    with pytest.raises(TypeError):
        oic.iman_conover([[1.0, 2.0]], np.identity(2))
    Xn = X.copy()
    Xn[3, 1] = np.nan
    with pytest.raises(ValueError):
        oic.iman_conover(Xn, np.identity(3))


def test_ndtri_restatement_is_bit_identical_to_scipy():
    rng = np.random.default_rng(0)
    q = np.concatenate([rng.random(20000), rng.random(5000) * 1e-6, 1 - rng.random(5000) * 1e-6,
                        [1e-300, 1e-17, 0.5, 0.13533528323661269189, 1 - 0.13533528323661269189]])
    ref = sp.special.ndtri(q)
    got = np.array([ndtri_scalar(float(v)) for v in q])
    np.testing.assert_array_equal(got, ref)
    vec = ndtri(q)
    ulp = np.abs(vec - ref) / np.spacing(np.abs(ref))
    assert ulp.max() <= 4
    assert np.mean(vec != ref) < 1e-3
    assert ndtri_scalar(0.0) == -np.inf and ndtri_scalar(1.0) == np.inf


def test_cholesky_correlator_doctest_goldens():
    """Reference doctest correlation.py:216-238 (values quoted there)."""
    C = np.array([[1, 0.7], [0.7, 1]])
    X = np.random.default_rng(4).normal(size=(9, 2))
    assert sp.stats.pearsonr(*X.T).statistic.round(6) == -0.025582
    Y = oic.cholesky_correlator(X, C)
    assert sp.stats.pearsonr(*Y.T).statistic.round(6) == 0.7
    np.testing.assert_allclose(np.mean(Y, axis=0), [-0.63531692, 0.70114825], atol=5e-9)
    np.testing.assert_allclose(np.std(Y, axis=0), [1.11972638, 0.75668173], atol=5e-9)
