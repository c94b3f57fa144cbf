"""`nearest_correlation_matrix` without cvxpy: the reference's own tests
(tests/test_correlation.py:7-78) and doctest values (src/probabilit/correlation.py:93-107), restated for the
NumPy implementation.  Host-only: K x K, not on the hot path."""
import numpy as np
import pytest

from probabilit_b200.correlation import nearest_correlation_matrix


@pytest.mark.parametrize("variables", range(2, 100, 10))
def test_solution_is_cholesky_decomposable(variables):
    """reference tests/test_correlation.py:7-35"""
    rng = np.random.default_rng(variables)
    observations = rng.normal(size=(variables * 2, variables))
    matrix = np.corrcoef(observations, rowvar=False)
    np.linalg.cholesky(matrix)
    matrix = matrix + rng.normal(size=matrix.shape, scale=0.1)
    matrix = matrix - np.identity(variables) * np.mean(np.diag(matrix))
    with pytest.raises(np.linalg.LinAlgError):
        np.linalg.cholesky(matrix)
    correlation_matrix = nearest_correlation_matrix(matrix)
    np.linalg.cholesky(correlation_matrix)
    assert np.allclose(np.diag(correlation_matrix), 1.0)
    assert np.allclose(correlation_matrix, correlation_matrix.T)


def test_matlab_nearcorr_example_with_elementwise_weights():
    """reference tests/test_correlation.py:37-78 (matrices from the MATLAB `nearcorr` documentation)"""
    A = np.array([[1.0, 0.0, 0.0, 0.0, -0.936],
                  [0.0, 1.0, -0.55, -0.3645, -0.53],
                  [0.0, -0.55, 1.0, -0.0351, 0.0875],
                  [0.0, -0.3645, -0.0351, 1.0, 0.4557],
                  [-0.936, -0.53, 0.0875, 0.4557, 1.0]])
    W = np.array([[0.0, 1.0, 0.1, 0.15, 0.25],
                  [1.0, 0.0, 0.05, 0.025, 0.15],
                  [0.1, 0.05, 0.0, 0.25, 1.0],
                  [0.15, 0.025, 0.25, 0.0, 0.25],
                  [0.25, 0.15, 1.0, 0.25, 0.0]])
    matlab_Y = np.array([[1.0, 0.0014, 0.0287, -0.0222, -0.8777],
                         [0.0014, 1.0, -0.498, -0.7268, -0.4567],
                         [0.0287, -0.498, 1.0, -0.0358, 0.0878],
                         [-0.0222, -0.7268, -0.0358, 1.0, 0.4465],
                         [-0.8777, -0.4567, 0.0878, 0.4465, 1.0]])
    Y = nearest_correlation_matrix(A, weights=W)
    assert np.allclose(Y, matlab_Y, atol=1e-4)  # MATLAB prints 4 digits
    assert np.linalg.eigvalsh(Y).min() > 0 and np.allclose(np.diag(Y), 1.0) and np.allclose(Y, Y.T)


def test_reference_doctest_values():
    """src/probabilit/correlation.py:93-107"""
    X = np.array([[1.0, 1, 0], [1, 1, 1], [0, 1, 1]])
    Y = nearest_correlation_matrix(X)
    assert abs(Y[0, 1] - 0.76068) < 2e-5 and abs(Y[0, 2] - 0.15729) < 2e-5
    H = np.array([[1, 0.5, 0.1], [0.5, 1, 0.5], [0.1, 0.5, 1]])
    Yw = nearest_correlation_matrix(X, weights=H)
    assert abs(Yw[0, 1] - 0.94171) < 2e-5 and abs(Yw[0, 2] - 0.77365) < 2e-5
    np.linalg.cholesky(Yw)


def test_weighted_solver_agrees_with_higham_for_uniform_weights_and_validates_arguments():
    from probabilit_b200.correlation import _nearest_correlation_matrix_weighted

    rng = np.random.default_rng(5)
    G = np.corrcoef(rng.normal(size=(12, 6)), rowvar=False) + rng.normal(scale=0.2, size=(6, 6))
    G = 0.5 * (G + G.T)
    a = nearest_correlation_matrix(G)
    b = _nearest_correlation_matrix_weighted(G, np.ones_like(G), 10.0 * 1e-6 / 6)
    assert np.allclose(a, b, atol=1e-6)
    with pytest.raises(TypeError):
        nearest_correlation_matrix([[1.0]])
    with pytest.raises(ValueError):
        nearest_correlation_matrix(G, weights=np.ones((2, 2)))
    ok = np.array([[1.0, 0.3], [0.3, 1.0]])
    assert np.array_equal(nearest_correlation_matrix(ok), ok)  # already feasible: returned unchanged
