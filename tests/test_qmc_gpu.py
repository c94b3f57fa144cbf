"""CUDA unit-cube generators: Sobol' and Halton bit-exact with scipy (same seed), Latin hypercube and
Philox streams statistically (KS, moments, exact stratification)."""
import warnings

import numpy as np
import pytest
from scipy import stats
from scipy.stats import qmc

from oracle import qmc as oq

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d,bits,scramble,seed", [(16, 30, False, None), (16, 30, True, 0), (3, 64, True, 7),
                                                  (1, 30, True, 1), (200, 30, True, 3), (5, 12, False, None)])
def test_sobol_bit_exact(d, bits, scramble, seed):
    from probabilit_b200.qmc import Sobol

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = qmc.Sobol(d, scramble=scramble, bits=bits, seed=seed)
        mine = Sobol(d, scramble=scramble, bits=bits, seed=seed)
        np.testing.assert_array_equal(mine._sv, ref._sv.astype(np.uint64))
        np.testing.assert_array_equal(mine._shift, ref._shift.astype(np.uint64))
        for n in (1, 70, 1025, 3):  # odd block boundaries exercise the half-used row pairs
            np.testing.assert_array_equal(mine.random(n), ref.random(n))
        if bits <= 32:  # scipy's own fast_forward rejects 64-bit engines (dtype mismatch in _sobol.pyx)
            ref.fast_forward(1000)
            mine.fast_forward(1000)
            np.testing.assert_array_equal(mine.random(64), ref.random(64))


def test_sobol_generator_argument_spawns_like_scipy():
    """Node.sample passes rng=Generator: scipy spawns a child stream (scipy/stats/_qmc.py:943-946)."""
    from probabilit_b200.qmc import Sobol

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = qmc.Sobol(4, rng=np.random.default_rng(11)).random(33)
        b = Sobol(4, rng=np.random.default_rng(11)).random(33)
    np.testing.assert_array_equal(a, b)


def test_sobol_large_device_resident():
    torch = pytest.importorskip("torch")
    from probabilit_b200.qmc import Sobol

    n, d = 1 << 22, 16
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x = Sobol(d, scramble=True, seed=0).random(n, device=True)
        ref = qmc.Sobol(d, scramble=True, seed=0).random(n)
    assert x.is_cuda and x.shape == (n, d) and x.stride() == (1, n)
    np.testing.assert_array_equal(x.cpu().numpy(), ref)


@pytest.mark.parametrize("d,scramble,seed", [(6, False, None), (20, True, 5), (1, True, 0)])
def test_halton_bit_exact(d, scramble, seed):
    from probabilit_b200.qmc import Halton

    ref = qmc.Halton(d, scramble=scramble, seed=seed)
    mine = Halton(d, scramble=scramble, seed=seed)
    for n in (1, 50, 4099):
        np.testing.assert_array_equal(mine.random(n), ref.random(n))


def test_latin_hypercube_properties():
    from probabilit_b200.qmc import LatinHypercube

    n, d = 100_003, 5
    x = LatinHypercube(d, seed=42).random(n)
    assert x.shape == (n, d) and x.min() > 0 and x.max() <= 1
    for c in range(d):  # exactly one point per stratum
        assert np.array_equal(np.sort(np.ceil(x[:, c] * n).astype(np.int64)), np.arange(1, n + 1))
        assert stats.kstest(x[:, c], "uniform").pvalue > 1e-3
    r = np.corrcoef(x, rowvar=False)
    assert np.abs(r - np.eye(d)).max() < 0.02
    centred = LatinHypercube(2, scramble=False, seed=1).random(10)
    np.testing.assert_allclose(np.sort(centred[:, 0]), (np.arange(10) + 0.5) / 10)
    # reproducible for a given seed, different across seeds
    assert np.array_equal(LatinHypercube(3, seed=5).random(64), LatinHypercube(3, seed=5).random(64))
    assert not np.array_equal(LatinHypercube(3, seed=5).random(64), LatinHypercube(3, seed=6).random(64))
    # the permutation looks random: serial correlation of consecutive rows is small
    y = LatinHypercube(1, seed=9).random(200_000)[:, 0]
    assert abs(np.corrcoef(y[:-1], y[1:])[0, 1]) < 0.01


def test_philox_uniform_moments_and_ks():
    from probabilit_b200.qmc import PhiloxUniform

    n, d = 400_000, 4
    g = PhiloxUniform(d, seed=123)
    x = g.random(n)
    assert x.min() >= 0 and x.max() < 1
    assert abs(x.mean() - 0.5) < 2e-3 and abs(x.var() - 1 / 12) < 1e-3
    for c in range(d):
        assert stats.kstest(x[:, c], "uniform").pvalue > 1e-3
    assert np.abs(np.corrcoef(x, rowvar=False) - np.eye(d)).max() < 0.01
    # continuation = one longer stream (row shards need no communication)
    y = g.random(1000)
    whole = PhiloxUniform(d, seed=123).random(n + 1000)
    np.testing.assert_array_equal(np.vstack([x, y]), whole)
