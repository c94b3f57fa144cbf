"""Pin the QMC oracle restatements bit-for-bit against scipy.stats.qmc (the library the reference
calls at src/probabilit/modeling.py:479-489)."""
import warnings

import numpy as np
import pytest
from scipy.stats import qmc

from oracle import qmc as oq


@pytest.mark.parametrize("d,bits", [(1, 30), (16, 30), (40, 30), (5, 64), (3, 12)])
def test_sobol_unscrambled(d, bits):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s = qmc.Sobol(d, scramble=False, bits=bits)
        sv = oq.sobol_direction_numbers(d, bits)
        np.testing.assert_array_equal(sv, s._sv.astype(np.uint64))
        z = np.zeros(d, dtype=np.uint64)
        np.testing.assert_array_equal(oq.sobol_points(sv, z, bits, 70), s.random(70))
        np.testing.assert_array_equal(oq.sobol_points(sv, z, bits, 31, skip=70), s.random(31))


@pytest.mark.parametrize("d,bits,seed", [(16, 30, 0), (7, 30, 123), (3, 64, 5)])
def test_sobol_scrambled(d, bits, seed):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s = qmc.Sobol(d, scramble=True, bits=bits, seed=seed)
        sv, shift = oq.sobol_scramble(oq.sobol_direction_numbers(d, bits), bits, np.random.default_rng(seed))
        np.testing.assert_array_equal(sv, s._sv.astype(np.uint64))
        np.testing.assert_array_equal(shift, s._shift.astype(np.uint64))
        np.testing.assert_array_equal(oq.sobol_points(sv, shift, bits, 33), s.random(33))
        np.testing.assert_array_equal(oq.sobol_points(sv, shift, bits, 20, skip=33), s.random(20))


@pytest.mark.parametrize("d,seed", [(6, 1), (20, 5)])
def test_halton(d, seed):
    np.testing.assert_array_equal(qmc.Halton(d, scramble=False).random(50), oq.halton_points(d, 50))
    h = qmc.Halton(d, scramble=True, seed=seed)
    perms = oq.halton_permutations(oq.n_primes(d), np.random.default_rng(seed))
    np.testing.assert_array_equal(h.random(50), oq.halton_points(d, 50, 0, perms))
    np.testing.assert_array_equal(h.random(10), oq.halton_points(d, 10, 50, perms))


def test_latin_hypercube():
    got = oq.latin_hypercube(3, 20, np.random.default_rng(42))
    np.testing.assert_array_equal(got, qmc.LatinHypercube(d=3, seed=42).random(20))
    # README.md:113 / correlation.py:14-15 configuration
    got = oq.latin_hypercube(2, 100, np.random.default_rng(42), scramble=True)
    np.testing.assert_array_equal(got, qmc.LatinHypercube(d=2, seed=42, scramble=True).random(n=100))
