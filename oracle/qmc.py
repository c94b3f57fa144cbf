"""CPU restatements of the unit-cube generators the reference calls (scipy.stats.qmc).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference call sites: src/probabilit/modeling.py:479-489 (``sampler = qmc.X(d=d, rng=random_state);
sampler.random(n)``), README.md:113-114, src/probabilit/correlation.py:14-15.  The algorithms live
in SciPy 1.18.1 (compiled ``_sobol`` / ``_qmc_cy`` extension modules, sources not on the box), so
the published algorithms are restated here and pinned bit-for-bit against ``scipy.stats.qmc`` in
tests/test_oracle_qmc.py:

* Sobol': Joe-Kuo direction numbers (tables: scipy/stats/_sobol_direction_numbers.npz), Gray-code
  construction, LMS + digital-shift scrambling (scipy/stats/_qmc.py:1761-1886).
* Halton: van der Corput radical inverse per prime base with Owen-style digit permutations
  (scipy/stats/_qmc.py:686-800, :1215-1283).
* Latin hypercube: (perm - U) / n (scipy/stats/_qmc.py:1546-1559).
"""
import math
import os

import numpy as np


def _tables():
    import scipy.stats
    path = os.path.join(os.path.dirname(scipy.stats.__file__), "_sobol_direction_numbers.npz")
    z = np.load(path)
    return z["poly"], z["vinit"]


def sobol_direction_numbers(d, bits=30):
    """scipy.stats._sobol._initialize_v: (d, bits) direction numbers, already shifted into place."""
    poly, vinit = _tables()
    sv = np.zeros((d, bits), dtype=np.uint64)
    for i in range(d):
        v = [0] * bits
        if i == 0:
            v = [1] * bits
        else:
            p = int(poly[i])
            m = p.bit_length() - 1
            for j in range(min(m, bits)):
                v[j] = int(vinit[i, j])
            for j in range(m, bits):
                new = v[j - m]
                for k in range(m):
                    if (p >> (m - 1 - k)) & 1:
                        new ^= (2 << k) * v[j - k - 1]
                v[j] = new
        for j in range(bits):
            sv[i, j] = (v[j] << (bits - 1 - j)) & ((1 << 64) - 1)
    return sv


def sobol_scramble(sv, bits, rng):
    """LMS + shift (scipy/stats/_qmc.py:1812-1828 + _sobol._cscramble).  Consumes rng like scipy."""
    d = sv.shape[0]
    dt = np.uint32 if bits <= 32 else np.uint64
    shift_bits = rng.integers(2, size=(d, bits), dtype=dt)
    shift = np.array([sum(int(shift_bits[i, b]) << b for b in range(bits)) for i in range(d)], dtype=np.uint64)
    ltm = np.tril(rng.integers(2, size=(d, bits, bits), dtype=dt))
    out = np.zeros_like(sv)
    for i in range(d):
        rows = []
        for p in range(bits):
            r = 0
            for k in range(bits):  # row p as an integer, column k at bit (bits-1-k); unit diagonal
                bit = 1 if k == p else int(ltm[i, p, k])
                r |= bit << (bits - 1 - k)
            rows.append(r)
        for j in range(bits):
            vdj = int(sv[i, j])
            t2 = 0
            for p in range(bits):  # output bit (bits-1-p) = parity(row_p & v)
                t2 |= (bin(rows[p] & vdj).count("1") & 1) << (bits - 1 - p)
            out[i, j] = t2
    return out, shift


def sobol_points(sv, shift, bits, n, skip=0):
    """Points skip .. skip+n-1: point j = (shift ^ XOR_{b in gray(j)} sv[:, b]) * 2^-bits."""
    d = sv.shape[0]
    out = np.empty((n, d))
    scale = 1.0 / 2 ** bits
    for r in range(n):
        j = skip + r
        g = j ^ (j >> 1)
        q = shift.copy()
        b = 0
        while g:
            if g & 1:
                q ^= sv[:, b]
            g >>= 1
            b += 1
        out[r] = q.astype(np.float64) * scale
    return out


def n_primes(d):
    primes, c = [], 2
    while len(primes) < d:
        if all(c % p for p in primes if p * p <= c):
            primes.append(c)
        c += 1
    return primes


def halton_permutations(bases, rng):
    """scipy/stats/_qmc.py:686-730, one (count, base) table per dimension, consuming rng like scipy."""
    perms = []
    for base in bases:
        count = math.ceil(54 / math.log2(base)) - 1
        p = np.repeat(np.arange(base)[None], count, axis=0)
        for row in p:
            rng.shuffle(row)
        perms.append(p.astype(np.int64))
    return perms


def van_der_corput(n, base, start_index=0, permutations=None):
    """scipy.stats._qmc_cy._cy_van_der_corput(_scrambled)."""
    seq = np.zeros(n)
    for i in range(n):
        quotient = start_index + i
        b2r = 1.0 / base
        acc = 0.0
        if permutations is None:
            while (1.0 - b2r) < 1.0:
                remainder = quotient % base
                acc += remainder * b2r
                b2r /= base
                quotient = (quotient - remainder) // base
        else:
            for j in range(permutations.shape[0]):
                remainder = quotient % base
                acc += float(permutations[j, remainder]) * b2r
                b2r /= base
                quotient = (quotient - remainder) // base
        seq[i] = acc
    return seq


def halton_points(d, n, start_index=0, permutations=None):
    bases = n_primes(d)
    cols = [van_der_corput(n, b, start_index, None if permutations is None else permutations[i])
            for i, b in enumerate(bases)]
    return np.array(cols).T.reshape(n, d)


def latin_hypercube(d, n, rng, scramble=True):
    """scipy/stats/_qmc.py:1546-1559 (strength 1, no optimisation)."""
    samples = rng.uniform(size=(n, d)) if scramble else 0.5
    perms = np.tile(np.arange(1, n + 1), (d, 1))
    for i in range(d):
        rng.shuffle(perms[i, :])
    return (perms.T - samples) / n
