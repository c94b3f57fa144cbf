"""Restatement of Cephes ``ndtri`` (inverse of the standard normal CDF).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference reaches it through ``sp.stats.norm.ppf(ranks)``
(src/probabilit/correlation.py:395) and ``Distribution("norm").ppf``
(src/probabilit/modeling.py:805-812); SciPy's ``norm._ppf`` is ``special.ndtri``
(scipy/stats/_continuous_distns.py:378-379,442), which is the Cephes routine compiled
inside scipy (xsf); its source is not vendored in /root/reference, so the published
Cephes algorithm is restated here (coefficients: SURVEY.md section 8c) and pinned
against ``scipy.special.ndtri`` in tests/test_oracle_golden.py.

``ndtri_scalar`` uses ``math.log``/``math.sqrt`` (glibc, like scipy's C code) and is
bit-identical to scipy; ``ndtri`` is the vectorised NumPy version (np.log may differ
from glibc's log by 1 ulp on a few tail inputs, giving <= 4 ulp differences).
The CUDA kernel (probabilit_b200/csrc/ndtri.cuh) follows the same operation order
with FMA contraction disabled.
"""
import math

import numpy as np

S2PI = 2.50662827463100050242e0
EXPM2 = 0.13533528323661269189  # exp(-2)

P0 = (-5.99633501014107895267e1, 9.80010754185999661536e1, -5.66762857469070293439e1,
      1.39312609387279679503e1, -1.23916583867381258016e0)
Q0 = (1.95448858338141759834e0, 4.67627912898881538453e0, 8.63602421390890590575e1,
      -2.25462687854119370527e2, 2.00260212380060660359e2, -8.20372256168333339912e1,
      1.59056225126211695515e1, -1.18331621121330003142e0)
P1 = (4.05544892305962419923e0, 3.15251094599893866154e1, 5.71628192246421288162e1,
      4.40805073893200834700e1, 1.46849561928858024014e1, 2.18663306850790267539e0,
      -1.40256079171354495875e-1, -3.50424626827848203418e-2, -8.57456785154685413611e-4)
Q1 = (1.57799883256466749731e1, 4.53907635128879210584e1, 4.13172038254672030440e1,
      1.50425385692907503408e1, 2.50464946208309415979e0, -1.42182922854787788574e-1,
      -3.80806407691578277194e-2, -9.33259480895457427372e-4)
P2 = (3.23774891776946035970e0, 6.91522889068984211695e0, 3.93881025292474443415e0,
      1.33303460815807542389e0, 2.01485389549179081538e-1, 1.23716634817820021358e-2,
      3.01581553508235416007e-4, 2.65806974686737550832e-6, 6.23974539184983293730e-9)
Q2 = (6.02427039364742014255e0, 3.67983563856160859403e0, 1.37702099489081330271e0,
      2.16236993594496635890e-1, 1.34204006088543189037e-2, 3.28014464682127739104e-4,
      2.89247864745380683936e-6, 6.79019408009981274425e-9)


def _polevl(x, c):
    """Horner, leading coefficient first (Cephes polevl)."""
    r = c[0]
    for a in c[1:]:
        r = r * x + a
    return r


def _p1evl(x, c):
    """Horner with an implicit leading coefficient of 1 (Cephes p1evl)."""
    r = x + c[0]
    for a in c[1:]:
        r = r * x + a
    return r


def ndtri_scalar(y0):
    """Bit-identical to scipy.special.ndtri for 0 < y0 < 1 (plus the edge values)."""
    if y0 == 0.0:
        return -math.inf
    if y0 == 1.0:
        return math.inf
    if not (0.0 < y0 < 1.0):
        return math.nan
    negate = True
    y = y0
    if y > 1.0 - EXPM2:
        y = 1.0 - y
        negate = False
    if y > EXPM2:
        y = y - 0.5
        y2 = y * y
        x = y + y * (y2 * _polevl(y2, P0) / _p1evl(y2, Q0))
        return x * S2PI
    x = math.sqrt(-2.0 * math.log(y))
    x0 = x - math.log(x) / x
    z = 1.0 / x
    if x < 8.0:
        x1 = z * _polevl(z, P1) / _p1evl(z, Q1)
    else:
        x1 = z * _polevl(z, P2) / _p1evl(z, Q2)
    x = x0 - x1
    if negate:
        x = -x
    return x


def ndtri(y0):
    """Vectorised restatement (NumPy)."""
    y0 = np.asarray(y0, dtype=np.float64)
    out = np.full(y0.shape, np.nan)
    out[y0 == 0.0] = -np.inf
    out[y0 == 1.0] = np.inf
    ok = (y0 > 0.0) & (y0 < 1.0)
    y = np.where(ok, y0, 0.5)
    flip = y > 1.0 - EXPM2
    y = np.where(flip, 1.0 - y, y)
    central = y > EXPM2
    # central branch
    yc = np.where(central, y, 0.5) - 0.5
    y2 = yc * yc
    xc = (yc + yc * (y2 * _polevl(y2, P0) / _p1evl(y2, Q0))) * S2PI
    # tail branch
    yt = np.where(central, 0.1, y)
    x = np.sqrt(-2.0 * np.log(yt))
    x0 = x - np.log(x) / x
    z = 1.0 / x
    x1a = z * _polevl(z, P1) / _p1evl(z, Q1)
    x1b = z * _polevl(z, P2) / _p1evl(z, Q2)
    xt = x0 - np.where(x < 8.0, x1a, x1b)
    xt = np.where(flip, xt, -xt)
    res = np.where(central, xt * 0 + xc, xt)
    # central branch with flip: Cephes returns x*s2pi without negation logic
    # (flip can only be true when y0 > 1-exp(-2), whose mirrored y is < exp(-2): tail)
    out[ok] = res[ok]
    return out
