"""Convenience constructors of the reference's src/probabilit/distributions.py, over the device-backed
``Distribution`` of probabilit_b200.modeling.  They only compute scipy parameters on the host (control
plane); sampling runs in the fused graph kernel.

``PERT`` (beta) and ``TruncatedNormal`` (truncnorm) map onto the four-parameter device opcodes
(PBL_PPF_BETA / PBL_PPF_TRUNCNORM, csrc/special.cuh: safeguarded Halley on the incomplete beta, erfc-based
truncated-normal quantile).
"""
import warnings

import numpy as np
import scipy.optimize

from .modeling import Distribution, Exp, Log, Sign


def Uniform(min=0, max=1):
    """Uniform on [min, max) (reference distributions.py:7-9)."""
    return Distribution("uniform", loc=min, scale=max - min)


def Normal(loc, scale):
    """Normal by mean and standard deviation (reference :12-14)."""
    return Distribution("norm", loc=loc, scale=scale)


def TruncatedNormal(loc, scale, low, high):
    """Normal(loc, scale) truncated to [low, high) (reference :17-29)."""
    a, b = (low - loc) / scale, (high - loc) / scale
    return Distribution("truncnorm", a=a, b=b, loc=loc, scale=scale)


class Lognormal(Distribution):
    """Lognormal whose ``mean`` and ``std`` are the moments of the lognormal itself (reference :32-59);
    the parameters may be numbers or nodes (composite)."""

    def __init__(self, mean, std):
        variance = Sign(std) * std ** 2  # keeps the sign, so a negative std fails downstream
        sigma_squared = Log(1 + variance / (mean ** 2))
        sigma = sigma_squared ** (1 / 2)
        mu = Log(mean) - sigma_squared / 2
        super().__init__(distr="lognorm", s=sigma, scale=Exp(mu))

    @classmethod
    def from_log_params(cls, mu, sigma):
        """From the mean / std of log(X) (reference :61-76)."""
        return Distribution("lognorm", s=sigma, scale=Exp(mu))


def pert_to_beta(minimum, mode, maximum, gamma=4.0):
    """(a, b, loc, scale) of the beta distribution behind PERT (reference :187-215)."""
    if not (minimum < mode < maximum):
        raise ValueError(f"Must have {minimum=} < {mode=} < {maximum=}")
    if gamma <= 0:
        raise ValueError(f"Gamma must be positive, got {gamma=}")
    scale = maximum - minimum
    return (1 + gamma * (mode - minimum) / scale, 1 + gamma * (maximum - mode) / scale, minimum, scale)


_pert_to_beta = pert_to_beta


def PERT(minimum, mode, maximum, gamma=4.0):
    """Beta distribution in the PERT parametrisation (reference :79-94)."""
    a, b, loc, scale = pert_to_beta(minimum, mode, maximum, gamma=gamma)
    return Distribution("beta", a=a, b=b, loc=loc, scale=scale)


def fit_triangular_distribution(low, mode, high, low_perc=0.10, high_perc=0.90):
    """(loc, scale, c) of the triangular distribution with the given mode whose ``low_perc`` /
    ``high_perc`` quantiles are ``low`` / ``high`` (reference :135-184): two equations in the support
    end points, solved with fsolve from the reference's starting point."""

    def cdf(x, a, b):
        if x <= a:
            return x * 0
        if x >= b:
            return x * 0 + 1.0
        if x <= mode:
            return ((x - a) ** 2) / ((b - a) * (mode - a))
        return 1 - ((b - x) ** 2) / ((b - a) * (b - mode))

    def residuals(params):
        a, b = params
        return (cdf(low, a, b) - low_perc, cdf(high, a, b) - high_perc)

    start = (low - abs(mode - low), high + abs(high - mode))
    a, b = scipy.optimize.fsolve(residuals, start)
    rmse = np.sqrt(np.sum(np.array(residuals([a, b])) ** 2))
    if rmse > 1e-6:
        warnings.warn(f"Optimization of Triangular params has {rmse=}")
    return float(a), float(b - a), float((mode - a) / (b - a))


_fit_triangular_distribution = fit_triangular_distribution


def Triangular(low, mode, high, low_perc=0.1, high_perc=0.9):
    """``Distribution("triang", ...)`` from (low, mode, high) given as percentiles (reference :97-132)."""
    if not (low < mode < high):
        raise ValueError(f"Must have {low=} < {mode=} < {high=}")
    if not ((0 <= low_perc <= 1.0) and (0 <= high_perc <= 1.0)):
        raise ValueError("Percentiles must be between 0 and 1.")
    if np.isclose(low_perc, 0.0) and np.isclose(high_perc, 1.0):
        loc, scale, c = low, high - low, (mode - low) / (high - low)
    else:
        loc, scale, c = fit_triangular_distribution(low, mode, high, low_perc=low_perc, high_perc=high_perc)
    return Distribution("triang", loc=loc, scale=scale, c=c)
