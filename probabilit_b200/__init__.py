"""probabilit_b200 -- the B200-native sampling hot path of tommyod/probabilit.

Public names mirror the reference's package (src/probabilit/__init__.py:1-26, correlation.py); the
compute runs in libprobabilit_b200.so (hand-written sm_100a CUDA behind a C ABI,
include/probabilit_b200.h).  The modeling names are imported lazily (they need networkx).
"""
from .correlation import (Cholesky, Correlator, CorrelatorError, ImanConover,  # noqa: F401
                          PermutationCorrelator, nearest_correlation_matrix)

_MODELING = ("Distribution", "Constant", "EmpiricalDistribution", "CumulativeDistribution",
             "DiscreteDistribution", "Equal", "scalar_transform")

__all__ = ["Cholesky", "Correlator", "CorrelatorError", "ImanConover", "PermutationCorrelator",
           "nearest_correlation_matrix", "PERT", *_MODELING]


def __getattr__(name):
    if name in _MODELING:
        from . import modeling

        return getattr(modeling, name)
    if name == "PERT":
        from .distributions import PERT

        return PERT
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
