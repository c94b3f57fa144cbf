"""probabilit_b200 -- the B200-native sampling hot path of tommyod/probabilit.

Public names mirror the reference's (src/probabilit/correlation.py); the compute runs in
libprobabilit_b200.so (hand-written sm_100a CUDA behind a C ABI, include/probabilit_b200.h).
"""
from .correlation import (Cholesky, Correlator, CorrelatorError, ImanConover,  # noqa: F401
                          PermutationCorrelator, nearest_correlation_matrix)

__all__ = ["Cholesky", "Correlator", "CorrelatorError", "ImanConover", "PermutationCorrelator",
           "nearest_correlation_matrix"]
