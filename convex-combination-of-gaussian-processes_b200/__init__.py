"""ccgp_b200 -- B200-native engine for the data-parallel core of the Combined-GP
reference scripts (batched covariance build + Cholesky + NLL, ME subset
determinants, predictive mean/variance).  All numerics run in libccgp.so
(hand-written sm_100a CUDA behind the C ABI of include/ccgp.h); importing this
package does not need a GPU, creating an Engine does."""
from ._capi import CcgpError, LIB_PATH, SYMBOLS  # noqa: F401
from .engine import (Engine, GAUSS_ISO, GAUSS_ANISO_LAMBDA, GAUSS_ISO_RAW2, MATERN1D, MATERN_SPLINE1D,  # noqa: F401
                     NATURAL, LOGSCALE,
                     MEAN_GLS_BETA, MEAN_ZERO_PLUS_TAU2)
from . import reference_api, workloads, sharding, samplers, me_design  # noqa: F401

__all__ = ["Engine", "CcgpError", "reference_api", "workloads", "sharding", "samplers", "me_design"]
