"""baynorm_b200 -- B200-native (sm_100a) implementation of bayNorm's data-parallel hot path.

The product is libbaynorm_b200.so (hand-written CUDA behind a C ABI, include/baynorm_b200.h).
This package holds the library, its sources (csrc/) and a host-side mirror of the reference's
R interface (api.py).  Importing the package does not touch the GPU; the first call loads the
library and fails loudly if it, or an sm_100 device, is missing.  There is no CPU fallback.
"""
from .api import (AdjustSIZE_fun, BB_Fun, BetaFun, Check_input, Context, DeviceMatrix, DownSampling, EstPrior,  # noqa: F401
                  EstPrior_rcpp, EstPrior_sprcpp, GradientFun_NB_1D, GradientFun_NB_2D, Main_mean_NB_Bay,
                  Main_mean_NB_spBay, Main_mode_NB_Bay, Main_NB_Bay, MarginalF_NB_1D, MarginalF_NB_2D, Prior_fun,
                  SyntheticControl, bayNorm, bayNorm_sup, default_context, rowMeansFast, rowVarsFast)
from ._lib import BayNormError  # noqa: F401

__version__ = "0.1.0"
