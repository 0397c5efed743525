"""mimo_b200 -- B200-native mixture-inference sweep behind the mimo.mixtures /
mimo.distributions API.

    from mimo_b200.distributions import StackedNormalWisharts, ...
    from mimo_b200.mixtures import BayesianMixtureOfGaussians, ...

Host code is Python; all per-point and per-component arithmetic runs in hand-written
sm_100a CUDA (mimo_b200/csrc) behind the C-ABI of include/mimo_b200.h.  There is no CPU
fallback: compute entry points raise if libmimo_b200.so is missing or no sm_100 GPU is
visible.  `set_default_precision('fp64')` selects the 1e-9 parity mode.
"""
from ._engine import set_default_precision, default_precision  # noqa: F401

__version__ = '0.1.0'
