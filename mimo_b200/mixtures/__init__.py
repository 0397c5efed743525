from .gmm import MixtureOfGaussians, BayesianMixtureOfGaussians  # noqa: F401
from .ilr import MixtureOfLinearGaussians, BayesianMixtureOfLinearGaussians  # noqa: F401
