from .gmm import MixtureOfGaussians, BayesianMixtureOfGaussians  # noqa: F401
from .ilr import MixtureOfLinearGaussians, BayesianMixtureOfLinearGaussians  # noqa: F401
from .hgmm import (BayesianMixtureOfGaussiansWithHierarchicalPrior, MixtureOfMixtureOfGaussians,  # noqa: F401
                   BayesianMixtureOfMixtureOfGaussians)
from .hilr import (BayesianMixtureOfLinearGaussiansWithTiedActivation, MixtureOfMixtureOfLinearGaussians,  # noqa: F401
                   BayesianMixtureOfMixtureOfLinearGaussians)
