from .dirichlet import Dirichlet, TruncatedStickBreaking  # noqa: F401
from .wishart import Wishart  # noqa: F401
from .gamma import Gamma  # noqa: F401
from .categorical import Categorical  # noqa: F401
from .matrix import MatrixNormalWithPrecision  # noqa: F401
from .gaussian import (GaussianWithPrecision, StackedGaussiansWithPrecision, TiedGaussiansWithPrecision,  # noqa: F401
                       GaussianWithDiagonalPrecision, StackedGaussiansWithDiagonalPrecision,
                       TiedGaussiansWithDiagonalPrecision, GaussianWithScaledPrecision, TiedGaussiansWithScaledPrecision)
from .lingauss import (LinearGaussianWithPrecision, StackedLinearGaussiansWithPrecision,  # noqa: F401
                       TiedLinearGaussiansWithPrecision, StackedAffineLinearGaussiansWithPrecision,
                       AffineLinearGaussianWithPrecision)
from .composite import (NormalWishart, StackedNormalWisharts, TiedNormalWisharts,  # noqa: F401
                        NormalGamma, StackedNormalGammas, TiedNormalGammas,
                        MatrixNormalWishart, StackedMatrixNormalWisharts, TiedMatrixNormalWisharts)
from .bayesian import (CategoricalWithDirichlet, CategoricalWithStickBreaking,  # noqa: F401
                       GaussianWithNormalWishart, StackedGaussiansWithNormalWisharts, TiedGaussiansWithNormalWisharts,
                       StackedGaussiansWithNormalGammas, TiedGaussiansWithNormalGammas,
                       StackedLinearGaussiansWithMatrixNormalWisharts, TiedLinearGaussiansWithMatrixNormalWisharts,
                       GaussianWithHierarchicalNormalWishart, TiedGaussiansWithHierarchicalNormalWisharts,
                       TiedAffineLinearGaussiansWithMatrixNormalWisharts, AffineLinearGaussianWithMatrixNormalWishart)
