"""Independent Gamma distributions (API of mimo/distributions/gamma.py:8-113): the
precision factor of the Normal-Gamma posterior of the diagonal family."""
import numpy as np
import numpy.random as npr
from scipy.special import gammaln, digamma

from ..utils.abstraction import Statistics as Stats


class Gamma:

    def __init__(self, dim, alphas, betas):
        self.dim = dim
        self.alphas = alphas   # shape
        self.betas = betas     # rate

    @property
    def params(self):
        return self.alphas, self.betas

    @params.setter
    def params(self, values):
        self.alphas, self.betas = values

    @property
    def nat_param(self):
        return self.std_to_nat(self.params)

    @nat_param.setter
    def nat_param(self, natparam):
        self.params = self.nat_to_std(natparam)

    @staticmethod
    def std_to_nat(params):
        return Stats([params[0] - 1, -params[1]])

    @staticmethod
    def nat_to_std(natparam):
        return natparam[0] + 1, -natparam[1]

    def mean(self):
        return self.alphas / self.betas

    def mode(self):
        assert np.all(self.alphas >= 1.)
        return (self.alphas - 1.) / self.betas

    def rvs(self, size=1):
        return npr.gamma(self.alphas, 1. / self.betas)

    @property
    def base(self):
        return 1.

    def log_base(self):
        return 0.

    def log_partition(self):
        return np.sum(gammaln(self.alphas) - self.alphas * np.log(self.betas))

    def log_likelihood(self, x):
        return np.sum((self.alphas - 1.) * np.log(x) - self.betas * x) - self.log_partition()

    def expected_statistics(self):
        return digamma(self.alphas) - np.log(self.betas), self.alphas / self.betas

    def _dot(self, nat):
        s = self.expected_statistics()
        return np.dot(nat[0], s[0]) + np.dot(nat[1], s[1])

    def entropy(self):
        return self.log_partition() - self._dot(self.nat_param)

    def cross_entropy(self, dist):
        return dist.log_partition() - self._dot(dist.nat_param)
