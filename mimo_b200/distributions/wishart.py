"""Wishart distribution (API of mimo/distributions/wishart.py:11-153): the scale-matrix
factor of the Normal-Wishart / Matrix-Normal-Wishart posteriors.  Inside the sweep its
Cholesky factor, log-determinant, E[log det] and Bartlett draw live in the batched
posterior kernels; this class is the small host-side object of the API."""
import numpy as np
import numpy.random as npr
from scipy.special import multigammaln, digamma

from ..utils.abstraction import Statistics as Stats


class Wishart:

    def __init__(self, dim, psi=None, nu=None):
        self.dim = dim
        self.psi = psi
        self.nu = nu

    @property
    def params(self):
        return self.psi, self.nu

    @params.setter
    def params(self, values):
        self.psi, self.nu = values

    @property
    def nat_param(self):
        return self.std_to_nat(self.params)

    @nat_param.setter
    def nat_param(self, natparam):
        self.params = self.nat_to_std(natparam)

    @staticmethod
    def std_to_nat(params):
        a = -0.5 * np.linalg.inv(params[0])
        return Stats([a, 0.5 * (params[1] - a.shape[0] - 1)])

    @staticmethod
    def nat_to_std(natparam):
        psi = -0.5 * np.linalg.inv(natparam[0])
        return psi, 2. * natparam[1] + psi.shape[0] + 1

    @property
    def psi_chol(self):
        return np.linalg.cholesky(self.psi)

    def mean(self):
        return self.nu * self.psi

    def mode(self):
        assert self.nu >= (self.dim + 1)
        return (self.nu - self.dim - 1) * self.psi

    def rvs(self, size=1):
        """Bartlett decomposition; variate order as wishart.py:72-80."""
        d = self.dim
        A = np.zeros((d, d))
        A[np.tril_indices(d, k=-1)] = npr.normal(size=d * (d - 1) // 2)
        A[np.diag_indices(d)] = [npr.chisquare(self.nu - i, size=1)[0] ** 0.5 for i in range(d)]
        T = self.psi_chol @ A
        return T @ T.T

    @property
    def base(self):
        return 1.

    def log_base(self):
        return 0.

    def log_partition(self):
        return 0.5 * self.nu * self.dim * np.log(2) + multigammaln(self.nu / 2., self.dim) \
            + self.nu * np.sum(np.log(np.diag(self.psi_chol)))

    def log_likelihood(self, x):
        return 0.5 * (self.nu - self.dim - 1) * np.linalg.slogdet(x)[1] \
            - 0.5 * np.trace(np.linalg.solve(self.psi, x)) - self.log_partition()

    def expected_statistics(self):
        return self.nu * self.psi, np.sum(digamma((self.nu - np.arange(self.dim)) / 2.)) \
            + self.dim * np.log(2.) + 2. * np.sum(np.log(np.diag(self.psi_chol)))

    def _dot(self, nat):
        s = self.expected_statistics()
        return np.tensordot(nat[0], s[0]) + nat[1] * s[1]

    def entropy(self):
        return self.log_partition() - self._dot(self.nat_param)

    def cross_entropy(self, dist):
        return dist.log_partition() - self._dot(dist.nat_param)
