"""Matrix-normal distribution with row precision V (o x o) and column precision K (c x c):
vec_F(A) ~ N(vec_F(M), kron(K, V)^-1)  -- the conditional of the expert matrix A in the
Matrix-Normal-Wishart prior (reference: mimo/distributions/matrix.py:10-176; the Gibbs draw of the
linear-Gaussian experts goes through it, composite.py:606-612).

Host-side, per-component parameter algebra only (o*c <= a few hundred): nothing here touches the
per-point path.  Draws consume `numpy.random.normal(size=o*c)` exactly as the reference does, so a
seeded NumPy stream replays.
"""
import numpy as np
import numpy.random as npr
from scipy.linalg import cholesky, solve_triangular

from ..utils.abstraction import Statistics as Stats


class MatrixNormalWithPrecision:

    def __init__(self, column_dim, row_dim, M=None, V=None, K=None):
        self.column_dim, self.row_dim = column_dim, row_dim
        self.M, self.V, self.K = M, V, K

    # -- parameters ---------------------------------------------------------------------------
    @property
    def params(self):
        return self.M, self.V, self.K

    @params.setter
    def params(self, values):
        self.M, self.V, self.K = values

    @property
    def nb_params(self):
        n = self.column_dim * self.row_dim
        return n + n * (n + 1) / 2

    @staticmethod
    def std_to_nat(params):
        """(V M, -V/2) for params = (M, V, ...)   (matrix.py:48-52)."""
        M, V = params[0], params[1]
        return V @ M, -0.5 * V

    @staticmethod
    def nat_to_std(natparam):
        """(-(1/2) b^-1 a, -2 b)   (matrix.py:54-58)."""
        a, b = natparam
        return Stats([-0.5 * np.linalg.solve(b, a), -2. * b])

    @property
    def nat_param(self):
        return self.std_to_nat(self.params)

    @nat_param.setter
    def nat_param(self, natparam):
        self.params = self.nat_to_std(natparam)

    # -- vec form -----------------------------------------------------------------------------
    def _vec(self, A):
        return np.reshape(A, (-1, self.row_dim * self.column_dim), order='F')

    @property
    def lmbda(self):
        return np.kron(self.K, self.V)

    @property
    def V_chol(self):
        return cholesky(self.V, lower=False)

    @property
    def K_chol(self):
        return cholesky(self.K, lower=False)

    @property
    def lmbda_chol(self):
        return cholesky(self.lmbda, lower=False)

    @property
    def lmbda_chol_inv(self):
        n = self.row_dim * self.column_dim
        return solve_triangular(self.lmbda_chol, np.eye(n), lower=False)

    @property
    def sigma(self):
        Ui = self.lmbda_chol_inv
        return Ui @ Ui.T

    def mean(self):
        return self.M

    def mode(self):
        return self.M

    def rvs(self):
        """M + unvec_F(U^-1 z), z = npr.normal(o c), U the upper Cholesky factor of kron(K, V)   (matrix.py:123-125)."""
        z = npr.normal(size=self.row_dim * self.column_dim)
        step = solve_triangular(self.lmbda_chol, z, lower=False)
        return self.M + np.reshape(step, (self.row_dim, self.column_dim), order='F')

    # -- densities ----------------------------------------------------------------------------
    @property
    def base(self):
        return np.power(2. * np.pi, -0.5 * self.row_dim * self.column_dim)

    def log_base(self):
        return np.log(self.base)

    def log_partition(self):
        m = self._vec(self.M)[0]
        return 0.5 * m @ self.lmbda @ m - np.sum(np.log(np.diag(self.lmbda_chol)))

    def log_likelihood(self, x):
        xs = self._vec(np.asarray(x, dtype=np.float64))
        bad = np.isnan(xs).any(axis=1)
        xs = np.where(np.isnan(xs), 0., xs)
        L, m = self.lmbda, self._vec(self.M)[0]
        ll = xs @ (L @ m) - 0.5 * np.einsum('nd,dl,nl->n', xs, L, xs)
        ll[bad] = 0.
        return ll - self.log_partition() + self.log_base()

    def expected_statistics(self):
        m = self._vec(self.M)[0]
        return m, np.outer(m, m) + self.sigma

    def _vec_nat(self):
        """natural parameters of the vec-form Gaussian: (Lambda vec(M), -Lambda / 2)."""
        L = self.lmbda
        return L @ self._vec(self.M)[0], -0.5 * L

    def _dot(self, nat):
        s = self.expected_statistics()
        return np.dot(nat[0], s[0]) + np.tensordot(nat[1], s[1])

    def entropy(self):
        """A(eta) - log h - <eta, E t(x)> with the vec-form natural parameters.  (The reference pairs its (o, c)-shaped
        nat_param with the vec-form statistics and raises on the shape mismatch, matrix.py:160-168; this is the
        quantity that code is after.)"""
        return self.log_partition() - self.log_base() - self._dot(self._vec_nat())

    def cross_entropy(self, dist):
        return dist.log_partition() - dist.log_base() - self._dot(dist._vec_nat())

    def relative_entropy(self, dist):
        """the column-precision part of KL(self || dist) as the reference computes it (matrix.py:170-176)."""
        o, c = self.row_dim, self.column_dim
        kl = 0.5 * np.trace(np.linalg.solve(self.K, dist.K)) - c * o
        kl += o * np.sum(np.log(np.diag(dist.K_chol)))
        kl -= o * np.sum(np.log(np.diag(self.K_chol)))
        return kl
