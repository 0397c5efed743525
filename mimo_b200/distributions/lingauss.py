"""Linear-Gaussian likelihoods y = A [x; 1] + eps (API of mimo/distributions/lingauss.py,
full-precision variants; arithmetic in libmimo_b200.so).

On the device a point is z = [x | y] and zt = [x | y | 1]; the affine input
xt = [x ; 1] therefore lives in columns 0..d_in-1 and D = d_in + d_out of zt."""
import numpy as np
import numpy.random as npr

from .. import _engine as E
from ..utils.abstraction import Statistics as Stats
from .gaussian import _clean_rows, soft_stats_quad


class ExpertLayout:
    """index maps of a linear-Gaussian expert inside zt = [x | y | 1]."""

    def __init__(self, column_dim, row_dim, affine=True):
        self.c, self.o, self.affine = column_dim, row_dim, affine
        self.din = column_dim - 1 if affine else column_dim
        self.D = self.din + row_dim
        self.x_cols = list(range(self.din)) + ([self.D] if affine else [])
        self.y_cols = list(range(self.din, self.D))
        self._dev = None

    def dev(self):
        if self._dev is None:
            self._dev = dict(stat_idx=E._i32(self.x_cols + self.y_cols + [self.D]),   # xt, y, constant
                             col_map=E._i32(self.x_cols + self.y_cols),
                             basis_idx=E._i32(list(range(self.din)) + [self.D]))
        return self._dev

    def split(self, S):
        """(K, D+1, D+1) second moments of zt -> (yxT, xxT, yyT, n)."""
        xi, yi, D = np.array(self.x_cols), np.array(self.y_cols), self.D
        return (S[:, yi[:, None], xi[None, :]].copy(), S[:, xi[:, None], xi[None, :]].copy(),
                S[:, yi[:, None], yi[None, :]].copy(), S[:, D, D].copy())


class StackedLinearGaussiansWithPrecision:
    _tied = False

    def __init__(self, size, column_dim, row_dim, As=None, lmbdas=None, affine=True, precision=None):
        self.size = size
        self.column_dim = column_dim
        self.row_dim = row_dim
        self.affine = affine
        self.precision = precision
        self.As = None if As is None else np.array(As, dtype=np.float64)
        self.lmbdas = None if lmbdas is None else np.array(lmbdas, dtype=np.float64)
        self.layout = ExpertLayout(column_dim, row_dim, affine)

    def _precision(self):
        return self.precision or E.default_precision()

    @property
    def params(self):
        return self.As, self.lmbdas

    @params.setter
    def params(self, values):
        self.As, self.lmbdas = (np.array(v, dtype=np.float64) for v in values)

    @property
    def dists(self):
        return [LinearGaussianWithPrecision(self.column_dim, self.row_dim, self.As[k], self.lmbdas[k],
                                            affine=self.affine, precision=self.precision) for k in range(self.size)]

    @property
    def input_dim(self):
        return self.column_dim - 1 if self.affine else self.column_dim

    @property
    def output_dim(self):
        return self.row_dim

    @property
    def lmbdas_chol(self):
        return np.transpose(np.linalg.cholesky(self.lmbdas), (0, 2, 1))

    @property
    def lmbdas_chol_inv(self):
        return np.linalg.inv(self.lmbdas_chol)

    @property
    def sigmas(self):
        return np.linalg.inv(self.lmbdas)

    def predict(self, x):
        x = np.asarray(x)
        if self.affine:
            return np.einsum('kdl,...l->k...d', self.As[:, :, :-1], x) + self.As[:, None, :, -1] \
                if x.ndim == 2 else np.einsum('kdl,l->kd', self.As[:, :, :-1], x) + self.As[:, :, -1]
        return np.einsum('kdl,...l->k...d', self.As, x)

    def mean(self, x):
        return self.predict(x)

    def mode(self, x):
        return self.predict(x)

    def rvs(self, x):
        Uinv = self.lmbdas_chol_inv
        mu = self.predict(x)
        return np.array([mu[k] + npr.normal(size=mu[k].shape).dot(Uinv[k].T) for k in range(self.size)])

    @property
    def base(self):
        return np.power(2. * np.pi, -self.output_dim / 2.) * np.ones((self.size,))

    def log_base(self):
        return np.log(self.base)

    # -- statistics ----------------------------------------------------------------
    def statistics(self, x, y, fold=True):
        if not (isinstance(x, np.ndarray) and isinstance(y, np.ndarray)):
            stats = [self.statistics(a, b, fold=fold) for a, b in zip(x, y)]
            return sum(stats[1:], stats[0]) if fold else stats
        good = _clean_rows(x, y)
        x, y = x[good], y[good]
        rep = lambda a: np.array([a for _ in range(self.size)])
        if fold:
            S = soft_stats_quad(np.hstack((x, y)), np.ones((1, len(x))), 'fp64')
            return Stats([rep(a[0]) for a in self.layout.split(S)])
        xt = np.hstack((x, np.ones((len(x), 1)))) if self.affine else x
        return Stats([rep(np.einsum('nd,nl->ndl', y, xt)), rep(np.einsum('nd,nl->ndl', xt, xt)),
                      rep(np.einsum('nd,nl->ndl', y, y)), rep(np.ones((len(y),)))])

    def weighted_statistics(self, x, y, weights):
        if not (isinstance(x, np.ndarray) and isinstance(y, np.ndarray)):
            stats = [self.weighted_statistics(a, b, w) for a, b, w in zip(x, y, weights)]
            return sum(stats[1:], stats[0])
        good = _clean_rows(x, y)
        S = soft_stats_quad(np.hstack((x[good], y[good])), np.asarray(weights)[:, good], self._precision())
        return Stats(self.layout.split(S))

    # -- log-likelihood --------------------------------------------------------------
    def _operands(self, precision, logw=None):
        lay = self.layout
        ops = E.QuadOperands(self.size, lay.D, self.row_dim, precision)
        if logw is not None:
            E.set_log_weights(ops, logw)
        E.operands_lingauss(ops, E.to_dev(self.As), E.to_dev(self.lmbdas), 0, lay.dev()['col_map']).check()
        return ops

    def log_partition(self, x):
        mu = self.predict(x)
        logdet_half = np.sum(np.log(np.diagonal(self.lmbdas_chol, axis1=1, axis2=2)), axis=1)
        return 0.5 * np.einsum('knd,kdl,knl->kn', mu, self.lmbdas, mu) - logdet_half[:, None]

    def log_likelihood(self, x, y):
        if not (isinstance(x, np.ndarray) and isinstance(y, np.ndarray)):
            return [self.log_likelihood(a, b) for a, b in zip(x, y)]
        x = np.atleast_2d(x).reshape((-1, self.input_dim))
        y = np.atleast_2d(y).reshape((-1, self.output_dim))
        precision = self._precision()
        z = np.nan_to_num(np.hstack((x, y)))
        ops = self._operands(precision)
        return E.to_host(E.loglik(E.to_dev(z, E.tdtype(precision)), ops)).astype(np.float64)

    # -- EM ----------------------------------------------------------------------------
    def max_likelihood(self, x, y, weights):
        good = _clean_rows(x, y)
        precision = self._precision()
        lay = self.layout
        feats = E.quad_features(lay.D)
        Z = E.to_dev(np.hstack((x[good], y[good])), E.tdtype(precision))
        R = E.to_dev(np.asarray(weights)[:, good], E.tdtype(precision))
        stat = E.stats_soft(Z, R, feats, precision)
        A, lmbda, info = E.mstep_lingauss(stat, feats.F, lay.dev()['stat_idx'], lay.D + 1,
                                          self.size, self.column_dim, self.row_dim, tied=self._tied)
        try:
            info.check()
        except np.linalg.LinAlgError as e:
            raise AssertionError(str(e))
        self.As, self.lmbdas = E.to_host(A), E.to_host(lmbda)


class TiedLinearGaussiansWithPrecision(StackedLinearGaussiansWithPrecision):
    _tied = True


class LinearGaussianWithPrecision:

    def __init__(self, column_dim, row_dim, A=None, lmbda=None, affine=True, precision=None):
        self.column_dim = column_dim
        self.row_dim = row_dim
        self.A = A
        self.lmbda = lmbda
        self.affine = affine
        self.precision = precision

    def _stack(self):
        return StackedLinearGaussiansWithPrecision(1, self.column_dim, self.row_dim, As=np.asarray(self.A)[None],
                                                   lmbdas=np.asarray(self.lmbda)[None], affine=self.affine,
                                                   precision=self.precision)

    @property
    def params(self):
        return self.A, self.lmbda

    @params.setter
    def params(self, values):
        self.A, self.lmbda = values

    @property
    def nb_params(self):
        return self.column_dim * self.row_dim + self.row_dim * (self.row_dim + 1) / 2

    @property
    def input_dim(self):
        return self.column_dim - 1 if self.affine else self.column_dim

    @property
    def output_dim(self):
        return self.row_dim

    @property
    def lmbda_chol(self):
        return np.linalg.cholesky(self.lmbda).T

    @property
    def lmbda_chol_inv(self):
        return np.linalg.inv(self.lmbda_chol)

    @property
    def sigma(self):
        return np.linalg.inv(self.lmbda)

    def predict(self, x):
        if self.affine:
            return np.einsum('dl,...l->...d', self.A[:, :-1], x) + self.A[:, -1]
        return np.einsum('dl,...l->...d', self.A, x)

    def mean(self, x):
        return self.predict(x)

    def mode(self, x):
        return self.predict(x)

    def rvs(self, x):
        size = self.output_dim if x.ndim == 1 else (x.shape[0], self.output_dim)
        return self.mean(x) + npr.normal(size=size).dot(self.lmbda_chol_inv.T)

    def statistics(self, x, y, fold=True):
        return Stats([s[0] for s in self._stack().statistics(x, y, fold=fold)])

    def weighted_statistics(self, x, y, weights):
        return Stats([s[0] for s in self._stack().weighted_statistics(x, y, np.asarray(weights)[None, :])])

    def log_likelihood(self, x, y):
        if not (isinstance(x, np.ndarray) and isinstance(y, np.ndarray)):
            return [self.log_likelihood(a, b) for a, b in zip(x, y)]
        return self._stack().log_likelihood(x, y)[0]

    def max_likelihood(self, x, y, weights=None):
        w = np.ones((len(x),)) if weights is None else np.asarray(weights)
        st = self._stack()
        st.max_likelihood(x, y, w[None, :])
        self.A, self.lmbda = st.As[0], st.lmbdas[0]


class StackedAffineLinearGaussiansWithPrecision:
    """y = A_k x + c_k + eps with slope and offset kept apart (lingauss.py:576-745): the likelihood object of the
    tied-slope experts of the hierarchical mixtures.  On the device it is the affine expert
    [A_k | c_k] of StackedLinearGaussiansWithPrecision: same operands, same kernels."""

    def __init__(self, size, column_dim, row_dim, As=None, cs=None, lmbdas=None, precision=None):
        self.size, self.column_dim, self.row_dim, self.precision = size, column_dim, row_dim, precision
        self.As, self.cs, self.lmbdas = As, cs, lmbdas
        self.layout = ExpertLayout(column_dim + 1, row_dim, affine=True)
        self.affine = True

    @property
    def params(self):
        return self.As, self.cs, self.lmbdas

    @params.setter
    def params(self, values):
        self.As, self.cs, self.lmbdas = values

    @property
    def input_dim(self):
        return self.column_dim

    @property
    def output_dim(self):
        return self.row_dim

    def _affine_As(self):
        """(K, o, c + 1): [A_k | c_k], what the operand kernels take."""
        return np.concatenate((np.asarray(self.As, dtype=np.float64), np.asarray(self.cs, dtype=np.float64)[:, :, None]), axis=2)

    def _combined(self):
        return StackedLinearGaussiansWithPrecision(self.size, self.column_dim + 1, self.row_dim, As=self._affine_As(),
                                                   lmbdas=self.lmbdas, affine=True, precision=self.precision)

    def predict(self, x):
        return self._combined().predict(x)

    def mean(self, x):
        return self.predict(x)

    def mode(self, x):
        return self.predict(x)

    def rvs(self, x):
        return self._combined().rvs(x)

    def log_likelihood(self, x, y):
        return self._combined().log_likelihood(x, y)

    def weighted_statistics(self, x, y, weights):
        return self._combined().weighted_statistics(x, y, weights)


class AffineLinearGaussianWithPrecision:
    """one expert y = A x + c + eps with slope and offset kept apart (lingauss.py:401-573) = a stack of one."""

    def __init__(self, column_dim, row_dim, A=None, c=None, lmbda=None, precision=None):
        self.column_dim, self.row_dim, self.precision = column_dim, row_dim, precision
        self.A, self.c, self.lmbda = A, c, lmbda

    @property
    def params(self):
        return self.A, self.c, self.lmbda

    @params.setter
    def params(self, values):
        self.A, self.c, self.lmbda = values

    def _stack(self):
        return StackedAffineLinearGaussiansWithPrecision(1, self.column_dim, self.row_dim, np.asarray(self.A)[None],
                                                         np.asarray(self.c)[None], np.asarray(self.lmbda)[None], precision=self.precision)

    def predict(self, x):
        return self._stack().predict(x)[0]

    def mean(self, x):
        return self.predict(x)

    def rvs(self, x):
        return self._stack().rvs(x)[0]

    def log_likelihood(self, x, y):
        return self._stack().log_likelihood(x, y)[0]
