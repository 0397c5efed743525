"""Gaussian likelihoods, full and diagonal precision, single / stacked / tied
(API of mimo/distributions/gaussian.py; the per-point and per-component arithmetic runs
in libmimo_b200.so).

Every heavy method uploads its inputs, calls the C-ABI kernels and downloads the result:
  log_likelihood       -> mimo_operands_gauss[_diag] + mimo_loglik_quad / _diag
  weighted_statistics  -> mimo_stats_soft           (packed FP64 statistics, unpacked here)
  max_likelihood       -> mimo_stats_soft + mimo_mstep_gauss[_diag]
The sweep drivers in mimo_b200.mixtures do not go through these object methods; they keep
data and operands resident and call the fused sweep (see mixtures/gmm.py).
"""
import numpy as np
import numpy.random as npr

from .. import _engine as E
from ..utils.abstraction import Statistics as Stats

LOG_2PI = np.log(2. * np.pi)


def _clean_rows(*arrays):
    """rows without NaN in any of the arrays (the reference drops them from statistics:
    gaussian.py:493-494, lingauss.py:308-310)."""
    good = np.ones(len(arrays[0]), dtype=bool)
    for a in arrays:
        good &= ~np.isnan(a).any(axis=1)
    return good


def unpack_quad(stat, Dp):
    """packed lower-triangular (K, F) -> symmetric (K, Dp, Dp)."""
    K = stat.shape[0]
    il = np.tril_indices(Dp)
    S = np.zeros((K, Dp, Dp))
    S[:, il[0], il[1]] = stat
    S[:, il[1], il[0]] = stat
    return S


def soft_stats_quad(z, weights, precision):
    """sum_n r_kn [z;1][z;1]^T  as (K, D+1, D+1) host array, computed on the GPU."""
    D = z.shape[1]
    feats = E.quad_features(D)
    Z = E.to_dev(z, E.tdtype(precision))
    R = E.to_dev(weights, E.tdtype(precision))
    return unpack_quad(E.to_host(E.stats_soft(Z, R, feats, precision)), D + 1)


class _StackedBase:
    """shared plumbing of the stacked likelihoods."""

    def _precision(self):
        return self.precision or E.default_precision()

    @property
    def nb_params(self):
        raise NotImplementedError

    @property
    def nat_param(self):
        return self.std_to_nat(self.params)

    @nat_param.setter
    def nat_param(self, natparam):
        self.params = self.nat_to_std(natparam)

    def mean(self):
        return self.mus

    def mode(self):
        return self.mus

    @property
    def base(self):
        return np.power(2. * np.pi, -self.dim / 2.) * np.ones((self.size,))

    def log_base(self):
        return np.log(self.base)


class StackedGaussiansWithPrecision(_StackedBase):
    _tied = False

    def __init__(self, size, dim, mus=None, lmbdas=None, precision=None):
        self.size = size
        self.dim = dim
        self.precision = precision
        self.mus = None if mus is None else np.array(mus, dtype=np.float64)
        self.lmbdas = None if lmbdas is None else np.array(lmbdas, dtype=np.float64)

    @property
    def params(self):
        return self.mus, self.lmbdas

    @params.setter
    def params(self, values):
        self.mus, self.lmbdas = (np.array(v, dtype=np.float64) for v in values)

    @property
    def dists(self):
        return [GaussianWithPrecision(self.dim, self.mus[k], self.lmbdas[k], precision=self.precision)
                for k in range(self.size)]

    def std_to_nat(self, params):
        mus, lmbdas = params
        return Stats([np.einsum('kdl,kl->kd', lmbdas, mus), -0.5 * lmbdas])

    def nat_to_std(self, natparam):
        lmbdas = -2. * natparam[1]
        return np.linalg.solve(lmbdas, natparam[0][..., None])[..., 0], lmbdas

    @property
    def lmbdas_chol(self):
        return np.transpose(np.linalg.cholesky(self.lmbdas), (0, 2, 1))     # upper factors

    @property
    def lmbdas_chol_inv(self):
        return np.linalg.inv(self.lmbdas_chol)

    @property
    def sigmas(self):
        return np.linalg.inv(self.lmbdas)

    def rvs(self, sizes):
        Uinv = self.lmbdas_chol_inv
        return np.vstack([self.mus[k] + npr.normal(size=(int(n), self.dim)).dot(Uinv[k].T)
                          for k, n in enumerate(sizes)])

    # -- statistics ------------------------------------------------------------------
    def statistics(self, data, fold=True):
        if not isinstance(data, np.ndarray):
            stats = [self.statistics(d, fold=fold) for d in data]
            return sum(stats[1:], stats[0]) if fold else stats
        data = data[_clean_rows(data)]
        if fold:
            S = soft_stats_quad(data, np.ones((1, len(data))), 'fp64')[0]
            d = self.dim
            rep = lambda a: np.array([a for _ in range(self.size)])
            return Stats([rep(S[d, :d]), rep(S[d, d]), rep(S[:d, :d]), rep(S[d, d])])
        # per-point statistics replicated K times (K*N*d*d values: small N only; the
        # sweep never needs this -- see expected_log_likelihood in bayesian.py)
        xxT = np.einsum('nd,nl->ndl', data, data)
        n = np.ones((data.shape[0],))
        rep = lambda a: np.array([a for _ in range(self.size)])
        return Stats([rep(data), rep(n), rep(xxT), rep(n)])

    def weighted_statistics(self, data, weights):
        if not isinstance(data, np.ndarray):
            stats = [self.weighted_statistics(d, w) for d, w in zip(data, weights)]
            return sum(stats[1:], stats[0])
        good = _clean_rows(data)
        S = soft_stats_quad(data[good], np.asarray(weights)[:, good], self._precision())
        d = self.dim
        return Stats([S[:, d, :d].copy(), S[:, d, d].copy(), S[:, :d, :d].copy(), S[:, d, d].copy()])

    # -- log-likelihood ---------------------------------------------------------------
    def _operands(self, precision, logw=None):
        ops = E.QuadOperands(self.size, self.dim, self.dim, precision)
        if logw is not None:
            E.set_log_weights(ops, logw)
        E.operands_gauss(ops, E.to_dev(self.mus), E.to_dev(self.lmbdas)).check()
        return ops

    def log_partition(self):
        ops = self._operands('fp64')
        return -(E.to_host(ops.cst) + 0.5 * self.dim * LOG_2PI) \
            + 0.5 * np.einsum('kd,kdl,kl->k', self.mus, self.lmbdas, self.mus)

    def log_likelihood(self, x):
        if not isinstance(x, np.ndarray):
            return [self.log_likelihood(xi) for xi in x]
        x = np.atleast_2d(x).reshape((-1, self.dim))
        bads = np.isnan(x).any(axis=1)
        precision = self._precision()
        ops = self._operands(precision)
        out = E.to_host(E.loglik(E.to_dev(np.nan_to_num(x), E.tdtype(precision)), ops)).astype(np.float64)
        if bads.any():   # the reference zeroes the data term of NaN rows (gaussian.py:518)
            zero = E.to_host(E.loglik(E.zeros((1, self.dim), E.tdtype(precision)), ops)).astype(np.float64)
            out[:, bads] = zero
        return out

    # -- EM ---------------------------------------------------------------------------
    def max_likelihood(self, data, weights):
        good = _clean_rows(data)
        precision = self._precision()
        feats = E.quad_features(self.dim)
        Z = E.to_dev(data[good], E.tdtype(precision))
        R = E.to_dev(np.asarray(weights)[:, good], E.tdtype(precision))
        stat = E.stats_soft(Z, R, feats, precision)
        mu, lmbda, info = E.mstep_gauss(stat, feats.F, E.identity_map(self.dim, self.dim), self.dim + 1,
                                        self.size, self.dim, tied=self._tied)
        try:
            info.check()
        except np.linalg.LinAlgError as e:     # the reference asserts eigvalsh(sigma) > 0
            raise AssertionError(str(e))
        self.mus, self.lmbdas = E.to_host(mu), E.to_host(lmbda)


class TiedGaussiansWithPrecision(StackedGaussiansWithPrecision):
    _tied = True


class GaussianWithPrecision:
    """single component = a stack of one."""

    def __init__(self, dim, mu=None, lmbda=None, precision=None):
        self.dim = dim
        self.mu = mu
        self.lmbda = lmbda
        self.precision = precision

    def _stack(self):
        return StackedGaussiansWithPrecision(1, self.dim, mus=np.asarray(self.mu)[None], lmbdas=np.asarray(self.lmbda)[None],
                                             precision=self.precision)

    @property
    def params(self):
        return self.mu, self.lmbda

    @params.setter
    def params(self, values):
        self.mu, self.lmbda = values

    @property
    def nb_params(self):
        return self.dim + self.dim * (self.dim + 1) / 2

    @staticmethod
    def std_to_nat(params):
        return Stats([params[1] @ params[0], -0.5 * params[1]])

    @staticmethod
    def nat_to_std(natparam):
        lmbda = -2. * natparam[1]
        return np.linalg.solve(lmbda, natparam[0]), lmbda

    @property
    def nat_param(self):
        return self.std_to_nat(self.params)

    @nat_param.setter
    def nat_param(self, natparam):
        self.params = self.nat_to_std(natparam)

    @property
    def lmbda_chol(self):
        return np.linalg.cholesky(self.lmbda).T

    @property
    def lmbda_chol_inv(self):
        return np.linalg.inv(self.lmbda_chol)

    @property
    def sigma(self):
        return np.linalg.inv(self.lmbda)

    def mean(self):
        return self.mu

    def mode(self):
        return self.mu

    def rvs(self, size=1):
        size = self.dim if size == 1 else (size, self.dim)
        return self.mu + npr.normal(size=size).dot(self.lmbda_chol_inv.T)

    @property
    def base(self):
        return np.power(2. * np.pi, -self.dim / 2.)

    def log_base(self):
        return np.log(self.base)

    def log_partition(self):
        return self._stack().log_partition()[0]

    def log_likelihood(self, x):
        if not isinstance(x, np.ndarray):
            return [self.log_likelihood(xi) for xi in x]
        return self._stack().log_likelihood(x)[0]

    def statistics(self, data, fold=True):
        if not isinstance(data, np.ndarray):
            stats = [self.statistics(d, fold=fold) for d in data]
            return sum(stats[1:], stats[0]) if fold else stats
        return Stats([s[0] for s in self._stack().statistics(data, fold=fold)])

    def weighted_statistics(self, data, weights):
        if not isinstance(data, np.ndarray):
            stats = [self.weighted_statistics(d, w) for d, w in zip(data, weights)]
            return sum(stats[1:], stats[0])
        return Stats([s[0] for s in self._stack().weighted_statistics(data, np.asarray(weights)[None, :])])

    def expected_statistics(self):
        return self.mu, np.outer(self.mu, self.mu) + self.sigma

    def entropy(self):
        """d/2 (1 + log 2 pi) - 1/2 log det Lambda   (gaussian.py:96-99 in closed form)."""
        return 0.5 * self.dim * (1. + np.log(2. * np.pi)) - 0.5 * np.linalg.slogdet(self.lmbda)[1]

    def cross_entropy(self, dist):
        """E_self[-log dist]   (gaussian.py:101-104 in closed form)."""
        diff = self.mu - dist.mu
        return 0.5 * self.dim * np.log(2. * np.pi) - 0.5 * np.linalg.slogdet(dist.lmbda)[1] \
            + 0.5 * np.trace(dist.lmbda @ self.sigma) + 0.5 * diff @ dist.lmbda @ diff

    def max_likelihood(self, data, weights=None):
        w = np.ones((len(data),)) if weights is None else np.asarray(weights)
        st = self._stack()
        st.max_likelihood(data, w[None, :])
        self.mu, self.lmbda = st.mus[0], st.lmbdas[0]


# ---------------------------------------------------------------------------------------
# diagonal precision
# ---------------------------------------------------------------------------------------
def soft_stats_diag(z, weights, precision):
    """(K, 2D+1) host array [sum r z | sum r z^2 | sum r]."""
    feats = E.diag_features(z.shape[1])
    Z = E.to_dev(z, E.tdtype(precision))
    R = E.to_dev(weights, E.tdtype(precision))
    return E.to_host(E.stats_soft(Z, R, feats, precision))


class StackedGaussiansWithDiagonalPrecision(_StackedBase):
    _tied = False

    def __init__(self, size, dim, mus=None, lmbdas_diags=None, precision=None):
        self.size = size
        self.dim = dim
        self.precision = precision
        self.mus = None if mus is None else np.array(mus, dtype=np.float64)
        self.lmbdas_diags = None if lmbdas_diags is None else np.array(lmbdas_diags, dtype=np.float64)

    @property
    def params(self):
        return self.mus, self.lmbdas_diags

    @params.setter
    def params(self, values):
        self.mus, self.lmbdas_diags = (np.array(v, dtype=np.float64) for v in values)

    @property
    def dists(self):
        return [GaussianWithDiagonalPrecision(self.dim, self.mus[k], self.lmbdas_diags[k], precision=self.precision)
                for k in range(self.size)]

    def std_to_nat(self, params):
        return Stats([params[1] * params[0], -0.5 * params[1]])

    def nat_to_std(self, natparam):
        return -0.5 * natparam[0] / natparam[1], -2. * natparam[1]

    @property
    def lmbdas(self):
        return np.array([np.diag(l) for l in self.lmbdas_diags])

    @property
    def lmbdas_chol(self):
        return np.array([np.diag(np.sqrt(l)) for l in self.lmbdas_diags])

    @property
    def lmbdas_chol_inv(self):
        return np.array([np.diag(1. / np.sqrt(l)) for l in self.lmbdas_diags])

    @property
    def sigmas_diags(self):
        return 1. / self.lmbdas_diags

    @property
    def sigmas(self):
        return np.array([np.diag(1. / l) for l in self.lmbdas_diags])

    def rvs(self, sizes):
        return np.vstack([self.mus[k] + npr.normal(size=(int(n), self.dim)) / np.sqrt(self.lmbdas_diags[k])
                          for k, n in enumerate(sizes)])

    def statistics(self, data, fold=True):
        if not isinstance(data, np.ndarray):
            stats = [self.statistics(d, fold=fold) for d in data]
            return sum(stats[1:], stats[0]) if fold else stats
        data = data[_clean_rows(data)]
        rep = lambda a: np.array([a for _ in range(self.size)])
        d = self.dim
        if fold:
            S = soft_stats_diag(data, np.ones((1, len(data))), 'fp64')[0]
            nd = np.broadcast_to(S[2 * d], (d,))
            return Stats([rep(S[:d]), rep(nd), rep(nd), rep(S[d:2 * d])])
        nd = np.ones(data.shape)
        return Stats([rep(data), rep(nd), rep(nd), rep(data * data)])

    def weighted_statistics(self, data, weights):
        if not isinstance(data, np.ndarray):
            stats = [self.weighted_statistics(d, w) for d, w in zip(data, weights)]
            return sum(stats[1:], stats[0])
        good = _clean_rows(data)
        S = soft_stats_diag(data[good], np.asarray(weights)[:, good], self._precision())
        d = self.dim
        ndk = np.broadcast_to(S[:, 2 * d:2 * d + 1], (self.size, d)).copy()
        return Stats([S[:, :d].copy(), ndk, ndk.copy(), S[:, d:2 * d].copy()])

    def _operands(self, precision, logw=None):
        ops = E.DiagOperands(self.size, self.dim, precision)
        if logw is not None:
            E.set_log_weights(ops, logw)
        E.operands_gauss_diag(ops, E.to_dev(self.mus), E.to_dev(self.lmbdas_diags))
        return ops

    def log_partition(self):
        return 0.5 * np.sum(self.mus * self.lmbdas_diags * self.mus, axis=1) \
            - 0.5 * np.sum(np.log(self.lmbdas_diags), axis=1)

    def log_likelihood(self, x):
        if not isinstance(x, np.ndarray):
            return [self.log_likelihood(xi) for xi in x]
        x = np.atleast_2d(x).reshape((-1, self.dim))
        bads = np.isnan(x).any(axis=1)
        precision = self._precision()
        ops = self._operands(precision)
        out = E.to_host(E.loglik(E.to_dev(np.nan_to_num(x), E.tdtype(precision)), ops)).astype(np.float64)
        if bads.any():
            zero = E.to_host(E.loglik(E.zeros((1, self.dim), E.tdtype(precision)), ops)).astype(np.float64)
            out[:, bads] = zero
        return out

    def max_likelihood(self, data, weights):
        good = _clean_rows(data)
        precision = self._precision()
        feats = E.diag_features(self.dim)
        Z = E.to_dev(data[good], E.tdtype(precision))
        R = E.to_dev(np.asarray(weights)[:, good], E.tdtype(precision))
        stat = E.stats_soft(Z, R, feats, precision)
        mu, lam = E.mstep_gauss_diag(stat, feats.F, self.size, self.dim, tied=self._tied)
        self.mus, self.lmbdas_diags = E.to_host(mu), E.to_host(lam)


class TiedGaussiansWithDiagonalPrecision(StackedGaussiansWithDiagonalPrecision):
    _tied = True


class GaussianWithDiagonalPrecision:

    def __init__(self, dim, mu=None, lmbda_diag=None, precision=None):
        self.dim = dim
        self.mu = mu
        self.lmbda_diag = lmbda_diag
        self.precision = precision

    def _stack(self):
        return StackedGaussiansWithDiagonalPrecision(1, self.dim, mus=np.asarray(self.mu)[None],
                                                     lmbdas_diags=np.asarray(self.lmbda_diag)[None],
                                                     precision=self.precision)

    @property
    def params(self):
        return self.mu, self.lmbda_diag

    @params.setter
    def params(self, values):
        self.mu, self.lmbda_diag = values

    @property
    def nb_params(self):
        return self.dim + self.dim

    @property
    def lmbda(self):
        return np.diag(self.lmbda_diag)

    @property
    def sigma_diag(self):
        return 1. / self.lmbda_diag

    @property
    def sigma(self):
        return np.diag(self.sigma_diag)

    def mean(self):
        return self.mu

    def mode(self):
        return self.mu

    def rvs(self, size=1):
        size = self.dim if size == 1 else (size, self.dim)
        return self.mu + npr.normal(size=size) / np.sqrt(self.lmbda_diag)

    def log_partition(self):
        return self._stack().log_partition()[0]

    def log_likelihood(self, x):
        if not isinstance(x, np.ndarray):
            return [self.log_likelihood(xi) for xi in x]
        return self._stack().log_likelihood(x)[0]

    def statistics(self, data, fold=True):
        if not isinstance(data, np.ndarray):
            stats = [self.statistics(d, fold=fold) for d in data]
            return sum(stats[1:], stats[0]) if fold else stats
        return Stats([s[0] for s in self._stack().statistics(data, fold=fold)])

    def weighted_statistics(self, data, weights):
        if not isinstance(data, np.ndarray):
            stats = [self.weighted_statistics(d, w) for d, w in zip(data, weights)]
            return sum(stats[1:], stats[0])
        return Stats([s[0] for s in self._stack().weighted_statistics(data, np.asarray(weights)[None, :])])

    def expected_statistics(self):
        return self.mu, self.mu ** 2 + self.sigma_diag

    def max_likelihood(self, data, weights=None):
        w = np.ones((len(data),)) if weights is None else np.asarray(weights)
        st = self._stack()
        st.max_likelihood(data, w[None, :])
        self.mu, self.lmbda_diag = st.mus[0], st.lmbdas_diags[0]


# ---------------------------------------------------------------------------------------
# scaled precision: the prior over a Gaussian mean that shares the likelihood's precision up to a factor kappa
# (hierarchical mixtures).  Host-side parameter objects of a few K d^2 numbers: the per-point work of the
# models that use them runs in the sweep kernels (distributions/bayesian.py: TiedGaussiansWithHierarchicalNormalWisharts).
# ---------------------------------------------------------------------------------------
class GaussianWithScaledPrecision:
    """N(mu, (kappa lmbda)^-1)   (gaussian.py:890-1035).  The reference caches the Cholesky factor of kappa * lmbda
    and resets the cache only when lmbda is assigned, not kappa (:947-961): `omega_chol` keeps that behaviour
    (quirk q11) because the lower bound of the hierarchical mixture depends on it; `omega` is always current."""

    def __init__(self, dim, kappa, mu=None, lmbda=None):
        self.dim, self.mu, self.kappa = dim, mu, kappa
        self._lmbda = lmbda
        self._omega_chol = None

    @property
    def params(self):
        return self.mu, self.kappa

    @params.setter
    def params(self, values):
        self.mu, self.kappa = values

    @property
    def nb_params(self):
        raise NotImplementedError

    @staticmethod
    def std_to_nat(params):
        return Stats([params[1] * params[0], params[1]])

    @staticmethod
    def nat_to_std(natparam):
        return natparam[0] / natparam[1], natparam[1]

    @property
    def nat_param(self):
        return self.std_to_nat(self.params)

    @nat_param.setter
    def nat_param(self, natparam):
        self.params = self.nat_to_std(natparam)

    @property
    def lmbda(self):
        return self._lmbda

    @lmbda.setter
    def lmbda(self, value):
        self._lmbda = value
        self._omega_chol = None

    @property
    def omega(self):
        return self.kappa * self.lmbda

    @property
    def omega_chol(self):
        if self._omega_chol is None:
            self._omega_chol = np.linalg.cholesky(self.omega).T
        return self._omega_chol

    @property
    def omega_chol_inv(self):
        return np.linalg.inv(self.omega_chol)

    @property
    def sigma(self):
        ci = self.omega_chol_inv
        return ci @ ci.T

    def mean(self):
        return self.mu

    def mode(self):
        return self.mu

    @property
    def base(self):
        return np.power(2. * np.pi, -self.dim / 2.)

    def log_base(self):
        return np.log(self.base)

    def rvs(self, size=1):
        size = self.dim if size == 1 else (size, self.dim)
        return self.mu + npr.normal(size=size).dot(self.omega_chol_inv.T)

    def statistics(self, data, fold=True):
        if not isinstance(data, np.ndarray):
            stats = [self.statistics(d, fold=fold) for d in data]
            return sum(stats[1:], stats[0]) if fold else stats
        data = data[_clean_rows(data)]
        return Stats([data.sum(0), data.shape[0]]) if fold else Stats([data, np.ones((data.shape[0],))])

    def weighted_statistics(self, data, weights):
        if not isinstance(data, np.ndarray):
            stats = [self.weighted_statistics(d, w) for d, w in zip(data, weights)]
            return sum(stats[1:], stats[0])
        good = _clean_rows(data)
        w = np.asarray(weights)[good]
        return Stats([w @ data[good], np.sum(w)])

    def log_partition(self):
        return 0.5 * self.mu @ self.lmbda @ self.mu - np.sum(np.log(np.diag(self.omega_chol)))

    def log_likelihood(self, x):
        if not isinstance(x, np.ndarray):
            return [self.log_likelihood(xi) for xi in x]
        bads = np.isnan(np.atleast_2d(x)).any(axis=1)
        x = np.nan_to_num(x).reshape((-1, self.dim))
        om = self.omega
        out = x @ (om @ self.mu) - 0.5 * np.einsum('nd,nd->n', x @ om, x)
        out[bads] = 0.
        return out - self.log_partition() + self.log_base()

    def entropy(self):
        return 0.5 * self.dim * np.log(2. * np.pi * np.e) - np.sum(np.log(np.diag(self.omega_chol)))

    def max_likelihood(self, data, weights=None):
        raise NotImplementedError


class TiedGaussiansWithScaledPrecision:
    """K of the above (gaussian.py:1038-1202); `dists` are live per-component objects, as in the reference."""

    def __init__(self, size, dim, kappas, mus=None, lmbdas=None):
        self.size, self.dim = size, dim
        pick = lambda v, k: None if v is None else v[k]                      # noqa: E731
        self.dists = [GaussianWithScaledPrecision(dim, pick(kappas, k), pick(mus, k), pick(lmbdas, k)) for k in range(size)]

    def _get(self, name):
        return np.array([getattr(dist, name) for dist in self.dists])

    def _set(self, name, value):
        for k, dist in enumerate(self.dists):
            setattr(dist, name, value[k, ...])

    mus = property(lambda self: self._get('mu'), lambda self, v: self._set('mu', v))
    kappas = property(lambda self: self._get('kappa'), lambda self, v: self._set('kappa', v))
    lmbdas = property(lambda self: self._get('lmbda'), lambda self, v: self._set('lmbda', v))
    omegas = property(lambda self: self._get('omega'))
    omegas_chol = property(lambda self: self._get('omega_chol'))
    omegas_chol_ins = property(lambda self: self._get('omega_chol_inv'))
    sigmas = property(lambda self: self._get('sigma'))
    base = property(lambda self: self._get('base'))

    @property
    def params(self):
        return self.mus, self.kappas

    @params.setter
    def params(self, values):
        self.mus, self.kappas = values

    @property
    def nb_params(self):
        raise NotImplementedError

    def std_to_nat(self, params):
        mus, kappas = params
        return Stats([np.asarray(kappas)[:, None] * np.asarray(mus), np.asarray(kappas)])

    def nat_to_std(self, natparam):
        return natparam[0] / natparam[1][:, None], natparam[1]

    @property
    def nat_param(self):
        return self.std_to_nat(self.params)

    @nat_param.setter
    def nat_param(self, natparam):
        self.params = self.nat_to_std(natparam)

    def rvs(self, sizes):
        return np.vstack([dist.rvs(size) for dist, size in zip(self.dists, sizes)])

    def mean(self):
        return self.mus

    def mode(self):
        return self.mus

    def log_base(self):
        return np.log(self.base)

    def statistics(self, data, fold=True):
        if not isinstance(data, np.ndarray):
            stats = [self.statistics(d, fold=fold) for d in data]
            return sum(stats[1:], stats[0]) if fold else stats
        x, n = self.dists[0].statistics(data, fold=fold)
        return Stats([np.array(self.size * [x]), np.array(self.size * [n])])

    def weighted_statistics(self, data, weights):
        if not isinstance(data, np.ndarray):
            stats = [self.weighted_statistics(d, w) for d, w in zip(data, weights)]
            return sum(stats[1:], stats[0])
        good = _clean_rows(data)
        w = np.asarray(weights)[:, good]
        return Stats([w @ data[good], np.sum(w, axis=1)])

    def log_partition(self):
        return np.array([dist.log_partition() for dist in self.dists])

    def log_likelihood(self, x):
        if not isinstance(x, np.ndarray):
            return [self.log_likelihood(xi) for xi in x]
        return np.stack([dist.log_likelihood(x) for dist in self.dists])

    def max_likelihood(self, data, weights):
        raise NotImplementedError
