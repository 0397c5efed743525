"""Conjugate priors / posteriors in natural-parameter form: Normal-Wishart, Normal-Gamma,
Matrix-Normal-Wishart; single, stacked (K independent) and tied (shared scale) variants
(API of mimo/distributions/composite.py).

State is kept as stacked NumPy arrays.  Inside a sweep the conjugate update, the expected
natural parameters, sampling and the lower-bound terms run in the batched posterior kernels
(mimo_nw_posterior / mimo_ng_posterior / mimo_mnw_posterior); `rvs()` here goes through the
same kernels (zero statistics + host-drawn variates in the reference's RNG order), while
the small closed-form helpers (`expected_statistics`, `log_partition`, `entropy`, ...) serve
the object API in FP64 on the host.
"""
import numpy as np
import numpy.random as npr
from scipy.special import digamma, gammaln, multigammaln

from .. import _engine as E
from ..utils.abstraction import Statistics as Stats

LOG_2PI = np.log(2. * np.pi)


def _logdet_spd(a):
    return 2. * np.sum(np.log(np.diagonal(np.linalg.cholesky(a), axis1=-2, axis2=-1)), axis=-1)


def _wishart_log_partition(psis, nus):
    d = psis.shape[-1]
    return 0.5 * nus * d * np.log(2.) + np.array([multigammaln(nu / 2., d) for nu in np.atleast_1d(nus)]) \
        + 0.5 * nus * _logdet_spd(psis)


def _wishart_elogdet(psis, nus):
    d = psis.shape[-1]
    return np.sum(digamma((np.atleast_1d(nus)[:, None] - np.arange(d)[None, :]) / 2.), axis=1) \
        + d * np.log(2.) + _logdet_spd(psis)


def draw_wishart_variates(nus, d, extra):
    """Host draws in the reference's stream order, per component:
    normal(d(d-1)/2), d x chisquare(nu - i), normal(extra)
    (wishart.py:72-80 then gaussian.py:311-313 / matrix.py:123-125)."""
    K = len(nus)
    nt = d * (d - 1) // 2
    var = np.empty((K, nt + d + extra))
    for k in range(K):
        var[k, :nt] = npr.normal(size=nt)
        for i in range(d):
            var[k, nt + i] = npr.chisquare(nus[k] - i, size=1)[0]
        var[k, nt + d:] = npr.normal(size=extra)
    return var


class _StackedPrior:
    """shared accessors: params tuple <-> natural parameters."""
    _tied = False

    @property
    def nb_params(self):
        raise NotImplementedError

    @property
    def nat_param(self):
        return self.std_to_nat(self.params)

    @nat_param.setter
    def nat_param(self, natparam):
        self.params = self.nat_to_std(natparam)

    @property
    def base(self):
        raise NotImplementedError

    def entropy(self):
        return self.log_partition() - self._dot(self.nat_param, self.expected_statistics())

    def cross_entropy(self, other):
        return other.log_partition() - self._dot(other.nat_param, self.expected_statistics())


# ---------------------------------------------------------------------------------------
class StackedNormalWisharts(_StackedPrior):

    def __init__(self, size, dim, mus=None, kappas=None, psis=None, nus=None):
        self.size, self.dim = size, dim
        self.mus, self.kappas, self.psis, self.nus = (None if v is None else np.array(v, dtype=np.float64)
                                                     for v in (mus, kappas, psis, nus))

    @property
    def params(self):
        return self.mus, self.kappas, self.psis, self.nus

    @params.setter
    def params(self, values):
        self.mus, self.kappas, self.psis, self.nus = (np.array(v, dtype=np.float64) for v in values)

    @property
    def dists(self):
        return [NormalWishart(self.dim, self.mus[k], self.kappas[k], self.psis[k], self.nus[k]) for k in range(self.size)]

    def std_to_nat(self, params):
        mus, kappas, psis, nus = params
        return Stats([kappas[:, None] * mus, np.array(kappas),
                      np.linalg.inv(psis) + kappas[:, None, None] * np.einsum('kd,kl->kdl', mus, mus),
                      nus - self.dim])

    def nat_to_std(self, natparam):
        a, b, c, e = natparam
        mus = a / b[:, None]
        inner = c - b[:, None, None] * np.einsum('kd,kl->kdl', mus, mus)
        if self._tied:
            psi = np.linalg.inv(np.mean(inner, axis=0))
            nu = np.mean(e + self.dim)
            return mus, np.array(b), np.array(self.size * [psi]), np.array(self.size * [nu])
        return mus, np.array(b), np.linalg.inv(inner), e + self.dim

    def mean(self):
        return self.mus, self.nus[:, None, None] * self.psis

    def mode(self):
        return self.mus, (self.nus - self.dim)[:, None, None] * self.psis

    def rvs(self):
        """(mus, lmbdas) ~ NW; Bartlett + Cholesky algebra on the device, variates from the
        global numpy.random stream in the reference's order."""
        var = draw_wishart_variates(self.nus, self.dim, self.dim)
        feats = E.quad_features(self.dim)
        idx = E.identity_map(self.dim, self.dim)
        out = E.nw_posterior([E.to_dev(p) for p in self.params], E.zeros((self.size, feats.F)), feats.F, idx,
                             self.dim + 1, mode=1, variates=var, want_lik=True, want_vlb=False)
        out['info'].check()
        return E.to_host(out['lik_mu']), E.to_host(out['lik_lmbda'])

    def log_base(self):
        return -0.5 * self.dim * LOG_2PI * np.ones((self.size,))

    def log_partition(self):
        return -0.5 * self.dim * np.log(self.kappas) + _wishart_log_partition(self.psis, self.nus)

    def log_likelihood(self, x):
        """sum_k log NW(mu_k, lmbda_k)  (composite.py:100-104, 244-246)."""
        mus, lmbdas = x
        d = self.dim
        diff = mus - self.mus
        ll_mu = -0.5 * self.kappas * np.einsum('kd,kdl,kl->k', diff, lmbdas, diff) \
            + 0.5 * (d * np.log(self.kappas) + _logdet_spd(lmbdas)) - 0.5 * d * LOG_2PI
        ll_w = 0.5 * (self.nus - d - 1.) * _logdet_spd(lmbdas) \
            - 0.5 * np.trace(np.linalg.solve(self.psis, lmbdas), axis1=1, axis2=2) \
            - _wishart_log_partition(self.psis, self.nus)
        return np.sum(ll_mu + ll_w)

    def expected_statistics(self):
        E_lm = self.nus[:, None] * np.einsum('kdl,kl->kd', self.psis, self.mus)
        E_mlm = -0.5 * (self.dim / self.kappas + np.einsum('kd,kd->k', self.mus, E_lm))
        return E_lm, E_mlm, -0.5 * self.nus[:, None, None] * self.psis, 0.5 * _wishart_elogdet(self.psis, self.nus)

    @staticmethod
    def _dot(nat, stats):
        return np.einsum('kd,kd->k', nat[0], stats[0]) + nat[1] * stats[1] \
            + np.einsum('kdl,kdl->k', nat[2], stats[2]) + nat[3] * stats[3]


class TiedNormalWisharts(StackedNormalWisharts):
    _tied = True


class NormalWishart:
    """single component view (composite.py:19-134)."""

    def __init__(self, dim, mu=None, kappa=None, psi=None, nu=None):
        self.dim = dim
        self.mu, self.kappa, self.psi, self.nu = mu, kappa, psi, nu

    def _stack(self):
        return StackedNormalWisharts(1, self.dim, np.asarray(self.mu)[None], np.atleast_1d(self.kappa),
                                     np.asarray(self.psi)[None], np.atleast_1d(self.nu))

    @property
    def gaussian(self):
        """the location factor (composite.py:26); a view: assign through `params`, not through the view."""
        from .gaussian import GaussianWithPrecision
        return GaussianWithPrecision(self.dim, mu=self.mu)

    @property
    def wishart(self):
        """the precision factor (composite.py:27); a view."""
        from .wishart import Wishart
        return Wishart(self.dim, psi=self.psi, nu=self.nu)

    @property
    def params(self):
        return self.mu, self.kappa, self.psi, self.nu

    @params.setter
    def params(self, values):
        self.mu, self.kappa, self.psi, self.nu = values

    @property
    def nat_param(self):
        return self.std_to_nat(self.params)

    @nat_param.setter
    def nat_param(self, natparam):
        self.params = self.nat_to_std(natparam)

    def std_to_nat(self, params):
        mu, kappa, psi, nu = params
        return Stats([kappa * mu, kappa, np.linalg.inv(psi) + kappa * np.outer(mu, mu), nu - self.dim])

    def nat_to_std(self, natparam):
        mu = natparam[0] / natparam[1]
        return mu, natparam[1], np.linalg.inv(natparam[2] - natparam[1] * np.outer(mu, mu)), natparam[3] + self.dim

    def mean(self):
        return self.mu, self.nu * self.psi

    def mode(self):
        return self.mu, (self.nu - self.dim) * self.psi

    def rvs(self):
        mus, lmbdas = self._stack().rvs()
        return mus[0], lmbdas[0]

    @property
    def base(self):
        """(2 pi)^(-d/2): Gaussian base measure times the Wishart's 1 (composite.py:88-93)."""
        return np.power(2. * np.pi, -0.5 * self.dim)

    def log_base(self):
        return np.log(self.base)

    def log_partition(self):
        return self._stack().log_partition()[0]

    def log_likelihood(self, x):
        return self._stack().log_likelihood((np.asarray(x[0])[None], np.asarray(x[1])[None]))

    def expected_statistics(self):
        return tuple(s[0] for s in self._stack().expected_statistics())

    def entropy(self):
        return self._stack().entropy()[0]

    def cross_entropy(self, dist):
        return self._stack().cross_entropy(dist._stack())[0]


# ---------------------------------------------------------------------------------------
class StackedNormalGammas(_StackedPrior):
    """kappas, alphas, betas are per dimension: all parameters are (K, d)."""

    def __init__(self, size, dim, mus=None, kappas=None, alphas=None, betas=None):
        self.size, self.dim = size, dim
        self.mus, self.kappas, self.alphas, self.betas = (None if v is None else np.array(v, dtype=np.float64)
                                                         for v in (mus, kappas, alphas, betas))

    @property
    def params(self):
        return self.mus, self.kappas, self.alphas, self.betas

    @params.setter
    def params(self, values):
        self.mus, self.kappas, self.alphas, self.betas = (np.array(v, dtype=np.float64) for v in values)

    def std_to_nat(self, params):
        mus, kappas, alphas, betas = params
        return Stats([kappas * mus, np.array(kappas), 2. * alphas - 1., 2. * betas + kappas * mus ** 2])

    def nat_to_std(self, natparam):
        a, b, c, e = natparam
        mus = a / b
        alphas, betas = 0.5 * (c + 1.), 0.5 * (e - b * mus ** 2)
        if self._tied:
            alphas = np.array(self.size * [np.mean(alphas, axis=0)])
            betas = np.array(self.size * [np.mean(betas, axis=0)])
        return mus, np.array(b), alphas, betas

    def mean(self):
        return self.mus, self.alphas / self.betas

    def mode(self):
        return self.mus, (self.alphas - 0.5) / self.betas

    def rvs(self):
        """per component: gamma draws then normals (composite.py:347-351)."""
        K, d = self.size, self.dim
        var = np.empty((K, 2 * d))
        for k in range(K):
            var[k, :d] = npr.gamma(self.alphas[k], 1. / self.betas[k])
            var[k, d:] = npr.normal(size=d)
        feats = E.diag_features(d)
        out = E.ng_posterior([E.to_dev(p) for p in self.params], E.zeros((K, feats.F)), feats.F, mode=1,
                             variates=var, want_lik=True, want_vlb=False)
        return E.to_host(out['lik_mu']), E.to_host(out['lik_lmbda'])

    def log_partition(self):
        return -0.5 * np.sum(np.log(self.kappas), axis=1) \
            + np.sum(gammaln(self.alphas) - self.alphas * np.log(self.betas), axis=1)

    def log_likelihood(self, x):
        mus, lmbdas = x
        prec = self.kappas * lmbdas
        ll_mu = np.sum(-0.5 * prec * (mus - self.mus) ** 2 + 0.5 * np.log(prec) - 0.5 * LOG_2PI, axis=1)
        ll_g = np.sum((self.alphas - 1.) * np.log(lmbdas) - self.betas * lmbdas
                      - gammaln(self.alphas) + self.alphas * np.log(self.betas), axis=1)
        return np.sum(ll_mu + ll_g)

    def expected_statistics(self):
        E_lm = self.alphas / self.betas * self.mus
        return (E_lm, -0.5 * (1. / self.kappas + self.mus * E_lm),
                0.5 * (digamma(self.alphas) - np.log(self.betas)), -0.5 * self.alphas / self.betas)

    @staticmethod
    def _dot(nat, stats):
        return sum(np.sum(n * s, axis=1) for n, s in zip(nat, stats))


class TiedNormalGammas(StackedNormalGammas):
    _tied = True


class NormalGamma:

    def __init__(self, dim, mu=None, kappas=None, alphas=None, betas=None):
        self.dim = dim
        self.mu, self.kappas, self.alphas, self.betas = mu, kappas, alphas, betas

    def _stack(self):
        return StackedNormalGammas(1, self.dim, *[np.asarray(p)[None] for p in self.params])

    @property
    def params(self):
        return self.mu, self.kappas, self.alphas, self.betas

    @params.setter
    def params(self, values):
        self.mu, self.kappas, self.alphas, self.betas = values

    @property
    def nat_param(self):
        return self.std_to_nat(self.params)

    @nat_param.setter
    def nat_param(self, natparam):
        self.params = self.nat_to_std(natparam)

    @staticmethod
    def std_to_nat(params):
        return Stats([params[1] * params[0], params[1], 2. * params[2] - 1., 2. * params[3] + params[1] * params[0] ** 2])

    @staticmethod
    def nat_to_std(natparam):
        mu = natparam[0] / natparam[1]
        return mu, natparam[1], 0.5 * (natparam[2] + 1.), 0.5 * (natparam[3] - natparam[1] * mu ** 2)

    def mean(self):
        return self.mu, self.alphas / self.betas

    def mode(self):
        return self.mu, (self.alphas - 0.5) / self.betas

    def rvs(self):
        mus, lmbdas = self._stack().rvs()
        return mus[0], lmbdas[0]

    @property
    def base(self):
        """(2 pi)^(-d/2): Gaussian base measure times the Gamma's 1 (composite.py:353-358)."""
        return np.power(2. * np.pi, -0.5 * self.dim)

    def log_base(self):
        return np.log(self.base)

    def log_partition(self):
        return self._stack().log_partition()[0]

    def expected_statistics(self):
        return tuple(s[0] for s in self._stack().expected_statistics())

    def entropy(self):
        return self._stack().entropy()[0]

    def cross_entropy(self, dist):
        return self._stack().cross_entropy(dist._stack())[0]


# ---------------------------------------------------------------------------------------
class StackedMatrixNormalWisharts(_StackedPrior):

    def __init__(self, size, column_dim, row_dim, Ms=None, Ks=None, psis=None, nus=None):
        self.size, self.column_dim, self.row_dim = size, column_dim, row_dim
        self.Ms, self.Ks, self.psis, self.nus = (None if v is None else np.array(v, dtype=np.float64)
                                                 for v in (Ms, Ks, psis, nus))

    @property
    def params(self):
        return self.Ms, self.Ks, self.psis, self.nus

    @params.setter
    def params(self, values):
        self.Ms, self.Ks, self.psis, self.nus = (np.array(v, dtype=np.float64) for v in values)

    def std_to_nat(self, params):
        Ms, Ks, psis, nus = params
        MK = np.einsum('kdl,klm->kdm', Ms, Ks)
        return Stats([MK, np.array(Ks), np.linalg.inv(psis) + np.einsum('kdm,khm->kdh', MK, Ms),
                      nus - self.row_dim - 1. + self.column_dim])

    def nat_to_std(self, natparam):
        a, b, c, e = natparam
        Ms = np.einsum('kdl,klh->kdh', a, np.linalg.inv(b))
        inner = c - np.einsum('kdl,klm,khm->kdh', Ms, b, Ms)
        nus = e + self.row_dim + 1. - self.column_dim
        if self._tied:
            psi = np.linalg.inv(np.mean(inner, axis=0))
            return Ms, np.array(b), np.array(self.size * [psi]), np.array(self.size * [np.mean(nus)])
        return Ms, np.array(b), np.linalg.inv(inner), nus

    def mean(self):
        return self.Ms, self.nus[:, None, None] * self.psis

    def mode(self):
        return self.Ms, (self.nus - self.row_dim)[:, None, None] * self.psis

    def rvs(self):
        from .lingauss import ExpertLayout
        o, c = self.row_dim, self.column_dim
        var = draw_wishart_variates(self.nus, o, o * c)
        lay = ExpertLayout(c, o, affine=True)          # any layout works: statistics are zero
        feats = E.quad_features(lay.D)
        out = E.mnw_posterior([E.to_dev(p) for p in self.params], E.zeros((self.size, feats.F)), feats.F,
                              lay.dev()['stat_idx'], lay.D + 1, mode=1, variates=var, want_lik=True, want_vlb=False)
        out['info'].check()
        return E.to_host(out['lik_A']), E.to_host(out['lik_lmbda'])

    def log_partition(self):
        return -0.5 * self.row_dim * _logdet_spd(self.Ks) + _wishart_log_partition(self.psis, self.nus)

    def expected_statistics(self):
        E_LA = self.nus[:, None, None] * np.einsum('kdl,klm->kdm', self.psis, self.Ms)
        E_ALA = -0.5 * (self.row_dim * np.linalg.inv(self.Ks) + np.einsum('kdl,kdm->klm', self.Ms, E_LA))
        return E_LA, E_ALA, -0.5 * self.nus[:, None, None] * self.psis, 0.5 * _wishart_elogdet(self.psis, self.nus)

    @staticmethod
    def _dot(nat, stats):
        return np.einsum('kdl,kdl->k', nat[0], stats[0]) + np.einsum('kdl,kdl->k', nat[1], stats[1]) \
            + np.einsum('kdl,kdl->k', nat[2], stats[2]) + nat[3] * stats[3]


class TiedMatrixNormalWisharts(StackedMatrixNormalWisharts):
    _tied = True


class MatrixNormalWishart:

    def __init__(self, column_dim, row_dim, M=None, K=None, psi=None, nu=None):
        self.column_dim, self.row_dim = column_dim, row_dim
        self.M, self.K, self.psi, self.nu = M, K, psi, nu

    def _stack(self):
        return StackedMatrixNormalWisharts(1, self.column_dim, self.row_dim, np.asarray(self.M)[None],
                                           np.asarray(self.K)[None], np.asarray(self.psi)[None], np.atleast_1d(self.nu))

    @property
    def params(self):
        return self.M, self.K, self.psi, self.nu

    @params.setter
    def params(self, values):
        self.M, self.K, self.psi, self.nu = values

    @property
    def nat_param(self):
        return Stats([s[0] for s in self._stack().nat_param])

    @nat_param.setter
    def nat_param(self, natparam):
        self.M, self.K, self.psi, self.nu = self.nat_to_std(natparam)

    def std_to_nat(self, params):
        """[M K, K, psi^-1 + M K M^T, nu - o - 1 + c]  (composite.py:577-592)."""
        M, K, psi, nu = params
        st = StackedMatrixNormalWisharts(1, self.column_dim, self.row_dim)
        return Stats([s[0] for s in st.std_to_nat((np.asarray(M)[None], np.asarray(K)[None], np.asarray(psi)[None], np.atleast_1d(nu)))])

    def nat_to_std(self, natparam):
        """inverse of std_to_nat (composite.py:594-599)."""
        st = StackedMatrixNormalWisharts(1, self.column_dim, self.row_dim)
        out = st.nat_to_std([np.asarray(n)[None] for n in natparam])
        return tuple(o[0] for o in out)

    @property
    def base(self):
        """(2 pi)^(-o c / 2): matrix-normal base measure times the Wishart's 1 (composite.py:615-620, matrix.py:127-129)."""
        return np.power(2. * np.pi, -0.5 * self.row_dim * self.column_dim)

    def log_base(self):
        return np.log(self.base)

    def mean(self):
        return self.M, self.nu * self.psi

    def mode(self):
        return self.M, (self.nu - self.row_dim) * self.psi

    def rvs(self, size=1):
        As, lmbdas = self._stack().rvs()
        return As[0], lmbdas[0]

    def log_partition(self):
        return self._stack().log_partition()[0]

    def expected_statistics(self):
        return tuple(s[0] for s in self._stack().expected_statistics())

    def entropy(self):
        return self._stack().entropy()[0]

    def cross_entropy(self, dist):
        return self._stack().cross_entropy(dist._stack())[0]
