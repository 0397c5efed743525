"""Gating priors: Dirichlet and truncated stick-breaking (mirrors
mimo/distributions/dirichlet.py).  These are K-vectors of host state; inside a sweep the
posterior update, expected log-weights and lower-bound term run in the batched gating
kernel (mimo_gating_posterior) -- the scalar helpers here serve the object API."""
import warnings

import numpy as np
import numpy.random as npr
from scipy.special import digamma, gammaln, betaln


class _ShiftedByOne:
    """Both gating priors have natural parameters = standard parameters - 1 (Dirichlet: alpha - 1; stick-breaking:
    (gamma - 1, delta - 1)) and no base measure; `_fields` names the attributes that make up `params`."""
    _fields = ()
    base = 1.

    def log_base(self):
        return 0.

    def _get_params(self):
        vals = tuple(getattr(self, f) for f in self._fields)
        return vals[0] if len(vals) == 1 else vals

    def _set_params(self, values):
        if len(self._fields) == 1:
            values = (values,)
        for f, v in zip(self._fields, values):
            setattr(self, f, v)

    params = property(_get_params, _set_params)
    nat_param = property(lambda self: self.std_to_nat(self.params),
                         lambda self, nat: self._set_params(self.nat_to_std(nat)))

    @classmethod
    def std_to_nat(cls, params):
        return params - 1. if len(cls._fields) == 1 else tuple(p - 1. for p in params)

    @classmethod
    def nat_to_std(cls, natparam):
        return natparam + 1. if len(cls._fields) == 1 else tuple(n + 1. for n in natparam)


class Dirichlet(_ShiftedByOne):
    _fields = ('alphas',)

    def __init__(self, dim=None, alphas=None):
        self.dim, self.alphas = dim, alphas

    def mean(self):
        return self.alphas / np.sum(self.alphas)

    def mode(self):
        assert np.all(self.alphas > 1.), "Make sure alphas > 1."
        return (self.alphas - 1.) / (np.sum(self.alphas) - self.dim)

    def rvs(self, size=1):
        return npr.dirichlet(self.alphas)

    def statistics(self, data):
        """log x per point; a list of arrays maps to a list (dirichlet.py:52-58)."""
        if isinstance(data, np.ndarray):
            return np.log(data[~np.isnan(data).any(axis=1)])
        return [self.statistics(d) for d in data]

    def weighted_statistics(self, data, weights):
        """w_n log x_n per point (dirichlet.py:60-69)."""
        if isinstance(data, np.ndarray):
            keep = ~np.isnan(data).any(axis=1)
            return weights[keep][:, None] * np.log(data[keep])
        return [self.weighted_statistics(d, w) for d, w in zip(data, weights)]

    def log_partition(self):
        return np.sum(gammaln(self.alphas)) - gammaln(np.sum(self.alphas))

    def log_likelihood(self, x):
        return np.sum((self.alphas - 1.) * np.log(x)) - self.log_partition()

    def expected_statistics(self):
        return digamma(self.alphas) - digamma(np.sum(self.alphas))

    def entropy(self):
        return self.log_partition() - self.nat_param.dot(self.expected_statistics())

    def cross_entropy(self, dist):
        return dist.log_partition() - dist.nat_param.dot(self.expected_statistics())


class TruncatedStickBreaking(_ShiftedByOne):
    """Ishwaran & James (2001) / Blei & Jordan (2006) truncation."""
    _fields = ('gammas', 'deltas')

    def __init__(self, dim=None, gammas=None, deltas=None):
        self.dim, self.gammas, self.deltas = dim, gammas, deltas

    @staticmethod
    def _sticks_to_probs(betas):
        probs = np.empty(betas.shape)
        probs[0] = betas[0]
        probs[1:] = betas[1:] * np.cumprod(1. - betas[:-1])
        return probs

    def mean(self):
        betas = np.hstack((self.gammas[:-1] / (self.gammas[:-1] + self.deltas[:-1]), 1.))
        return self._sticks_to_probs(betas)

    def mode(self):
        betas = np.ones((self.dim,))
        for k in range(self.dim - 1):
            g, d = self.gammas[k], self.deltas[k]
            if g > 1. and d > 1.:
                betas[k] = (g - 1.) / (g + d - 2.)
            elif (g == 1. and d == 1.) or (g < 1. and d < 1.) or (g > 1. and d <= 1.):
                betas[k] = 1.
            elif g <= 1. and d > 1.:
                betas[k] = 0.
            else:
                warnings.warn("Mode of Dirichlet process not defined")
                raise ValueError
        return self._sticks_to_probs(betas)

    def rvs(self, size=1, truncate=True):
        betas = np.hstack((npr.beta(self.gammas[:-1], self.deltas[:-1]), 1.))
        return self._sticks_to_probs(betas)

    def log_partition(self):
        return np.sum(betaln(self.gammas, self.deltas))

    def log_likelihood(self, x):
        raise NotImplementedError

    def expected_statistics(self):
        both = digamma(self.gammas + self.deltas)
        return digamma(self.gammas) - both, digamma(self.deltas) - both

    def _dot(self, nat):
        e_stick, e_rest = self.expected_statistics()
        return nat[0].dot(e_stick) + nat[1].dot(e_rest)

    def entropy(self):
        return self.log_partition() - self._dot(self.nat_param)

    def cross_entropy(self, dist):
        return dist.log_partition() - self._dot(dist.nat_param)
