"""likelihood + prior + posterior conjugate wrappers (API of mimo/distributions/bayesian.py).

The reference writes every update as ``posterior.nat_param = prior.nat_param + stats`` and
then converts back with per-component Python loops and explicit inverses.  Here the whole
step -- conjugate update, Cholesky factors, expected natural parameters / posterior draw /
posterior mode, the packed E-step operands and the lower-bound term -- is ONE batched
kernel call per wrapper (mimo_nw_posterior, mimo_ng_posterior, mimo_mnw_posterior,
mimo_gating_posterior).  The public methods below run that call on freshly uploaded
statistics; the sweep drivers (mixtures/) call the same `_update` hooks on device-resident
statistics without any host round trip.
"""
import copy

import numpy as np
import numpy.random as npr

from .. import _engine as E
from ..utils.abstraction import Statistics as E_stats
from .categorical import Categorical
from .composite import draw_wishart_variates
from .gaussian import (StackedGaussiansWithPrecision, TiedGaussiansWithPrecision,
                       StackedGaussiansWithDiagonalPrecision, TiedGaussiansWithDiagonalPrecision)
from .lingauss import (StackedLinearGaussiansWithPrecision, TiedLinearGaussiansWithPrecision, ExpertLayout)

MEANFIELD, GIBBS, MAP, NONE = 0, 1, 2, 3


def torch_long():
    import torch
    return torch.int64


def torch_full_mean(t):
    """np.full_like(t, np.mean(t)) for a device tensor (tied degrees of freedom)."""
    return t.mean().expand_as(t).contiguous()


# ---------------------------------------------------------------------------------------
# gating
# ---------------------------------------------------------------------------------------
class _GatingBase:
    _kind = 0

    def __init__(self, dim, prior, likelihood=None):
        self.dim = dim
        self.prior = prior
        self.posterior = copy.deepcopy(prior)
        self.likelihood = likelihood if likelihood is not None else Categorical(dim=dim, probs=self.prior.rvs())

    def empirical_bayes(self, data):
        raise NotImplementedError

    def _counts(self, data, weights):
        return self.likelihood.statistics(data) if weights is None \
            else self.likelihood.weighted_statistics(data, weights)

    # -- device hooks used by the sweep drivers ---------------------------------------
    def _prior_dev(self):
        a, b = self._prior_arrays(self.prior)
        return E.to_dev(a), (E.to_dev(b) if b is not None else None)

    def _update(self, stat, F, count_feature, mode, ops=None, variates=None, prior_dev=None):
        pa, pb = prior_dev or self._prior_dev()
        return E.gating_posterior(self._kind, pa, pb, stat, F, count_feature, mode=mode, variates=variates, ops=ops)

    def _apply_counts(self, counts, mode, variates=None, set_probs=True):
        stat = E.to_dev(np.asarray(counts, dtype=np.float64).reshape(-1, 1))
        out = self._update(stat, 1, 0, mode, variates=variates)
        out['info'].check()
        self._store(out, set_probs=set_probs)
        return out

    def _store(self, out, set_probs=True):
        self._store_posterior(out)
        if set_probs:
            self.likelihood.params = E.to_host(out['probs'])

    # -- reference API -------------------------------------------------------------------
    def max_aposteriori(self, data, weights=None):
        try:
            self._apply_counts(self._counts(data, weights), MAP)
        except AssertionError:
            raise AssertionError("Make sure alphas > 1.")

    def resample(self, data):
        counts = self.likelihood.statistics(data)
        self._apply_counts(counts, GIBBS, variates=self._draw_variates(counts))

    def meanfield_update(self, data, weights=None):
        """bayesian.py:77-83 / 151-159: the posterior update, then likelihood.params = posterior.rvs() -- a valid
        probability vector drawn from the global numpy.random stream like the reference, NOT the
        unnormalised exp(E log pi) the mean-field kernel works with."""
        out = self._apply_counts(self._counts(data, weights), MEANFIELD, set_probs=False)
        self.likelihood.params = self.posterior.rvs()
        return out

    def meanfield_sgd(self, data, weights, scale, step_size):
        """bayesian.py:85-91 / 161-171 (ends with the same posterior.rvs() draw)."""
        counts = self._counts(data, weights)
        self._sgd_blend(counts, scale, step_size)
        self.likelihood.params = self.posterior.rvs()

    def variational_lowerbound(self):
        return self.posterior.entropy() - self.posterior.cross_entropy(self.prior)

    def expected_log_likelihood(self):
        return self.posterior.expected_statistics()


class CategoricalWithDirichlet(_GatingBase):
    """bayesian.py:36-99."""
    _kind = 0

    @staticmethod
    def _prior_arrays(prior):
        return prior.alphas, None

    def _store_posterior(self, out):
        self.posterior.alphas = E.to_host(out['a'])

    def _draw_variates(self, counts):
        # npr.dirichlet(alphas) == normalised standard_gamma(alphas) on the same stream
        return npr.standard_gamma(self.prior.alphas + counts)

    def _draw_variates_device(self, counts, gen, prior_dev):
        """the same gamma draws from the DEVICE generator (no host read of the counts): counts (K,) device FP64."""
        import torch
        return torch._standard_gamma(prior_dev[0] + counts, generator=gen)

    def _sgd_blend(self, counts, scale, step_size):
        self.posterior.nat_param = (1. - step_size) * self.posterior.nat_param \
            + step_size * (self.prior.nat_param + 1. / scale * counts)


class CategoricalWithStickBreaking(_GatingBase):
    """bayesian.py:102-179."""
    _kind = 1

    @staticmethod
    def _prior_arrays(prior):
        return prior.gammas, prior.deltas

    def _store_posterior(self, out):
        self.posterior.gammas = E.to_host(out['a'])
        self.posterior.deltas = E.to_host(out['b'])

    def _draw_variates(self, counts):
        acc = np.hstack((np.cumsum(counts[::-1])[-2::-1], 0))
        return npr.beta((self.prior.gammas + counts)[:-1], (self.prior.deltas + acc)[:-1])

    def _draw_variates_device(self, counts, gen, prior_dev):
        """K - 1 Beta(gamma_k, delta_k) draws as g1 / (g1 + g2) of two device gamma draws."""
        import torch
        tail = torch.flip(torch.cumsum(torch.flip(counts, [0]), 0), [0])       # sum_{j >= k} counts
        acc = torch.cat((tail[1:], tail.new_zeros(1)))
        g1 = torch._standard_gamma((prior_dev[0] + counts)[:-1], generator=gen)
        g2 = torch._standard_gamma((prior_dev[1] + acc)[:-1], generator=gen)
        return g1 / (g1 + g2)

    def _sgd_blend(self, counts, scale, step_size):
        acc = np.hstack((np.cumsum(counts[::-1])[-2::-1], 0))
        self.posterior.gammas = (1. - step_size) * self.posterior.gammas \
            + step_size * (self.prior.gammas + 1. / scale * counts)
        self.posterior.deltas = (1. - step_size) * self.posterior.deltas \
            + step_size * (self.prior.deltas + 1. / scale * acc)


# ---------------------------------------------------------------------------------------
# components
# ---------------------------------------------------------------------------------------
class _ComponentsBase:
    """common driver of the three conjugate families."""

    def empirical_bayes(self, *data):
        raise NotImplementedError

    def _precision(self):
        return self.likelihood.precision or E.default_precision()

    # the reference ends every mean-field / SGD update with likelihood.params = posterior.rvs()
    # (bayesian.py:230,238,391,399,844,852): the draw consumes the global numpy.random stream and leaves SAMPLED
    # likelihood parameters.  Kept here for the stand-alone wrappers; the fused sweep drivers skip it unless asked.
    sample_likelihood = True

    def _after_meanfield(self):
        if self.sample_likelihood:
            self.likelihood.params = self.posterior.rvs()

    def variational_lowerbound(self):
        return self.posterior.entropy() - self.posterior.cross_entropy(self.prior)

    def log_marginal_likelihood(self):
        return self.posterior.log_partition() - self.prior.log_partition()

    def _run(self, stat, mode, variates=None):
        """posterior update from device statistics of this wrapper's own layout; stores the
        result in the NumPy-facing objects."""
        out = self._update(stat, self._feats().F, self._own_layout(), mode, variates=variates,
                           want_lik=mode in (GIBBS, MAP))
        out['info'].check()
        self._store(out, mode)
        return out

    def _stats_from(self, weights, *data):
        """device statistics from host data and (K, N) weights."""
        precision = self._precision()
        z = np.hstack(data) if len(data) > 1 else data[0]
        good = ~np.isnan(z).any(axis=1)
        Z = E.to_dev(z[good], E.tdtype(precision))
        R = E.to_dev(np.asarray(weights)[:, good], E.tdtype(precision))
        return E.stats_soft(Z, R, self._feats(), precision)

    def _unit_weights(self, *data):
        return np.ones((self.size, len(data[0])))

    def _wishart_variates_device(self, nus, d, extra, gen):
        """draw_wishart_variates on the device: [normal(d(d-1)/2) | chisquare(nu - i), i < d | normal(extra)] per
        component, chi-square(df) = 2 Gamma(df / 2); nus (K,) device FP64.  Not the reference's stream."""
        import torch
        K = nus.shape[0]
        nt = d * (d - 1) // 2
        var = torch.empty((K, nt + d + extra), dtype=torch.float64, device=nus.device)
        var[:, :nt].normal_(generator=gen)
        df = nus[:, None] - torch.arange(d, dtype=torch.float64, device=nus.device)[None, :]
        var[:, nt:nt + d] = 2. * torch._standard_gamma(0.5 * df, generator=gen)
        var[:, nt + d:].normal_(generator=gen)
        return var

    def _counts_of(self, stat):
        return E.to_host(stat[:, self._feats().F - 1])


class StackedGaussiansWithNormalWisharts(_ComponentsBase):
    """bayesian.py:182-323 (stacked) -- Gaussian likelihoods with Normal-Wishart priors."""
    _tied = False
    _likelihood_cls = StackedGaussiansWithPrecision

    def __init__(self, size, dim, prior, likelihood=None):
        self.size, self.dim = size, dim
        self.prior = prior
        self.posterior = copy.deepcopy(prior)
        if likelihood is None:
            mus, lmbdas = prior.rvs()
            likelihood = self._likelihood_cls(size=size, dim=dim, mus=mus, lmbdas=lmbdas)
        self.likelihood = likelihood

    # -- device hooks --------------------------------------------------------------------
    def _feats(self):
        return E.quad_features(self.dim)

    def _own_layout(self):
        idx = E.identity_map(self.dim, self.dim)
        return dict(stat_idx=idx, col_map=idx, Dp=self.dim + 1, row_off=0)

    def _rows(self, mode):
        return self.dim

    def _prior_dev(self, dist=None):
        return [E.to_dev(p) for p in (dist or self.prior).params]

    _shardable = True          # one CTA per component, no coupling between components unless tied

    def _update(self, stat, F, layout, mode, ops=None, variates=None, prior_dev=None, want_lik=False, want_vlb=True,
                k_range=None):
        return E.nw_posterior(prior_dev or self._prior_dev(), stat, F, layout['stat_idx'], layout['Dp'], mode=mode,
                              tied=self._tied, variates=variates, ops=ops, row_off=layout['row_off'],
                              col_map=layout['col_map'], want_lik=want_lik, want_vlb=want_vlb, k_range=k_range)

    def _draw_variates(self, counts):
        nus = self.prior.nus + counts
        if self._tied:
            nus = np.full_like(nus, np.mean(nus))
        return draw_wishart_variates(nus, self.dim, self.dim)

    def _draw_variates_device(self, counts, stat, gen, prior_dev):
        nus = prior_dev[3] + counts
        if self._tied:
            nus = torch_full_mean(nus)
        return self._wishart_variates_device(nus, self.dim, self.dim, gen)

    def _store(self, out, mode):
        self.posterior.params = tuple(E.to_host(out[k]) for k in ('m', 'kappa', 'psi', 'nu'))
        if out.get('lik_mu') is not None:
            self.likelihood.params = (E.to_host(out['lik_mu']), E.to_host(out['lik_lmbda']))

    def _posterior_operands(self, ops, layout, dist=None):
        """operands of E_q[log N] under `dist` (default: the current posterior)."""
        F = E.quad_features(ops.D).F
        out = self._update(E.zeros((self.size, F)), F, layout, MEANFIELD, ops=ops,
                           prior_dev=self._prior_dev(dist or self.posterior), want_vlb=False)
        return out['info']

    def _likelihood_operands(self, ops, layout):
        return E.operands_gauss(ops, E.to_dev(self.likelihood.mus), E.to_dev(self.likelihood.lmbdas),
                                row_off=layout['row_off'], col_map=layout['col_map'])

    # -- reference API -------------------------------------------------------------------
    def _stats(self, data, weights):
        w = self._unit_weights(data) if weights is None else weights
        return self._stats_from(w, data)

    def max_aposteriori(self, data, weights=None):
        self._run(self._stats(data, weights), MAP)

    def resample(self, data, labels=None):
        stat = self._stats(data, labels)
        self._run(stat, GIBBS, variates=self._draw_variates(self._counts_of(stat)))

    def meanfield_update(self, data, weights=None):
        self._run(self._stats(data, weights), MEANFIELD)
        self._after_meanfield()

    def meanfield_sgd(self, data, weights, scale, step_size):
        stats = self.likelihood.weighted_statistics(data, self._unit_weights(data) if weights is None else weights)
        self.posterior.nat_param = (1. - step_size) * self.posterior.nat_param \
            + step_size * (self.prior.nat_param + 1. / scale * stats)
        self._after_meanfield()

    def expected_log_likelihood(self, x):
        precision = self._precision()
        ops = E.QuadOperands(self.size, self.dim, self.dim, precision)
        self._posterior_operands(ops, self._own_layout()).check()
        return E.to_host(E.loglik(E.to_dev(np.nan_to_num(x), E.tdtype(precision)), ops)).astype(np.float64)

    def posterior_predictive_gaussian(self):
        mus, kappas, psis, nus = self.posterior.params
        dfs = nus - self.dim + 1
        return mus, (dfs / (1. + 1. / kappas))[:, None, None] * psis

    def posterior_predictive_studentt(self):
        mus, lmbdas = self.posterior_predictive_gaussian()
        return mus, lmbdas, self.posterior.nus - self.dim + 1

    def log_posterior_predictive_gaussian(self, x):
        mus, lmbdas = self.posterior_predictive_gaussian()
        return StackedGaussiansWithPrecision(self.size, self.dim, mus, lmbdas,
                                             precision=self.likelihood.precision).log_likelihood(np.array(x))


class TiedGaussiansWithNormalWisharts(StackedGaussiansWithNormalWisharts):
    _tied = True
    _likelihood_cls = TiedGaussiansWithPrecision


class GaussianWithNormalWishart:
    """single Gaussian with a Normal-Wishart prior (bayesian.py:182-265) = stack of one."""

    def __init__(self, dim, prior, likelihood=None):
        from .composite import StackedNormalWisharts
        from .gaussian import GaussianWithPrecision
        self.dim = dim
        self.prior = prior
        self.posterior = copy.deepcopy(prior)
        if likelihood is None:
            mu, lmbda = prior.rvs()
            likelihood = GaussianWithPrecision(dim=dim, mu=mu, lmbda=lmbda)
        self.likelihood = likelihood
        self._S = StackedNormalWisharts

    def _stacked(self):
        sp = self._S(1, self.dim, *[np.asarray(p)[None] if np.ndim(p) else np.atleast_1d(p) for p in self.prior.params])
        w = StackedGaussiansWithNormalWisharts(1, self.dim, sp, likelihood=self.likelihood._stack())
        w.posterior = self._S(1, self.dim, *[np.asarray(p)[None] if np.ndim(p) else np.atleast_1d(p)
                                             for p in self.posterior.params])
        return w

    def _pull(self, w, lik=True):
        self.posterior.params = tuple(p[0] for p in w.posterior.params)
        if lik:
            self.likelihood.params = tuple(p[0] for p in w.likelihood.params)

    def max_aposteriori(self, data, weights=None):
        w = self._stacked()
        w.max_aposteriori(data, None if weights is None else np.asarray(weights)[None, :])
        self._pull(w)

    def resample(self, data, labels=None):
        w = self._stacked()
        w.resample(data, None if labels is None else np.asarray(labels)[None, :])
        self._pull(w)

    def meanfield_update(self, data, weights=None):
        w = self._stacked()
        w.meanfield_update(data, None if weights is None else np.asarray(weights)[None, :])
        self._pull(w)                       # likelihood.params = posterior.rvs() (bayesian.py:230)

    def variational_lowerbound(self):
        return self.posterior.entropy() - self.posterior.cross_entropy(self.prior)

    def expected_log_likelihood(self, x):
        return self._stacked().expected_log_likelihood(x)[0]

    def log_marginal_likelihood(self):
        return self.posterior.log_partition() - self.prior.log_partition()


class StackedGaussiansWithNormalGammas(_ComponentsBase):
    """bayesian.py:343-483 -- diagonal Gaussians with Normal-Gamma priors.

    bug_compat=True reproduces the reference's StackedNormalGammas setters
    (composite.py:472-484): the posterior's alphas / betas never leave their prior values
    (SURVEY q1).  The default is the textbook update of NormalGamma.nat_to_std (:332-337)."""
    _tied = False
    _likelihood_cls = StackedGaussiansWithDiagonalPrecision

    def __init__(self, size, dim, prior, likelihood=None, bug_compat=False):
        self.size, self.dim = size, dim
        self.prior = prior
        self.posterior = copy.deepcopy(prior)
        self.bug_compat = bug_compat
        if likelihood is None:
            mus, lmbdas_diags = prior.rvs()
            likelihood = self._likelihood_cls(size=size, dim=dim, mus=mus, lmbdas_diags=lmbdas_diags)
        self.likelihood = likelihood

    def _feats(self):
        return E.diag_features(self.dim)

    def _own_layout(self):
        return dict(Dp=self.dim + 1, row_off=0)

    def _prior_dev(self, dist=None):
        return [E.to_dev(p) for p in (dist or self.prior).params]

    def _update(self, stat, F, layout, mode, ops=None, variates=None, prior_dev=None, want_lik=False, want_vlb=True):
        return E.ng_posterior(prior_dev or self._prior_dev(), stat, F, mode=mode, tied=self._tied,
                              bug_compat=self.bug_compat, variates=variates, ops=ops,
                              want_lik=want_lik, want_vlb=want_vlb)

    def _draw_variates(self, counts, stat_host=None):
        """gamma draws need the posterior shape / rate: computed from the statistics on the
        host side of the boundary (K*d scalars), in the reference's per-component order."""
        K, d = self.size, self.dim
        m0, k0, a0, b0 = self.prior.params
        if self.bug_compat:
            al, be = a0, b0
        else:
            n = stat_host[:, 2 * d][:, None]
            kap = k0 + n
            m = (k0 * m0 + stat_host[:, :d]) / kap
            al = a0 + 0.5 * n
            be = b0 + 0.5 * (stat_host[:, d:2 * d] + k0 * m0 ** 2 - kap * m ** 2)
            if self._tied:
                al = np.array(K * [np.mean(al, axis=0)])
                be = np.array(K * [np.mean(be, axis=0)])
        var = np.empty((K, 2 * d))
        for k in range(K):
            var[k, :d] = npr.gamma(al[k], 1. / be[k])
            var[k, d:] = npr.normal(size=d)
        return var

    def _draw_variates_device(self, counts, stat, gen, prior_dev):
        """_draw_variates with the posterior shape / rate formed on the device from the packed statistics
        [sum r z | sum r z^2 | sum r] and the draws from the device generator: no host read per sweep."""
        import torch
        d = self.dim
        m0, k0, a0, b0 = prior_dev
        if self.bug_compat:
            al, be = a0, b0
        else:
            n = stat[:, 2 * d][:, None]
            kap = k0 + n
            m = (k0 * m0 + stat[:, :d]) / kap
            al = a0 + 0.5 * n
            be = b0 + 0.5 * (stat[:, d:2 * d] + k0 * m0 ** 2 - kap * m ** 2)
            if self._tied:
                al = al.mean(0, keepdim=True).expand_as(al)
                be = be.mean(0, keepdim=True).expand_as(be)
        var = torch.empty((self.size, 2 * d), dtype=torch.float64, device=stat.device)
        var[:, :d] = torch._standard_gamma(al.contiguous(), generator=gen) / be
        var[:, d:].normal_(generator=gen)
        return var

    def _store(self, out, mode):
        self.posterior.params = tuple(E.to_host(out[k]) for k in ('m', 'kappa', 'alpha', 'beta'))
        if out.get('lik_mu') is not None:
            self.likelihood.params = (E.to_host(out['lik_mu']), E.to_host(out['lik_lmbda']))

    def _posterior_operands(self, ops, layout=None, dist=None):
        F = self._feats().F
        saved = self.bug_compat
        self.bug_compat = False                 # operands of `dist` exactly as given
        try:
            out = self._update(E.zeros((self.size, F)), F, layout, MEANFIELD, ops=ops,
                               prior_dev=self._prior_dev(dist or self.posterior), want_vlb=False)
        finally:
            self.bug_compat = saved
        return out['info']

    def _likelihood_operands(self, ops, layout=None):
        E.operands_gauss_diag(ops, E.to_dev(self.likelihood.mus), E.to_dev(self.likelihood.lmbdas_diags))
        return E.Info()

    def _stats(self, data, weights):
        w = self._unit_weights(data) if weights is None else weights
        return self._stats_from(w, data)

    def max_aposteriori(self, data, weights=None):
        self._run(self._stats(data, weights), MAP)

    def resample(self, data, labels=None):
        stat = self._stats(data, labels)
        self._run(stat, GIBBS, variates=self._draw_variates(None, E.to_host(stat)))

    def meanfield_update(self, data, weights=None):
        self._run(self._stats(data, weights), MEANFIELD)
        self._after_meanfield()

    def meanfield_sgd(self, data, weights, scale, step_size):
        stats = self.likelihood.weighted_statistics(data, self._unit_weights(data) if weights is None else weights)
        self.posterior.nat_param = (1. - step_size) * self.posterior.nat_param \
            + step_size * (self.prior.nat_param + 1. / scale * stats)
        self._after_meanfield()

    def expected_log_likelihood(self, x):
        precision = self._precision()
        ops = E.DiagOperands(self.size, self.dim, precision)
        self._posterior_operands(ops).check()
        return E.to_host(E.loglik(E.to_dev(np.nan_to_num(x), E.tdtype(precision)), ops)).astype(np.float64)

    def posterior_predictive_gaussian(self):
        mus, kappas, alphas, betas = self.posterior.params
        return mus, (alphas / betas) / (1. + 1. / kappas)

    def posterior_predictive_studentt(self):
        mus, lmbda_diags = self.posterior_predictive_gaussian()
        return mus, lmbda_diags, 2. * self.posterior.alphas


class TiedGaussiansWithNormalGammas(StackedGaussiansWithNormalGammas):
    _tied = True
    _likelihood_cls = TiedGaussiansWithDiagonalPrecision


class StackedLinearGaussiansWithMatrixNormalWisharts(_ComponentsBase):
    """bayesian.py:796-985 -- linear-Gaussian experts with Matrix-Normal-Wishart priors."""
    _tied = False
    _likelihood_cls = StackedLinearGaussiansWithPrecision

    def __init__(self, size, column_dim, row_dim, prior, likelihood=None, affine=True):
        self.size, self.column_dim, self.row_dim, self.affine = size, column_dim, row_dim, affine
        self.prior = prior
        self.posterior = copy.deepcopy(prior)
        if likelihood is None:
            As, lmbdas = prior.rvs()
            likelihood = self._likelihood_cls(size, column_dim, row_dim, As=As, lmbdas=lmbdas, affine=affine)
        self.likelihood = likelihood
        self.layout = ExpertLayout(column_dim, row_dim, affine)

    def _feats(self):
        return E.quad_features(self.layout.D)

    def _own_layout(self):
        dev = self.layout.dev()
        return dict(stat_idx=dev['stat_idx'], col_map=dev['col_map'], Dp=self.layout.D + 1, row_off=0)

    def _rows(self, mode):
        return self.row_dim + self.column_dim if mode == MEANFIELD else self.row_dim

    def _prior_dev(self, dist=None):
        return [E.to_dev(p) for p in (dist or self.prior).params]

    def _update(self, stat, F, layout, mode, ops=None, variates=None, prior_dev=None, want_lik=False, want_vlb=True):
        return E.mnw_posterior(prior_dev or self._prior_dev(), stat, F, layout['stat_idx'], layout['Dp'], mode=mode,
                               tied=self._tied, variates=variates, ops=ops, row_off=layout['row_off'],
                               col_map=layout['col_map'], want_lik=want_lik, want_vlb=want_vlb)

    def _draw_variates(self, counts):
        nus = self.prior.nus + counts
        if self._tied:
            nus = np.full_like(nus, np.mean(nus))
        return draw_wishart_variates(nus, self.row_dim, self.row_dim * self.column_dim)

    def _draw_variates_device(self, counts, stat, gen, prior_dev):
        nus = prior_dev[3] + counts
        if self._tied:
            nus = torch_full_mean(nus)
        return self._wishart_variates_device(nus, self.row_dim, self.row_dim * self.column_dim, gen)

    def _store(self, out, mode):
        self.posterior.params = tuple(E.to_host(out[k]) for k in ('M', 'K', 'psi', 'nu'))
        if out.get('lik_A') is not None:
            self.likelihood.params = (E.to_host(out['lik_A']), E.to_host(out['lik_lmbda']))

    def _posterior_operands(self, ops, layout, dist=None):
        F = E.quad_features(ops.D).F
        out = self._update(E.zeros((self.size, F)), F, layout, MEANFIELD, ops=ops,
                           prior_dev=self._prior_dev(dist or self.posterior), want_vlb=False)
        return out['info']

    def _likelihood_operands(self, ops, layout):
        return E.operands_lingauss(ops, E.to_dev(self.likelihood.As), E.to_dev(self.likelihood.lmbdas),
                                   layout['row_off'], layout['col_map'])

    def _stats(self, x, y, weights):
        w = self._unit_weights(x) if weights is None else weights
        return self._stats_from(w, x, y)

    def max_aposteriori(self, x, y, weights=None):
        self._run(self._stats(x, y, weights), MAP)

    def resample(self, x, y, z=None):
        stat = self._stats(x, y, z)
        self._run(stat, GIBBS, variates=self._draw_variates(self._counts_of(stat)))

    def meanfield_update(self, x, y, weights=None):
        self._run(self._stats(x, y, weights), MEANFIELD)
        self._after_meanfield()

    def meanfield_sgd(self, x, y, weights, scale, step_size):
        stats = self.likelihood.weighted_statistics(x, y, self._unit_weights(x) if weights is None else weights)
        self.posterior.nat_param = (1. - step_size) * self.posterior.nat_param \
            + step_size * (self.prior.nat_param + 1. / scale * stats)
        self._after_meanfield()

    def expected_log_likelihood(self, x, y):
        precision = self._precision()
        ops = E.QuadOperands(self.size, self.layout.D, self._rows(MEANFIELD), precision)
        self._posterior_operands(ops, self._own_layout()).check()
        z = np.nan_to_num(np.hstack((np.atleast_2d(x), np.atleast_2d(y))))
        return E.to_host(E.loglik(E.to_dev(z, E.tdtype(precision)), ops)).astype(np.float64)

    def _augment(self, x):
        x = np.reshape(x, (-1, self.likelihood.input_dim))
        return np.hstack((x, np.ones((len(x), 1)))) if self.likelihood.affine else x

    def posterior_predictive_gaussian(self, x):
        """mus (K,N,o), lmbdas (K,N,o,o)   (bayesian.py:949-962)."""
        xt = self._augment(x)
        Ms, Ks, psis, nus = self.posterior.params
        dfs = nus - self.likelihood.row_dim + 1
        mus = np.einsum('kdl,nl->knd', Ms, xt)
        cs = 1. + np.einsum('nd,kdl,nl->kn', xt, np.linalg.inv(Ks), xt)
        return mus, np.einsum('kdl,k,kn->kndl', psis, dfs, 1. / cs)

    def posterior_predictive_studentt(self, x):
        mus, lmbdas = self.posterior_predictive_gaussian(x)
        return mus, lmbdas, self.posterior.nus - self.likelihood.row_dim + 1


class TiedLinearGaussiansWithMatrixNormalWisharts(StackedLinearGaussiansWithMatrixNormalWisharts):
    _tied = True
    _likelihood_cls = TiedLinearGaussiansWithPrecision


# ---------------------------------------------------------------------------------------
# hierarchical Normal-Wishart
# ---------------------------------------------------------------------------------------
def _hyper_nw_update(hyper_prior, kappas0, mus, xk, nk, sxx):
    """(rho, kappa, psi, nu) of the shared (tau, Lambda) given component means `mus` (K, d) and the statistics
    xk = sum r x (K, d), nk = sum r (K,), sxx = sum_k sum r x x^T (d, d): the closed form of bayesian.py:671-684
    (= :643-656, :709-722) with the sums over k taken once."""
    mu0, k0, psi0, nu0 = hyper_prior
    K = mus.shape[0]
    tot = np.sum(kappas0 + k0)
    rho = (kappas0 @ mus + K * k0 * mu0) / tot
    dm = mu0[None, :] - mus
    c = k0 * kappas0 / (k0 + kappas0)
    cross = mus.T @ xk
    inner = np.linalg.inv(psi0) + ((dm * c[:, None]).T @ dm + sxx - cross - cross.T + (mus * nk[:, None]).T @ mus) / K
    return rho, tot / K, np.linalg.inv(inner), np.sum(nu0 + nk + 1) / K


class TiedGaussiansWithHierarchicalNormalWisharts(_ComponentsBase):
    """K Gaussians with one shared precision Lambda: means mu_k ~ N(tau, (kappa_k Lambda)^-1) (`prior`, a
    TiedGaussiansWithScaledPrecision) and (tau, Lambda) ~ Normal-Wishart (`hyper_prior`)   (bayesian.py:595-793).

    What is per point runs on the GPU: the weighted statistics (mimo_stats_soft, or the fused sweep of the mixture
    drivers) and the expected log-likelihood (the whitened quad-form kernels with rows chol(nu psi) shared by all
    components, offsets -U m_k and the trace / log-det constants folded into cst).  The nested sub-iterations between
    the component means and the hyper-posterior only touch sum r x (K, d), sum r (K) and sum_k sum r x x^T (d, d): the
    (K, F) statistics are reduced over k on the device and these K (d + 1) + F numbers cross to the host, where the
    sub-iterations run in FP64 (O(nb_iter (K d^2 + d^3)), independent of N)."""
    _tied = True
    _shardable = False
    nb_iter = 5                     # sub-iterations of a sweep driven by mixtures/hgmm.py (its maxsubiter)
    track_bound = True              # evaluate the lower-bound term in every sweep update (hgmm.py:207 does; hilr.py:194 does
                                    # not -- and the FIRST evaluation fixes the cached factors of quirk q11)

    def __init__(self, size, dim, hyper_prior, prior, precision=None):
        self.size, self.dim = size, dim
        draws = [hyper_prior.rvs() for _ in range(size)]                     # the reference's stream order (:602-605)
        prior.mus = np.stack([t for t, _ in draws])
        prior.lmbdas = np.stack([l for _, l in draws])
        self.hyper_prior = hyper_prior
        self.hyper_posterior = copy.deepcopy(hyper_prior)
        self.prior = prior
        self.posterior = copy.deepcopy(prior)
        self.likelihood = TiedGaussiansWithPrecision(size=size, dim=dim, mus=self.prior.rvs(sizes=size * [1]),
                                                     lmbdas=prior.lmbdas, precision=precision)

    # -- statistics ------------------------------------------------------------------------
    def _feats(self):
        return E.quad_features(self.dim)

    def _own_layout(self):
        idx = E.identity_map(self.dim, self.dim)
        return dict(stat_idx=idx, col_map=idx, Dp=self.dim + 1, row_off=0)

    def _rows(self, mode):
        return self.dim

    def _prior_dev(self, dist=None):
        return None

    def _reduced(self, stat, layout=None):
        """device (K, F) packed statistics -> host (xk, nk, sum_k xxT_k): K (d + 1) + d (d + 1) / 2 numbers cross the
        boundary.  layout['stat_idx']: columns of zt holding this part's variables and the constant (a mixture of linear
        experts keeps the input density's variables inside [x | y | 1]); default: the part's own [z | 1]."""
        d = self.dim
        lid = None if layout is None else id(layout['stat_idx'])
        if getattr(self, '_red_layout', -1) != lid:                  # the index tensors of a layout are read back once
            self._red_layout = lid
            self._red_cols = list(range(d)) + [d] if layout is None else [int(c) for c in E.to_host(layout['stat_idx'])]
        cols = self._red_cols
        key = tuple(cols)
        if getattr(self, '_red_key', None) != key:
            v, c = cols[:d], cols[d]
            il = np.tril_indices(d)
            self._red_key = key
            self._red_lin = E.to_dev(np.array([E.tri(a, c) for a in v] + [E.tri(c, c)], dtype=np.int64), torch_long())
            self._red_quad = E.to_dev(np.array([E.tri(v[i], v[j]) for i, j in zip(*il)], dtype=np.int64), torch_long())
        lin = E.to_host(stat[:, self._red_lin])
        tri = E.to_host(stat[:, self._red_quad].sum(0))
        sxx = np.zeros((d, d))
        il = np.tril_indices(d)
        sxx[il] = tri
        sxx[(il[1], il[0])] = tri
        return lin[:, :d], lin[:, d], sxx

    def _suff(self, data, weights):
        return self._reduced(self._stats_from(weights, data))

    # -- the three coordinate updates on reduced statistics ----------------------------------------
    def _meanfield(self, xk, nk, sxx, nb_iter):
        """bayesian.py:662-684."""
        hp, k0 = tuple(self.hyper_prior.params), self.prior.kappas
        self.posterior.kappas = k0 + nk
        mus, hq = None, tuple(self.hyper_posterior.params)
        for _ in range(nb_iter):                         # (local arrays: the K per-component objects are written once)
            mus = (k0[:, None] * hq[0][None, :] + xk) / (k0 + nk)[:, None]
            hq = _hyper_nw_update(hp, k0, mus, xk, nk, sxx)
        if mus is not None:
            self.posterior.mus = mus
            self.hyper_posterior.params = hq

    def _gibbs(self, xk, nk, sxx, nb_iter):
        """bayesian.py:623-659; the draws come from the global numpy.random stream in the reference's order."""
        hp, k0 = tuple(self.hyper_prior.params), self.prior.kappas
        mus = lmbdas = None
        for _ in range(nb_iter):
            draws = [self.hyper_posterior.rvs() for _ in range(self.size)]
            self.prior.mus = np.stack([t for t, _ in draws])
            lmbdas = np.stack([l for _, l in draws])
            self.prior.lmbdas = lmbdas
            self.posterior.nat_param = self.prior.nat_param + E_stats([xk, nk])
            self.posterior.lmbdas = lmbdas
            mus = self.posterior.rvs(sizes=self.size * [1])
            self.hyper_posterior.params = _hyper_nw_update(hp, k0, mus, xk, nk, sxx)
        self.likelihood.mus, self.likelihood.lmbdas = mus, lmbdas

    def _sgd(self, xk, nk, sxx, nb_iter, scale, step_size):
        """bayesian.py:691-729."""
        hp, k0 = tuple(self.hyper_prior.params), self.prior.kappas
        xk, nk, sxx = xk / scale, nk / scale, sxx / scale
        for _ in range(nb_iter):
            tau, lmbda = self.hyper_posterior.mean()
            self.prior.mus = np.stack(self.size * [tau])
            self.prior.lmbdas = np.stack(self.size * [lmbda])
            self.posterior.nat_param = (1. - step_size) * self.posterior.nat_param \
                + step_size * (self.prior.nat_param + E_stats([xk, nk]))
            self.posterior.lmbdas = np.stack(self.size * [lmbda])
            params = _hyper_nw_update(hp, k0, self.posterior.mean(), xk, nk, sxx)
            self.hyper_posterior.nat_param = (1. - step_size) * self.hyper_posterior.nat_param \
                + step_size * self.hyper_posterior.std_to_nat(params)

    def _set_mode(self):
        _, lmbda = self.hyper_posterior.mode()
        self.likelihood.mus = self.posterior.mode()
        self.likelihood.lmbdas = np.stack(self.size * [lmbda])

    # -- reference API -------------------------------------------------------------------
    def resample(self, data, labels, nb_iter=5):
        self._gibbs(*self._suff(data, labels), nb_iter)

    def meanfield_update(self, data, weights, nb_iter=25):
        self._meanfield(*self._suff(data, weights), nb_iter)
        self._set_mode()

    def meanfield_sgd(self, data, weights, nb_iter, scale, step_size):
        self._sgd(*self._suff(data, weights), nb_iter, scale, step_size)
        self._set_mode()

    # -- operands of E_q log N(x | mu_k, Lambda) -------------------------------------------------
    def _expected_constants(self):
        """per component: 1/2 (E log det Lambda - log det E Lambda) - 1/2 tr(E Lambda omega_k^-1), what
        bayesian.py:731-749 adds to the Gaussian log-density of precision E Lambda = nu psi at the posterior means."""
        from scipy.special import digamma
        _, _, psi, nu = self.hyper_posterior.params
        d = self.dim
        logdet_psi = np.linalg.slogdet(psi)[1]
        e_logdet = np.sum(digamma((nu - np.arange(d)) / 2.)) + d * np.log(2.) + logdet_psi
        tr = np.einsum('dl,kld->k', nu * psi, np.linalg.inv(self.posterior.omegas))
        return 0.5 * (e_logdet - (d * np.log(nu) + logdet_psi)) - 0.5 * tr

    def _posterior_operands(self, ops, layout, dist=None):
        _, _, psi, nu = self.hyper_posterior.params
        info = E.operands_gauss(ops, E.to_dev(self.posterior.mus), E.to_dev(np.stack(self.size * [nu * psi])),
                                row_off=layout['row_off'], col_map=layout['col_map'])
        ops.cst += E.to_dev(self._expected_constants(), ops.cst.dtype)
        return info

    def _likelihood_operands(self, ops, layout):
        return E.operands_gauss(ops, E.to_dev(self.likelihood.mus), E.to_dev(self.likelihood.lmbdas),
                                row_off=layout['row_off'], col_map=layout['col_map'])

    def expected_log_likelihood(self, x):
        precision = self._precision()
        ops = E.QuadOperands(self.size, self.dim, self.dim, precision)
        self._posterior_operands(ops, self._own_layout()).check()
        return E.to_host(E.loglik(E.to_dev(np.nan_to_num(x), E.tdtype(precision)), ops)).astype(np.float64)

    # -- hooks of the sweep session (mixtures/_driver.py) ------------------------------------------
    def _update(self, stat, F, layout, mode, ops=None, variates=None, prior_dev=None, want_lik=False, want_vlb=True):
        red = self._reduced(stat, layout)
        if mode == GIBBS:
            self._gibbs(*red, self.nb_iter)
            info = self._likelihood_operands(ops, layout) if ops is not None else E.Info()
            return dict(info=info, vlb=None)
        assert mode == MEANFIELD, 'hierarchical components support Gibbs and mean-field sweeps'
        self._meanfield(*red, self.nb_iter)
        info = self._posterior_operands(ops, layout) if ops is not None else E.Info()
        vlb = E.zeros((self.size,))
        if want_vlb and self.track_bound:
            vlb[0] = float(self.variational_lowerbound())
        return dict(info=info, vlb=vlb)

    def _store(self, out, mode):
        if mode == MEANFIELD:
            self._set_mode()

    # -- lower bound, predictive -------------------------------------------------------------
    def variational_lowerbound(self):
        """bayesian.py:751-781 in closed form; the entropy of the means' posteriors uses the cached Cholesky factors of
        the reference (quirk q11, gaussian.py GaussianWithScaledPrecision.omega_chol)."""
        from scipy.special import digamma
        d, K = self.dim, self.size
        rho, kap, psi, nu = self.hyper_posterior.params
        k0 = self.prior.kappas
        hyper = self.hyper_posterior.entropy() - self.hyper_posterior.cross_entropy(self.hyper_prior)
        e_logdet = np.sum(digamma((nu - np.arange(d)) / 2.)) + d * np.log(2.) + np.linalg.slogdet(psi)[1]
        ent = sum(dist.entropy() for dist in self.posterior.dists)
        dm = self.posterior.mus - rho[None, :]
        el = nu * psi
        quad = np.einsum('kd,dl,kl->k', dm, el, dm)
        tr = np.einsum('dl,kld->k', el, np.linalg.inv(self.posterior.omegas))
        return K * hyper + ent - 0.5 * K * d * np.log(2. * np.pi) + np.sum(0.5 * d * np.log(k0) + 0.5 * e_logdet
                                                                        - 0.5 * k0 * d / kap - 0.5 * k0 * (quad + tr))

    def posterior_predictive_gaussian(self):
        _, _, psi, nu = self.hyper_posterior.params
        return self.posterior.mus, np.stack(self.size * [(nu - self.dim + 1) * psi])

    def log_posterior_predictive_gaussian(self, x):
        mus, lmbdas = self.posterior_predictive_gaussian()
        return StackedGaussiansWithPrecision(self.size, self.dim, mus, lmbdas,
                                             precision=self.likelihood.precision).log_likelihood(np.array(x))


class GaussianWithHierarchicalNormalWishart:
    """one Gaussian whose mean has a scaled-precision prior under a Normal-Wishart hyper-prior (bayesian.py:503-592):
    the K = 1 case of the updates above with the reference's slightly different bookkeeping (the mean's posterior
    carries the current precision draw / expectation)."""

    def __init__(self, dim, hyper_prior, prior, precision=None):
        from .gaussian import GaussianWithPrecision
        self.dim = dim
        tau, lmbda = hyper_prior.rvs()
        prior.mu, prior.lmbda = tau, lmbda
        self.hyper_prior = hyper_prior
        self.hyper_posterior = copy.deepcopy(hyper_prior)
        self.prior = prior
        self.posterior = copy.deepcopy(prior)
        self.likelihood = GaussianWithPrecision(dim=dim, mu=self.prior.rvs(), lmbda=lmbda, precision=precision)

    def empirical_bayes(self, data):
        raise NotImplementedError

    def _suff(self, data):
        """sum x, n, sum x x^T of the rows without NaN, computed on the GPU."""
        x, n, xx, _ = self.likelihood.statistics(data)
        return np.asarray(x)[None, :], np.atleast_1d(float(n)), np.asarray(xx)

    def _hyper(self, mu, xk, nk, sxx):
        return _hyper_nw_update(tuple(self.hyper_prior.params), np.atleast_1d(self.prior.kappa), mu[None, :], xk, nk, sxx)

    def resample(self, data, nb_iter=1):
        xk, nk, sxx = self._suff(data)
        mu = lmbda = None
        for _ in range(nb_iter):
            _, lmbda = self.hyper_posterior.rvs()
            self.posterior.kappa = self.prior.kappa + nk[0]
            self.posterior.mu = (self.prior.kappa * self.hyper_prior.mu + xk[0]) / (self.prior.kappa + nk[0])
            self.posterior.lmbda = lmbda
            mu = self.posterior.rvs()
            self.hyper_posterior.params = self._hyper(mu, xk, nk, sxx)
        self.likelihood.params = mu, lmbda

    def meanfield_update(self, data, nb_iter=25):
        xk, nk, sxx = self._suff(data)
        for _ in range(nb_iter):
            self.posterior.kappa = self.prior.kappa + nk[0]
            self.posterior.mu = (self.prior.kappa * self.hyper_posterior.mu + xk[0]) / (self.prior.kappa + nk[0])
            self.posterior.lmbda = self.hyper_posterior.wishart.mean()
            self.hyper_posterior.params = self._hyper(self.posterior.mu, xk, nk, sxx)
        _, lmbda = self.hyper_posterior.mean()
        self.likelihood.params = self.posterior.rvs(), lmbda
        return []

    def expected_log_likelihood(self, x):
        raise NotImplementedError

    def variational_lowerbound(self, x):
        raise NotImplementedError


class TiedAffineLinearGaussiansWithMatrixNormalWisharts(_ComponentsBase):
    """K linear-Gaussian experts y = A x + c_k + eps that share the slope A and the precision Lambda and keep their own
    offsets c_k   (bayesian.py:1222-1522): slope_prior Matrix-Normal on A, offset_prior scaled-precision Gaussians on
    c_k, precision_prior Wishart on Lambda.

    Per point: the weighted second moments of [x | y | 1] (mimo_stats_soft or the fused sweep) and the expected
    log-likelihood, which is the Matrix-Normal-Wishart expectation of the affine expert [A | c_k] with the block-
    diagonal column precision diag(K, kappa_k) -- the same mimo_mnw_posterior operands as the plain experts.  The
    nested sub-iterations between slope / precision and the offsets run on the (K, F) statistics in FP64 on the host
    (O(nb_iter K c^3), independent of N)."""
    _tied = True
    _shardable = False
    nb_iter = 25

    def __init__(self, size, column_dim, row_dim, slope_prior, offset_prior, precision_prior, likelihood=None, precision=None):
        from .lingauss import StackedAffineLinearGaussiansWithPrecision
        from .composite import StackedMatrixNormalWisharts
        self.size, self.column_dim, self.row_dim = size, column_dim, row_dim
        As = np.zeros((size, row_dim, column_dim))
        lmbdas = np.zeros((size, row_dim, row_dim))
        for k in range(size):                                   # the reference's stream order (:1233-1241)
            lmbdas[k] = precision_prior.rvs()
            slope_prior.V = lmbdas[k]
            As[k] = slope_prior.rvs()
        offset_prior.lmbdas = lmbdas
        cs = offset_prior.rvs(sizes=size * [1])
        self.slope_prior, self.offset_prior, self.precision_prior = slope_prior, offset_prior, precision_prior
        self.likelihood = StackedAffineLinearGaussiansWithPrecision(size, column_dim, row_dim, As, cs, lmbdas, precision=precision)
        self.slope_posterior = copy.deepcopy(slope_prior)
        self.offset_posterior = copy.deepcopy(offset_prior)
        self.precision_posterior = copy.deepcopy(precision_prior)
        self.layout = ExpertLayout(column_dim + 1, row_dim, affine=True)
        self._S = StackedMatrixNormalWisharts

    # -- the expert [A | c_k] as a stacked Matrix-Normal-Wishart (what :1391-1417 builds) -------------------------
    def _joint(self, slope, offset, precision):
        from scipy.linalg import block_diag
        K = self.size
        Ms = np.stack([np.hstack((slope.M, offset.mus[k][:, None])) for k in range(K)])
        Ks = np.stack([block_diag(slope.K, np.array([[offset.kappas[k]]])) for k in range(K)])
        return self._S(K, self.column_dim + 1, self.row_dim, Ms=Ms, Ks=Ks, psis=np.stack(K * [precision.psi]),
                       nus=np.array(K * [precision.nu], dtype=np.float64))

    def _mnw(self):
        prior = self._joint(self.slope_prior, self.offset_prior, self.precision_prior)
        w = StackedLinearGaussiansWithMatrixNormalWisharts(self.size, self.column_dim + 1, self.row_dim, prior=prior,
                                                           likelihood=self.likelihood._combined(), affine=True)
        w.posterior = self._joint(self.slope_posterior, self.offset_posterior, self.precision_posterior)
        return w

    # -- statistics ------------------------------------------------------------------------
    def _feats(self):
        return E.quad_features(self.layout.D)

    def _own_layout(self):
        dev = self.layout.dev()
        return dict(stat_idx=dev['stat_idx'], col_map=dev['col_map'], Dp=self.layout.D + 1, row_off=0)

    def _rows(self, mode):
        return self.row_dim + self.column_dim + 1 if mode == MEANFIELD else self.row_dim

    def _prior_dev(self, dist=None):
        return None

    def _moments(self, stat):
        """device (K, F) packed statistics of [x | y | 1] -> dict of the host arrays the updates are written in."""
        from .gaussian import unpack_quad
        c = self.column_dim
        yxt, xxt, yy, n = self.layout.split(unpack_quad(E.to_host(stat), self.layout.D + 1))     # xt = [x ; 1]
        return dict(xm=xxt[:, :c, c], ym=yxt[:, :, c], n=n, yx=yxt[:, :, :c], xx=xxt[:, :c, :c], yy=yy)

    def _suff(self, x, y, weights):
        return self._moments(self._stats_from(weights, x, y))

    # -- coordinate updates (bayesian.py:1260-1383) ---------------------------------------------
    def _slope_precision(self, m, cs):
        """slope posterior (M, K) and precision posterior (psi, nu) given the offsets cs (K, o)."""
        K = self.size
        M0, K0 = self.slope_prior.M, self.slope_prior.K
        psi0, nu0 = self.precision_prior.psi, self.precision_prior.nu
        G = (M0 @ K0)[None] + m['yx'] - cs[:, :, None] * m['xm'][:, None, :]
        Kk = K0[None] + m['xx']
        GKi = G @ np.linalg.inv(Kk)
        self.slope_posterior.M = GKi.sum(0) / K
        self.slope_posterior.K = Kk.sum(0) / K
        yc = np.einsum('kd,kl->kdl', m['ym'], cs)
        resid = m['yy'] - yc - yc.transpose(0, 2, 1) + m['n'][:, None, None] * np.einsum('kd,kl->kdl', cs, cs)
        dc = cs - self.offset_prior.mus
        pull = self.offset_prior.kappas[:, None, None] * np.einsum('kd,kl->kdl', dc, dc)
        inner = np.linalg.inv(psi0) + M0 @ self.slope_posterior.K @ M0.T \
            + (resid + pull - GKi @ G.transpose(0, 2, 1)).sum(0) / K
        self.precision_posterior.psi = np.linalg.inv(inner)
        self.precision_posterior.nu = np.sum(nu0 + m['n'] + 1) / K

    def _offsets(self, m, As, lmbdas):
        k0, mu0 = self.offset_prior.kappas, self.offset_prior.mus
        self.offset_posterior.mus = (k0[:, None] * mu0 + m['ym'] - np.einsum('kdl,kl->kd', As, m['xm'])) / (k0 + m['n'])[:, None]
        self.offset_posterior.kappas = k0 + m['n']
        self.offset_posterior.lmbdas = lmbdas

    def _gibbs(self, m, nb_iter):
        K = self.size
        As = lmbdas = cs = None
        for _ in range(nb_iter):
            cs = self.offset_posterior.rvs(sizes=K * [1])
            self._slope_precision(m, cs)
            As = np.zeros((K, self.row_dim, self.column_dim))
            lmbdas = np.zeros((K, self.row_dim, self.row_dim))
            for k in range(K):
                lmbdas[k] = self.precision_posterior.rvs()
                self.slope_posterior.V = lmbdas[k]
                As[k] = self.slope_posterior.rvs()
            self._offsets(m, As, lmbdas)
        self.likelihood.As, self.likelihood.cs, self.likelihood.lmbdas = As, cs, lmbdas

    def _meanfield(self, m, nb_iter):
        K = self.size
        for _ in range(nb_iter):
            self._slope_precision(m, self.offset_posterior.mean())
            lmbda = self.precision_posterior.mean()
            self.slope_posterior.V = lmbda
            self._offsets(m, np.stack(K * [self.slope_posterior.mean()]), np.stack(K * [lmbda]))

    def _set_mode(self):
        K = self.size
        self.likelihood.As = np.stack(K * [self.slope_posterior.mode()])
        self.likelihood.lmbdas = np.stack(K * [self.precision_posterior.mode()])
        self.likelihood.cs = self.offset_posterior.mode()

    # -- reference API -------------------------------------------------------------------
    def resample(self, x, y, z, nb_iter=25):
        self._gibbs(self._suff(x, y, z), nb_iter)

    def meanfield_update(self, x, y, weights, nb_iter=25):
        self._meanfield(self._suff(x, y, weights), nb_iter)
        self._set_mode()

    def meanfield_sgd(self, x, y, weights, nb_iter, scale, step_size):
        raise NotImplementedError

    def _consume_reference_draws(self):
        """The reference builds a fresh StackedLinearGaussiansWithMatrixNormalWisharts inside expected_log_likelihood and
        posterior_predictive_gaussian (bayesian.py:1411-1414, 1503-1506) whose constructor samples a likelihood from the
        prior (:805-809): K Wishart + matrix-normal draws that nothing reads.  The same variates are taken from
        numpy.random here so that a seeded script stays on the reference's stream (quirk q12)."""
        draw_wishart_variates(np.array(self.size * [self.precision_prior.nu], dtype=np.float64), self.row_dim,
                              self.row_dim * (self.column_dim + 1))

    def expected_log_likelihood(self, x, y):
        self._consume_reference_draws()
        return self._mnw().expected_log_likelihood(x, y)

    def variational_lowerbound(self):
        w = self._mnw()
        return w.posterior.entropy() - w.posterior.cross_entropy(w.prior)

    def posterior_predictive_gaussian(self, x):
        self._consume_reference_draws()
        return self._mnw().posterior_predictive_gaussian(x)

    def log_posterior_predictive_gaussian(self, x, y):
        from .gaussian import LOG_2PI
        mus, lmbdas = self.posterior_predictive_gaussian(x)
        dy = np.asarray(y)[None] - mus
        return -0.5 * self.row_dim * LOG_2PI + 0.5 * np.linalg.slogdet(lmbdas)[1] - 0.5 * np.einsum('knd,kndl,knl->kn', dy, lmbdas, dy)

    # -- hooks of the sweep session ------------------------------------------------------------
    def _posterior_operands(self, ops, layout, dist=None):
        return self._mnw()._posterior_operands(ops, layout)

    def _likelihood_operands(self, ops, layout):
        return E.operands_lingauss(ops, E.to_dev(self.likelihood._affine_As()), E.to_dev(self.likelihood.lmbdas),
                                   layout['row_off'], layout['col_map'])

    def _update(self, stat, F, layout, mode, ops=None, variates=None, prior_dev=None, want_lik=False, want_vlb=True):
        m = self._moments(stat)
        if mode == GIBBS:
            self._gibbs(m, self.nb_iter)
            info = self._likelihood_operands(ops, layout) if ops is not None else E.Info()
            return dict(info=info, vlb=None)
        assert mode == MEANFIELD, 'tied affine experts support Gibbs and mean-field sweeps'
        self._meanfield(m, self.nb_iter)
        info = self._posterior_operands(ops, layout) if ops is not None else E.Info()
        vlb = E.zeros((self.size,))
        if want_vlb:
            vlb.copy_(E.to_dev(np.asarray(self.variational_lowerbound(), dtype=np.float64)))
        return dict(info=info, vlb=vlb)

    def _store(self, out, mode):
        if mode == MEANFIELD:
            self._set_mode()


class AffineLinearGaussianWithMatrixNormalWishart(_ComponentsBase):
    """one expert y = A x + c + eps with a Matrix-Normal prior on the slope, a scaled-precision Gaussian prior on the
    offset and a Wishart prior on the precision: the Gibbs sampler of bayesian.py:1137-1219 (examples/lingauss).  The
    second moments of [x | y | 1] come from the statistics kernel; the alternation between (slope, precision) and the
    offset runs on them on the host."""

    def __init__(self, column_dim, row_dim, slope_prior, offset_prior, precision_prior, likelihood=None, precision=None):
        from .lingauss import AffineLinearGaussianWithPrecision
        self.column_dim, self.row_dim = column_dim, row_dim
        self.size = 1
        lmbda = precision_prior.rvs()
        slope_prior.V = lmbda
        A = slope_prior.rvs()
        offset_prior.lmbda = lmbda
        c = offset_prior.rvs()
        self.slope_prior, self.offset_prior, self.precision_prior = slope_prior, offset_prior, precision_prior
        self.likelihood = AffineLinearGaussianWithPrecision(column_dim, row_dim, A, c, lmbda, precision=precision)
        self.slope_posterior = copy.deepcopy(slope_prior)
        self.offset_posterior = copy.deepcopy(offset_prior)
        self.precision_posterior = copy.deepcopy(precision_prior)
        self.layout = ExpertLayout(column_dim + 1, row_dim, affine=True)

    def _feats(self):
        return E.quad_features(self.layout.D)

    def resample(self, x, y, nb_iter=25):
        from .gaussian import unpack_quad
        cdim = self.column_dim
        stat = self._stats_from(np.ones((1, len(x))), x, y)
        yxt, xxt, yy, n = (a[0] for a in self.layout.split(unpack_quad(E.to_host(stat), self.layout.D + 1)))
        xm, ym, yx, xx = xxt[:cdim, cdim], yxt[:, cdim], yxt[:, :cdim], xxt[:cdim, :cdim]
        M0, K0 = self.slope_prior.M, self.slope_prior.K
        psi0, nu0 = self.precision_prior.psi, self.precision_prior.nu
        k0, mu0 = self.offset_prior.kappa, self.offset_prior.mu
        A = c = lmbda = None
        for _ in range(nb_iter):
            c = self.offset_posterior.rvs()
            K = K0 + xx
            M = (M0 @ K0 + yx - np.outer(c, xm)) @ np.linalg.inv(K)
            self.slope_posterior.M, self.slope_posterior.K = M, K
            resid = yy - np.outer(ym, c) - np.outer(c, ym) + n * np.outer(c, c)
            self.precision_posterior.psi = np.linalg.inv(np.linalg.inv(psi0) + M0 @ K0 @ M0.T + resid
                                                         + k0 * np.outer(c - mu0, c - mu0) - M @ K @ M.T)
            self.precision_posterior.nu = nu0 + n + 1
            lmbda = self.precision_posterior.rvs()
            self.slope_posterior.V = lmbda
            A = self.slope_posterior.rvs()
            self.offset_posterior.mu = (k0 * mu0 + ym - A @ xm) / (k0 + n)
            self.offset_posterior.kappa = k0 + n
            self.offset_posterior.lmbda = lmbda
        self.likelihood.params = A, c, lmbda
