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
from .categorical import Categorical
from .composite import draw_wishart_variates
from .gaussian import (StackedGaussiansWithPrecision, TiedGaussiansWithPrecision,
                       StackedGaussiansWithDiagonalPrecision, TiedGaussiansWithDiagonalPrecision)
from .lingauss import (StackedLinearGaussiansWithPrecision, TiedLinearGaussiansWithPrecision, ExpertLayout)

MEANFIELD, GIBBS, MAP, NONE = 0, 1, 2, 3


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
