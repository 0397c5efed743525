"""Hierarchical mixtures of Gaussians (API of mimo/mixtures/hgmm.py).

  BayesianMixtureOfGaussiansWithHierarchicalPrior   K tied Gaussians under a hierarchical Normal-Wishart prior
  MixtureOfMixtureOfGaussians / BayesianMixtureOfMixtureOfGaussians   a mixture whose components are such mixtures

Everything per point x per component goes through the same device session as mixtures/gmm.py: the observations stay
resident, one fused sweep per iteration gives the responsibilities' statistics and sum_n logsumexp, and the nested
sub-iterations of the hierarchical prior (distributions/bayesian.py) see only K (d + 1) + F reduced statistics.
"""
import numpy as np
import numpy.random as npr
import torch
from tqdm import tqdm

from .. import _engine as E
from ..distributions.bayesian import MEANFIELD, GIBBS
from ..utils.data import batches
from ._driver import random_responsibilities
from .gmm import BayesianMixtureOfGaussians, _as_obs


class BayesianMixtureOfGaussiansWithHierarchicalPrior(BayesianMixtureOfGaussians):
    """gating: CategoricalWith{Dirichlet,StickBreaking}; components: TiedGaussiansWithHierarchicalNormalWisharts
    (hgmm.py:118-295).  expected_log_complete_likelihood / expected_responsibilities / expected_log_likelihood and the
    public lower-bound pieces are inherited: they only need the components' operand hooks."""

    def __init__(self, size, dim, gating, components, precision=None):
        assert components.size == size and components.dim == dim
        super().__init__(gating=gating, components=components, precision=precision)

    def used_labels(self, obs):
        raise NotImplementedError

    # -- Gibbs (hgmm.py:137-161) --------------------------------------------------------------
    def resample(self, obs, maxiter=250, maxsubiter=5, progress_bar=True, process_id=0):
        """labels from the current likelihood parameters -> gating -> components (with their sub-iterations); all
        variates from the global numpy.random stream in the reference's order."""
        s = self._session(obs)
        self.components.nb_iter = maxsubiter
        lay = s.parts[0].layout(s.D + 1, 0)
        buf = None
        with tqdm(total=maxiter, desc=f'Init #{process_id + 1}', position=process_id, disable=not progress_bar) as pbar:
            for _ in range(maxiter):
                ops = s.operands_from_likelihood(self.likelihood._log_probs())
                buf = s.sweep(ops, hard=True, uniforms=npr.random(size=(1, s.N)))
                counts = s.counts_host()
                g = self.gating._update(s.stat, s.F, s.count_feature, GIBBS, variates=self.gating._draw_variates(counts),
                                        prior_dev=s.gating_prior)
                g['info'].check()
                self.gating._store(g)
                self.components._update(s.stat, s.F, lay, GIBBS)
                pbar.update(1)
        if buf is not None:
            self.labels_ = E.to_host(buf.labels)

    def resample_components(self, obs, labels, maxsubiter):
        from ..utils.data import one_hot
        self.components.resample(obs, one_hot(labels, K=self.size), maxsubiter)

    # -- mean field (hgmm.py:186-225) ------------------------------------------------------------
    def meanfield_coordinate_descent(self, obs, randomize=True, weights=None, maxiter=250, maxsubiter=5, tol=1e-8,
                                     progress_bar=True, process_id=0, rtol=0., sample_likelihood=False):
        """Per iteration: reduced statistics -> sub-iterations of the hierarchical prior -> operands -> one fused E-step +
        statistics sweep; the lower bound is the parameter terms + sum_n logsumexp (the responsibilities are the E-step
        of the posterior the bound is evaluated at).  weights (N,): per-point weights multiplying the responsibilities
        in the parameter update (how a mixture of mixtures trains its components, hgmm.py:202, 422-431).
        sample_likelihood=True also performs the reference's per-iteration gating.likelihood.params = posterior.rvs()
        (bayesian.py:83): nothing in this loop reads it, but it advances numpy.random, which a mixture of
        mixtures draws its next cluster's random start from."""
        s = self._session(obs)
        self.components.nb_iter = maxsubiter
        w = None if weights is None else E.to_dev(np.asarray(weights, dtype=np.float64), E.tdtype(s.precision))
        resp = None
        if randomize:
            r0 = random_responsibilities(self.size, s.N)
            if w is None:
                s.stats_from_resp(r0)
            else:
                resp = E.to_dev(r0, E.tdtype(s.precision))
        elif w is None:
            s.sweep(s.operands_from_posterior(), hard=False)
        else:
            resp = s.loglik(s.operands_from_posterior())
            E.softmax(resp, s.precision, resp=True)
        vlb, outs = [], None
        with tqdm(total=maxiter, desc=f'VI #{process_id + 1}', position=process_id, disable=not progress_bar) as pbar:
            for _ in range(maxiter):
                if w is not None:
                    s.stats_from_resp((resp * w[None, :]).contiguous())
                ops, outs = s.update_from_stats(MEANFIELD)
                s.check(outs)
                if w is None:
                    s.sweep(ops, hard=False)
                    lse = s.lse_sum
                else:
                    resp = s.loglik(ops)
                    lse = E.softmax(resp, s.precision, resp=True, lse_sum=True)['lse_sum']
                vlb.append(float((lse.reshape(()) + outs['gating']['vlb'].reshape(()) + outs['parts'][0]['vlb'].sum()).item()))   # one read
                if sample_likelihood:
                    self.gating._store(outs['gating'], set_probs=False)
                    self.gating.likelihood.params = self.gating.posterior.rvs()
                if len(vlb) > 1 and (abs(vlb[-1] - vlb[-2]) < tol or abs(vlb[-1] - vlb[-2]) < rtol * abs(vlb[-1])):
                    break
                pbar.update(1)
        if outs is not None:
            s.store(outs, MEANFIELD, set_probs=False)
        return vlb

    def meanfield_update_parameters(self, obs, resp, maxsubiter):
        self.meanfield_update_components(obs, resp, maxsubiter)
        self.meanfield_update_gating(resp)

    def meanfield_update_components(self, obs, resp, maxsubiter):
        self.components.meanfield_update(obs, resp, maxsubiter)

    # -- SVI (hgmm.py:228-262: full-batch natural-gradient steps) -------------------------------------
    def meanfield_stochastic_descent(self, obs, randomize=True, weights=None, maxiter=250, maxsubiter=5, scale=1,
                                     step_size=1e-2, progress_bar=True, procces_id=0):
        obs = _as_obs(obs)
        resp = random_responsibilities(self.size, len(obs)) if randomize is True else self.expected_responsibilities(obs)
        with tqdm(total=maxiter, desc=f'SVI #{procces_id + 1}', position=procces_id, disable=not progress_bar) as pbar:
            for _ in range(maxiter):
                resp = resp if weights is None else resp * weights
                self.meanfield_sgd_parameters(obs, resp, maxsubiter, scale, step_size)
                resp = self.expected_responsibilities(obs)
                pbar.update(1)
        return []

    def meanfield_sgd_parameters(self, obs, resp, maxsubiter, scale, step_size):
        self.meanfield_sgd_components(obs, resp, maxsubiter, scale, step_size)
        self.meanfield_sgd_gating(resp, scale, step_size)

    def meanfield_sgd_components(self, obs, resp, maxsubiter, scale, step_size):
        self.components.meanfield_sgd(obs, resp, maxsubiter, scale, step_size)

    def _lower_bound_at_posterior(self, obs):
        s = self._session(obs)
        s.sweep(s.operands_from_posterior(), hard=False)
        return float(self.gating.variational_lowerbound() + self.components.variational_lowerbound() + s.lse_sum.item())


# ---------------------------------------------------------------------------------------------------------------
def _log_weights(gating):
    from ..distributions.bayesian import CategoricalWithDirichlet
    if isinstance(gating, CategoricalWithDirichlet):
        return gating.expected_log_likelihood()
    log_stick, log_rest = gating.expected_log_likelihood()
    return log_stick + np.hstack((0, np.cumsum(log_rest)[:-1]))


def _softmax_rows(log_lik):
    """(M, N) host log-joint -> responsibilities, on the device (utils/stats.py softmax kernel)."""
    a = E.to_dev(log_lik, E.tdtype(E.default_precision()))
    E.softmax(a, E.default_precision(), resp=True)
    return E.to_host(a).astype(np.float64)


class MixtureOfMixtureOfGaussians:
    """hgmm.py:16-89: `components` is a list of cluster_size MixtureOfGaussians."""

    def __init__(self, cluster_size, mixture_size, dim, gating, components):
        self.cluster_size, self.mixture_size, self.dim = cluster_size, mixture_size, dim
        self.gating, self.components = gating, components

    @property
    def params(self):
        raise NotImplementedError

    @property
    def nb_params(self):
        raise NotImplementedError

    def used_labels(self, obs):
        raise NotImplementedError

    def rvs(self, size=1):
        raise NotImplementedError

    def log_complete_likelihood(self, obs):
        comp = np.stack([c.log_likelihood(obs) for c in self.components])
        return comp + self.gating.log_likelihood(np.arange(self.cluster_size))[:, None]

    def log_likelihood(self, obs):
        a = E.to_dev(self.log_complete_likelihood(obs), torch.float64)
        return E.to_host(E.softmax(a, 'fp64', lse=True)['lse'])

    def responsibilities(self, obs):
        return _softmax_rows(self.log_complete_likelihood(obs))

    def max_likelihood(self, obs, randomize=True, maxiter=250, maxsubiter=5, progress_bar=True, process_id=0):
        resp = random_responsibilities(self.cluster_size, len(obs)) if randomize else self.responsibilities(obs)
        log_lik = []
        with tqdm(total=maxiter, desc=f'EM #{process_id + 1}', position=process_id, disable=not progress_bar) as pbar:
            for i in range(maxiter):
                for m in range(self.cluster_size):
                    self.components[m].max_likelihood(obs, weights=resp[m, :], randomize=randomize if i == 0 else False,
                                                      maxiter=maxsubiter, progress_bar=False)
                self.gating.max_likelihood(None, resp)
                resp = self.responsibilities(obs)
                log_lik.append(np.sum(self.log_likelihood(obs)))
                pbar.update(1)
        return log_lik

    def plot(self, *args, **kwargs):
        raise NotImplementedError('plotting is outside the scope of mimo_b200')


class BayesianMixtureOfMixtureOfGaussians:
    """hgmm.py:298-504: `components` is a list of cluster_size BayesianMixtureOfGaussiansWithHierarchicalPrior."""

    def __init__(self, cluster_size, mixture_size, dim, gating, components):
        self.cluster_size, self.mixture_size, self.dim = cluster_size, mixture_size, dim
        self.gating, self.components = gating, components
        self.likelihood = MixtureOfMixtureOfGaussians(cluster_size, mixture_size, dim, gating=gating.likelihood,
                                                      components=[c.likelihood for c in components])

    def used_labels(self, obs):
        raise NotImplementedError

    # -- Gibbs -------------------------------------------------------------------------------
    def resample(self, obs, init_labels='prior', maxiter=250, maxsubiter=100, maxsubsubiter=5, progress_bar=True, process_id=0):
        obs = _as_obs(obs)
        if init_labels == 'random':
            labels = npr.choice(self.cluster_size, size=(len(obs)))
        elif init_labels == 'prior':
            labels = self.gating.likelihood.rvs(len(obs))
        elif init_labels == 'posterior':
            _, labels = self.resample_labels(obs)
        with tqdm(total=maxiter, desc=f'Init #{process_id + 1}', position=process_id, disable=not progress_bar) as pbar:
            for _ in range(maxiter):
                self.resample_components(obs, labels, maxsubiter, maxsubsubiter)
                self.resample_gating(labels)
                _, labels = self.resample_labels(obs)
                pbar.update(1)
        self.labels_ = np.asarray(labels, dtype=np.int32)

    def resample_labels(self, obs):
        from ..utils.stats import sample_discrete_from_log
        log_prob = self.likelihood.log_complete_likelihood(obs)
        return log_prob, sample_discrete_from_log(log_prob, axis=0, precision='fp64')

    def resample_gating(self, labels):
        self.gating.resample(labels)

    def resample_components(self, obs, labels, maxsubiter, maxsubsubiter):
        for m in range(self.cluster_size):
            idx = np.where(labels == m)[0]
            self.components[m].resample(obs=obs[idx], maxiter=maxsubiter, maxsubiter=maxsubsubiter, progress_bar=False)

    # -- mean field ---------------------------------------------------------------------------
    def expected_log_complete_likelihood(self, obs):
        comp = np.stack([c.expected_log_likelihood(obs) for c in self.components])
        return comp + _log_weights(self.gating)[:, None]

    def expected_responsibilities(self, obs):
        return _softmax_rows(self.expected_log_complete_likelihood(obs))

    def meanfield_coordinate_descent(self, obs, randomize=True, maxiter=250, maxsubiter=5, maxsubsubiter=5, tol=1e-8,
                                     progress_bar=True, process_id=0):
        obs = _as_obs(obs)
        resp = random_responsibilities(self.cluster_size, len(obs)) if randomize else self.expected_responsibilities(obs)
        with tqdm(total=maxiter, desc=f'VI #{process_id + 1}', position=process_id, disable=not progress_bar) as pbar:
            for i in range(maxiter):
                self.meanfield_update_parameters(obs, resp, maxsubiter, maxsubsubiter, randomize if i == 0 else False)
                resp = self.expected_responsibilities(obs)
                pbar.update(1)
        return []

    def meanfield_update_parameters(self, obs, resp, maxsubiter, maxsubsubiter, randomize):
        self.meanfield_update_gating(resp)
        self.meanfield_update_components(obs, resp, maxsubiter, maxsubsubiter, randomize)

    def meanfield_update_gating(self, resp):
        self.gating.meanfield_update(None, resp)

    def meanfield_update_components(self, obs, resp, maxsubiter, maxsubsubiter, randomize):
        for m in range(self.cluster_size):
            self.components[m].meanfield_coordinate_descent(obs=obs, randomize=randomize, weights=resp[m, :], maxiter=maxsubiter,
                                                            maxsubiter=maxsubsubiter, progress_bar=False, sample_likelihood=True)

    # -- SVI ---------------------------------------------------------------------------------
    def meanfield_stochastic_descent(self, obs, randomize=True, maxiter=250, maxsubiter=5, maxsubsubiter=5, step_size=1e-2,
                                     batch_size=128, progress_bar=True, procces_id=0):
        obs = _as_obs(obs)
        with tqdm(total=maxiter, desc=f'SVI #{procces_id + 1}', position=procces_id, disable=not progress_bar) as pbar:
            scale = batch_size / float(len(obs))
            for i in range(maxiter):
                rnd = randomize if i == 0 else False
                for batch in batches(batch_size, len(obs)):
                    resp = random_responsibilities(self.cluster_size, len(batch)) if rnd is True \
                        else self.expected_responsibilities(obs[batch, :])
                    self.meanfield_sgd_parameters(obs[batch, :], resp, maxsubiter, maxsubsubiter, rnd, scale, step_size)
                pbar.update(1)
        return []

    def meanfield_sgd_parameters(self, obs, resp, maxsubiter, maxsubsubiter, randomize, scale, step_size):
        self.meanfield_sgd_components(obs, resp, maxsubiter, maxsubsubiter, randomize, scale, step_size)
        self.meanfield_sgd_gating(resp, scale, step_size)

    def meanfield_sgd_components(self, obs, resp, maxsubiter, maxsubsubiter, randomize, scale, step_size):
        for m in range(self.cluster_size):
            self.components[m].meanfield_stochastic_descent(obs=obs, randomize=randomize, weights=resp[m, :], maxiter=maxsubiter,
                                                            maxsubiter=maxsubsubiter, scale=scale, step_size=step_size,
                                                            progress_bar=False)

    def meanfield_sgd_gating(self, resp, scale, step_size):
        self.gating.meanfield_sgd(None, resp, scale, step_size)

    def variational_lowerbound_labels(self, resp):
        raise NotImplementedError

    def variational_lowerbound_obs(self, obs, resp):
        raise NotImplementedError

    def plot(self, *args, **kwargs):
        raise NotImplementedError('plotting is outside the scope of mimo_b200')
