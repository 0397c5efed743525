"""Hierarchical mixtures of linear-Gaussian experts (API of mimo/mixtures/hilr.py).

  BayesianMixtureOfLinearGaussiansWithTiedActivation   input densities under a hierarchical Normal-Wishart prior
                                                       (shared precision), experts with a shared slope and precision
                                                       and their own offsets

The sweep is the one of mixtures/ilr.py: z = [x | y] resident on the device, basis + experts + gating in one operand
block per component, one fused E-step + statistics pass per iteration; the nested sub-iterations of the two
hierarchical wrappers (distributions/bayesian.py) run on the packed statistics.
"""
import numpy as np
import numpy.random as npr
from tqdm import tqdm

from .. import _engine as E
from ..distributions.bayesian import MEANFIELD, GIBBS
from ._driver import random_responsibilities
from .ilr import BayesianMixtureOfLinearGaussians


class BayesianMixtureOfLinearGaussiansWithTiedActivation(BayesianMixtureOfLinearGaussians):
    """hilr.py:79-291.  gating: CategoricalWith{Dirichlet,StickBreaking}; basis: TiedGaussiansWithHierarchicalNormalWisharts;
    models: TiedAffineLinearGaussiansWithMatrixNormalWisharts.  expected_log_complete_likelihood,
    expected_responsibilities and the public lower bound are inherited (they need the wrappers' operand hooks only)."""

    def __init__(self, size, input_dim, output_dim, gating, basis, models, scale=False, precision=None):
        super().__init__(size, input_dim, output_dim, gating, basis, models, scale=scale, precision=precision)

    def used_labels(self, x, y):
        raise NotImplementedError

    def rvs(self, size=1):
        raise NotImplementedError

    def _sub_iterations(self, maxsubiter):
        self.basis.nb_iter = self.models.nb_iter = maxsubiter

    # -- Gibbs (hilr.py:121-148) ----------------------------------------------------------------
    def resample(self, x, y, maxiter=250, maxsubiter=5, progress_bar=True, process_id=0):
        """labels from the current likelihood parameters -> gating -> input densities -> experts; all variates from the
        global numpy.random stream in the reference's order."""
        xx, yy = self._scaled(x, y)
        s = self._session(xx, yy)
        self._sub_iterations(maxsubiter)
        lays = [s.parts[0].layout(s.D + 1, 0), s.parts[1].layout(s.D + 1, self.basis._rows(GIBBS))]
        buf = None
        with tqdm(total=maxiter, desc=f'Init #{process_id + 1}', position=process_id, disable=not progress_bar) as pbar:
            for _ in range(maxiter):
                ops = s.operands_from_likelihood(self.likelihood._log_probs())
                buf = s.sweep(ops, hard=True, uniforms=npr.random(size=(1, s.N)))
                g = self.gating._update(s.stat, s.F, s.count_feature, GIBBS, variates=self.gating._draw_variates(s.counts_host()),
                                        prior_dev=s.gating_prior)
                g['info'].check()
                self.gating._store(g)
                self.basis._update(s.stat, s.F, lays[0], GIBBS)
                self.models._update(s.stat, s.F, lays[1], GIBBS)
                pbar.update(1)
        if buf is not None:
            self.labels_ = E.to_host(buf.labels)

    def resample_basis(self, x, z, maxsubiter):
        from ..utils.data import one_hot
        self.basis.resample(x, one_hot(z, K=self.size), maxsubiter)

    def resample_models(self, x, y, z, maxsubiter):
        from ..utils.data import one_hot
        self.models.resample(x, y, one_hot(z, K=self.size), maxsubiter)

    # -- mean field (hilr.py:175-218) ------------------------------------------------------------
    def meanfield_coordinate_descent(self, x, y, randomize=True, weights=None, maxiter=250, maxsubiter=5, tol=1e-16,
                                     progress_bar=True, process_id=0, lower_bound=False, sample_likelihood=False):
        """The reference returns an empty list (its lower-bound line is commented out, hilr.py:194); lower_bound=True
        returns the bound of every iteration instead (parameter terms + sum_n logsumexp from the fused sweep) and
        stops on tol like the other drivers.  sample_likelihood=True replays the draws the reference's loop makes and
        never reads (the gating's posterior.rvs(), SURVEY q3, and the experts its E-step samples, quirk q12): a mixture of
        mixtures takes its next cluster's random start from the same stream."""
        xx, yy = self._scaled(x, y)
        s = self._session(xx, yy)
        self._sub_iterations(maxsubiter)
        self.basis.track_bound = bool(lower_bound)
        w = None if weights is None else E.to_dev(np.asarray(weights, dtype=np.float64), E.tdtype(s.precision))
        resp = None
        if randomize:
            r0 = random_responsibilities(self.size, s.N)
            if w is None:
                s.stats_from_resp(r0)
            else:
                resp = E.to_dev(r0, E.tdtype(s.precision))
        else:
            if sample_likelihood:
                self.models._consume_reference_draws()
            if w is None:
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
                if sample_likelihood:
                    self.gating._store(outs['gating'], set_probs=False)
                    self.gating.likelihood.params = self.gating.posterior.rvs()
                    self.models._consume_reference_draws()
                if lower_bound:
                    vlb.append(float(lse.item()) + float(outs['gating']['vlb'].item())
                               + sum(float(o['vlb'].sum().item()) for o in outs['parts']))
                    if len(vlb) > 1 and abs(vlb[-1] - vlb[-2]) < tol:
                        break
                pbar.update(1)
        if outs is not None:
            s.store(outs, MEANFIELD, set_probs=False)
        return vlb

    def expected_log_complete_likelihood(self, x, y):
        self.models._consume_reference_draws()          # quirk q12: the reference's E-step samples and discards K experts
        return super().expected_log_complete_likelihood(x, y)

    def expected_responsibilities(self, x, y):
        self.models._consume_reference_draws()
        return super().expected_responsibilities(x, y)

    def expected_log_likelihood(self, x, y):
        """(N,): log sum_k exp(E log joint)   (hilr.py:151-153)."""
        self.models._consume_reference_draws()
        s = self._session(x, y)
        a = s.loglik(s.operands_from_posterior())
        return E.to_host(E.softmax(a, s.precision, lse=True)['lse']).astype(np.float64)

    def meanfield_update_parameters(self, x, y, resp, maxsubiter):
        self.meanfield_update_basis(x, resp, maxsubiter)
        self.meanfield_update_models(x, y, resp, maxsubiter)
        self.meanfield_update_gating(resp)

    def meanfield_update_basis(self, x, resp, maxsubiter):
        self.basis.meanfield_update(x, resp, maxsubiter)

    def meanfield_update_models(self, x, y, resp, maxsubiter):
        self.models.meanfield_update(x, y, resp, maxsubiter)

    # -- SVI (hilr.py:221-259): the experts' natural-gradient step is not implemented in the reference either -------
    def meanfield_stochastic_descent(self, x, y, randomize=True, weights=None, maxiter=250, maxsubiter=5, scale=1,
                                     step_size=1e-2, progress_bar=True, procces_id=0):
        xx, yy = self._scaled(x, y)
        resp = random_responsibilities(self.size, len(xx)) if randomize is True else self.expected_responsibilities(xx, yy)
        for _ in range(maxiter):
            resp = resp if weights is None else resp * weights
            self.meanfield_sgd_parameters(xx, yy, resp, maxsubiter, scale, step_size)
            resp = self.expected_responsibilities(xx, yy)
        return []

    def meanfield_sgd_parameters(self, x, y, resp, maxsubiter, scale, step_size):
        self.meanfield_sgd_basis(x, resp, maxsubiter, scale, step_size)
        self.meanfield_sgd_models(x, y, resp, maxsubiter, scale, step_size)
        self.meanfield_sgd_gating(resp, scale, step_size)

    def meanfield_sgd_basis(self, x, resp, maxsubiter, scale, step_size):
        self.basis.meanfield_sgd(x, resp, maxsubiter, scale, step_size)

    def meanfield_sgd_models(self, x, y, resp, maxsubiter, scale, step_size):
        self.models.meanfield_sgd(x, y, resp, maxsubiter, scale, step_size)

    def _lower_bound_at_posterior(self, xx, yy):
        s = self._session(xx, yy)
        s.sweep(s.operands_from_posterior(), hard=False)
        return float(self.gating.variational_lowerbound() + self.basis.variational_lowerbound()
                     + np.sum(self.models.variational_lowerbound()) + s.lse_sum.item())


# ---------------------------------------------------------------------------------------------------------------
from .hgmm import _log_weights, _softmax_rows  # noqa: E402
from .ilr import _Scaler  # noqa: E402
from ..utils.data import batches  # noqa: E402


class MixtureOfMixtureOfLinearGaussians:
    """hilr.py:18-76: `components` is a list of cluster_size MixtureOfLinearGaussians."""

    def __init__(self, cluster_size, mixture_size, input_dim, output_dim, gating, components, scale=False):
        self.cluster_size, self.mixture_size = cluster_size, mixture_size
        self.input_dim, self.output_dim = input_dim, output_dim
        self.gating, self.components = gating, components
        self.scale = scale
        self.input_transform, self.output_transform = _Scaler(), _Scaler()

    @property
    def params(self):
        raise NotImplementedError

    @property
    def nb_params(self):
        raise NotImplementedError

    def init_transform(self, x, y):
        self.scale = True
        self.input_transform.fit(x)
        self.output_transform.fit(y)

    def used_labels(self, x, y):
        raise NotImplementedError

    def rvs(self, size=1):
        raise NotImplementedError

    def log_complete_likelihood(self, x, y):
        comp = np.stack([c.log_likelihood(x, y) for c in self.components])
        return comp + self.gating.log_likelihood(np.arange(self.cluster_size))[:, None]

    def log_likelihood(self, x, y):
        import torch
        a = E.to_dev(self.log_complete_likelihood(x, y), torch.float64)
        return E.to_host(E.softmax(a, 'fp64', lse=True)['lse'])

    def responsibilities(self, x, y):
        return _softmax_rows(self.log_complete_likelihood(x, y))

    def max_likelihood(self, x, y, randomize=True, maxiter=250, maxsubiter=5, progress_bar=True, process_id=0):
        raise NotImplementedError


class BayesianMixtureOfMixtureOfLinearGaussians:
    """hilr.py:293-609: `components` is a list of cluster_size BayesianMixtureOfLinearGaussiansWithTiedActivation."""

    def __init__(self, cluster_size, mixture_size, input_dim, output_dim, gating, components, scale=False):
        self.cluster_size, self.mixture_size = cluster_size, mixture_size
        self.input_dim, self.output_dim = input_dim, output_dim
        self.gating, self.components = gating, components
        self.likelihood = MixtureOfMixtureOfLinearGaussians(cluster_size, mixture_size, input_dim, output_dim,
                                                            gating=gating.likelihood, components=[c.likelihood for c in components])
        self.scale = scale
        self.input_transform, self.output_transform = _Scaler(), _Scaler()

    def _scaled(self, x, y=None):
        x = np.reshape(x, (-1, self.input_dim))
        xx = self.input_transform.transform(x) if self.scale else np.asarray(x, dtype=np.float64)
        if y is None:
            return xx
        y = np.reshape(y, (-1, self.output_dim))
        return xx, (self.output_transform.transform(y) if self.scale else np.asarray(y, dtype=np.float64))

    def used_labels(self, x, y):
        z = np.argmax(self.expected_responsibilities(*self._scaled(x, y)), axis=0)
        return np.where(np.bincount(z, minlength=self.cluster_size) > 0)[0]

    def init_transform(self, x, y):
        self.scale = True
        self.input_transform.fit(x)
        self.output_transform.fit(y)

    def max_aposteriori(self, x, y, randomize=True, maxiter=250, maxsubiter=5, progress_bar=True, process_id=0):
        raise NotImplementedError

    # -- Gibbs (hilr.py:346-387) ---------------------------------------------------------------
    def resample(self, x, y, init_labels='prior', maxiter=250, maxsubiter=100, maxsubsubiter=5, progress_bar=True, process_id=0):
        xx, yy = self._scaled(x, y)
        if init_labels == 'random':
            z = npr.choice(self.cluster_size, size=(len(xx)))
        elif init_labels == 'posterior':
            _, z = self.resample_labels(xx, yy)
        elif init_labels == 'prior':
            z = self.gating.likelihood.rvs(len(xx))
        with tqdm(total=maxiter, desc=f'Init #{process_id + 1}', position=process_id, disable=not progress_bar) as pbar:
            for _ in range(maxiter):
                self.resample_components(xx, yy, z, maxsubiter, maxsubsubiter)
                self.resample_gating(z)
                _, z = self.resample_labels(xx, yy)
                pbar.update(1)
        self.labels_ = np.asarray(z, dtype=np.int32)

    def resample_labels(self, x, y):
        from ..utils.stats import sample_discrete_from_log
        log_prob = self.likelihood.log_complete_likelihood(x, y)
        return log_prob, sample_discrete_from_log(log_prob, axis=0, precision='fp64')

    def resample_gating(self, z):
        self.gating.resample(z)

    def resample_components(self, x, y, z, maxsubiter, maxsubsubiter):
        for m in range(self.cluster_size):
            idx = np.where(z == m)[0]
            self.components[m].resample(x=x[idx], y=y[idx], maxiter=maxsubiter, maxsubiter=maxsubsubiter, progress_bar=False)

    # -- mean field (hilr.py:389-458) -----------------------------------------------------------
    def expected_log_complete_likelihood(self, x, y):
        comp = np.stack([c.expected_log_likelihood(x, y) for c in self.components])
        return comp + _log_weights(self.gating)[:, None]

    def expected_responsibilities(self, x, y):
        return _softmax_rows(self.expected_log_complete_likelihood(x, y))

    def meanfield_coordinate_descent(self, x, y, randomize=True, maxiter=250, maxsubiter=5, maxsubsubiter=5, tol=1e-16,
                                     progress_bar=True, process_id=0):
        xx, yy = self._scaled(x, y)
        resp = random_responsibilities(self.cluster_size, len(xx)) if randomize else self.expected_responsibilities(xx, yy)
        with tqdm(total=maxiter, desc=f'VI #{process_id + 1}', position=process_id, disable=not progress_bar) as pbar:
            for i in range(maxiter):
                self.meanfield_update_parameters(xx, yy, resp, maxsubiter, maxsubsubiter, randomize if i == 0 else False)
                resp = self.expected_responsibilities(xx, yy)
                pbar.update(1)
        return []

    def meanfield_update_parameters(self, x, y, resp, maxsubiter, maxsubsubiter, randomize):
        self.meanfield_update_components(x, y, resp, maxsubiter, maxsubsubiter, randomize)
        self.meanfield_update_gating(resp)

    def meanfield_update_gating(self, resp):
        self.gating.meanfield_update(None, resp)

    def meanfield_update_components(self, x, y, resp, maxsubiter, maxsubsubiter, randomize):
        for m in range(self.cluster_size):
            self.components[m].meanfield_coordinate_descent(x=x, y=y, randomize=randomize, weights=resp[m, :], maxiter=maxsubiter,
                                                            maxsubiter=maxsubsubiter, progress_bar=False, sample_likelihood=True)

    # -- SVI (hilr.py:460-515; the experts' natural-gradient step is not implemented in the reference either) ---------
    def meanfield_stochastic_descent(self, x, y, randomize=True, maxiter=250, maxsubiter=5, maxsubsubiter=5, step_size=1e-2,
                                     batch_size=128, progress_bar=True, procces_id=0):
        xx, yy = self._scaled(x, y)
        scale = batch_size / float(len(xx))
        for i in range(maxiter):
            rnd = randomize if i == 0 else False
            for batch in batches(batch_size, len(xx)):
                resp = random_responsibilities(self.cluster_size, len(batch)) if rnd is True \
                    else self.expected_responsibilities(xx[batch, :], yy[batch, :])
                self.meanfield_sgd_parameters(xx[batch, :], yy[batch, :], resp, maxsubiter, maxsubsubiter, rnd, scale, step_size)
        return []

    def meanfield_sgd_parameters(self, x, y, resp, maxsubiter, maxsubsubiter, randomize, scale, step_size):
        self.meanfield_sgd_components(x, y, resp, maxsubiter, maxsubsubiter, randomize, scale, step_size)
        self.meanfield_sgd_gating(resp, scale, step_size)

    def meanfield_sgd_components(self, x, y, resp, maxsubiter, maxsubsubiter, randomize, scale, step_size):
        for m in range(self.cluster_size):
            self.components[m].meanfield_stochastic_descent(x=x, y=y, randomize=randomize, weights=resp[m, :], maxiter=maxsubiter,
                                                            maxsubiter=maxsubsubiter, scale=scale, step_size=step_size,
                                                            progress_bar=False)

    def meanfield_sgd_gating(self, resp, scale, step_size):
        self.gating.meanfield_sgd(None, resp, scale, step_size)

    def variational_lowerbound_labels(self, resp):
        raise NotImplementedError

    def variational_lowerbound_data(self, x, y, resp):
        raise NotImplementedError

    def variational_lowerbound(self, x, y, upper_resp, lower_resp):
        raise NotImplementedError

    # -- prediction (hilr.py:527-609) -------------------------------------------------------------
    def _log_weights_all(self, xx):
        """(M, K, N): log E[cluster weight] + log E[local weight] + log posterior-predictive input density."""
        lw = np.stack([np.log(c.gating.posterior.mean())[:, None] + c.basis.log_posterior_predictive_gaussian(xx)
                       for c in self.components])
        return lw + np.log(self.gating.posterior.mean())[:, None, None]

    def meanfield_predictive_activation(self, x):
        lw = self._log_weights_all(self._scaled(x))
        return _softmax_rows(lw.reshape(-1, lw.shape[-1])).reshape(lw.shape)

    def meanfield_predictive_weights(self, x):
        lw = self._log_weights_all(np.asarray(x, dtype=np.float64))
        return _softmax_rows(lw.reshape(-1, lw.shape[-1])).reshape(lw.shape)

    def meanfield_predictive_moments(self, x):
        pairs = [c.models.posterior_predictive_gaussian(x) for c in self.components]
        return np.stack([p[0] for p in pairs]), np.linalg.inv(np.stack([p[1] for p in pairs]))

    @staticmethod
    def mixture_moments(mus, covars, weights):
        mean = np.einsum('mknd,mkn->nd', mus, weights)
        covar = np.einsum('mkndl,mkn->ndl', covars + np.einsum('mknd,mknl->mkndl', mus, mus), weights) \
            - np.einsum('nd,nl->ndl', mean, mean)
        return mean, covar

    def meanfield_prediction(self, x, prediction='average', incremental=False, variance='diagonal'):
        x = np.reshape(x, (-1, self.input_dim))
        xx = self._scaled(x)
        weights = self.meanfield_predictive_weights(xx)
        mus, sigmas = self.meanfield_predictive_moments(xx)
        if prediction == 'mode':
            n = len(xx)
            mk = np.argmax(weights.reshape(-1, n), axis=0)
            mean = mus.reshape(-1, n, self.output_dim)[mk, np.arange(n)]
            covar = sigmas.reshape(-1, n, self.output_dim, self.output_dim)[mk, np.arange(n)]
        elif prediction == 'average':
            mean, covar = self.mixture_moments(mus, sigmas, weights)
        else:
            raise NotImplementedError
        if self.scale:
            mean = self.output_transform.inverse_transform(mean)
            mat = np.diag(np.sqrt(self.output_transform.var_))
            covar = np.einsum('kh,...hj,ji->...ki', mat, covar, mat.T)
        if incremental:
            mean += x[:, :self.output_dim]
        var = np.vstack(list(map(np.diag, covar)))
        return (mean, var, np.sqrt(var)) if variance == 'diagonal' else (mean, covar, np.sqrt(var))
