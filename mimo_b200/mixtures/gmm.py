"""Mixtures of Gaussians: EM, MAP-EM, Gibbs, mean-field VI and SVI drivers
(API of mimo/mixtures/gmm.py; the sweeps run on the GPU).

Every loop keeps the observations resident on the device for its whole duration and moves
only K-sized quantities (counts, lower-bound scalars, host-drawn variates) across the
boundary per sweep.  The (K, N) arrays of the reference (log-probabilities,
responsibilities, one-hot labels) are produced only by the methods whose return value they
are.
"""
import numpy as np
import numpy.random as npr
from tqdm import tqdm

from .. import _engine as E
from ..distributions.bayesian import (CategoricalWithDirichlet, CategoricalWithStickBreaking,  # noqa: F401
                                      MEANFIELD, GIBBS, MAP)
from ..distributions.gaussian import StackedGaussiansWithDiagonalPrecision
from ..distributions.bayesian import StackedGaussiansWithNormalGammas
from ..utils.data import batches
from ._driver import Session, Part, random_responsibilities


def _as_obs(obs):
    obs = np.asarray(obs, dtype=np.float64)
    if np.isnan(obs).any():
        raise ValueError('mimo_b200 sweep drivers need finite observations (NaN rows are only '
                         'supported by the per-object log_likelihood / statistics methods)')
    return obs


class _LikelihoodPart:
    """adapter: a bare likelihood object as a session part (EM has no priors)."""

    def __init__(self, lik):
        self.lik = lik
        self.diag = isinstance(lik, StackedGaussiansWithDiagonalPrecision)

    def _rows(self, mode):
        return self.lik.dim

    def _prior_dev(self):
        return None

    def _likelihood_operands(self, ops, layout):
        if self.diag:
            E.operands_gauss_diag(ops, E.to_dev(self.lik.mus), E.to_dev(self.lik.lmbdas_diags))
            return E.Info()
        return E.operands_gauss(ops, E.to_dev(self.lik.mus), E.to_dev(self.lik.lmbdas),
                                row_off=layout['row_off'], col_map=layout['col_map'])


class MixtureOfGaussians:
    """gating: Categorical; components: stacked Gaussian likelihoods (full or diagonal)."""

    def __init__(self, gating, components):
        assert components.size == gating.dim
        self.gating = gating
        self.components = components

    @property
    def params(self):
        raise NotImplementedError

    @property
    def nb_params(self):
        raise NotImplementedError

    @property
    def size(self):
        return self.gating.dim

    @property
    def dim(self):
        return self.components.dim

    def _session(self, obs, precision=None):
        part = _LikelihoodPart(self.components)
        idx = None if part.diag else E.identity_map(self.dim, self.dim)
        return Session(_as_obs(obs), self.size, None, [Part(part, idx, idx)], 'diag' if part.diag else 'quad',
                       precision or self.components.precision)

    def _log_probs(self):
        with np.errstate(divide='ignore'):
            return np.log(self.gating.probs)

    def used_labels(self, obs):
        labels = np.argmax(self.responsibilities(obs), axis=0)
        return np.where(np.bincount(labels, minlength=self.size) > 0)[0]

    def rvs(self, size=1):
        labels = self.gating.rvs(size)
        counts = np.bincount(labels, minlength=self.size)
        obs = np.zeros((size, self.dim))
        for idx, (c, count) in enumerate(zip(self.components.dists, counts)):
            obs[labels == idx, ...] = np.reshape(c.rvs(int(count)), (-1, self.dim)) if count > 0 else obs[labels == idx]
        perm = npr.permutation(size)
        return obs[perm], labels[perm]

    def log_complete_likelihood(self, obs):
        """(K, N): component log-likelihood + log gating probability (gmm.py:67-70)."""
        s = self._session(obs)
        return E.to_host(s.loglik(s.operands_from_likelihood(self._log_probs()))).astype(np.float64)

    def log_likelihood(self, obs):
        s = self._session(obs)
        a = s.loglik(s.operands_from_likelihood(self._log_probs()))
        return E.to_host(E.softmax(a, s.precision, lse=True)['lse']).astype(np.float64)

    def responsibilities(self, obs):
        s = self._session(obs)
        a = s.loglik(s.operands_from_likelihood(self._log_probs()))
        E.softmax(a, s.precision, resp=True)
        return E.to_host(a).astype(np.float64)

    def max_likelihood(self, obs, randomize=True, weights=None, maxiter=250, progress_bar=True, process_id=0):
        """EM (gmm.py:77-103).  Per iteration: M-step kernels on the packed statistics, then
        one fused E-step + statistics sweep whose log-normaliser sum is the log-likelihood."""
        s = self._session(obs)
        part = s.parts[0].w
        if weights is not None:
            return self._max_likelihood_weighted(obs, randomize, weights, maxiter, progress_bar, process_id)
        if randomize:
            s.stats_from_resp(random_responsibilities(self.size, s.N))
        else:
            s.sweep(s.operands_from_likelihood(self._log_probs()), hard=False)
        log_lik = []
        with tqdm(total=maxiter, desc=f'EM #{process_id + 1}', position=process_id, disable=not progress_bar) as pbar:
            for _ in range(maxiter):
                self._mstep_from_stats(s, part)
                s.sweep(s.operands_from_likelihood(self._log_probs()), hard=False)
                log_lik.append(float(s.lse_sum.item()))
                pbar.update(1)
        return log_lik

    def _mstep_from_stats(self, s, part):
        counts = s.counts_host()
        if part.diag:
            mu, lam = E.mstep_gauss_diag(s.stat, s.F, self.size, self.dim, tied=self.components._tied)
            self.components.params = (E.to_host(mu), E.to_host(lam))
        else:
            idx = s.parts[0].stat_idx
            mu, lmbda, info = E.mstep_gauss(s.stat, s.F, idx, self.dim + 1, self.size, self.dim,
                                            tied=self.components._tied)
            try:
                info.check()
            except np.linalg.LinAlgError as e:
                raise AssertionError(str(e))
            self.components.params = (E.to_host(mu), E.to_host(lmbda))
        self.gating.probs = counts / counts.sum()

    def _max_likelihood_weighted(self, obs, randomize, weights, maxiter, progress_bar, process_id):
        resp = random_responsibilities(self.size, len(obs)) if randomize else self.responsibilities(obs)
        log_lik = []
        with tqdm(total=maxiter, desc=f'EM #{process_id + 1}', position=process_id, disable=not progress_bar) as pbar:
            for _ in range(maxiter):
                resp = resp * weights
                self.components.max_likelihood(obs, resp)
                self.gating.max_likelihood(None, resp)
                resp = self.responsibilities(obs)
                log_lik.append(np.sum(self.log_likelihood(obs)))
                pbar.update(1)
        return log_lik

    def plot(self, *args, **kwargs):
        raise NotImplementedError('plotting is outside the scope of mimo_b200')


class BayesianMixtureOfGaussians:
    """gating: CategoricalWith{Dirichlet,StickBreaking}; components: stacked Gaussians with
    Normal-Wishart or Normal-Gamma priors."""

    def __init__(self, gating, components, precision=None):
        self.gating = gating
        self.components = components
        self.precision = precision
        self.likelihood = MixtureOfGaussians(gating=self.gating.likelihood, components=self.components.likelihood)
        self.labels_ = None

    @property
    def size(self):
        return self.likelihood.size

    @property
    def dim(self):
        return self.likelihood.dim

    def _family(self):
        return 'diag' if isinstance(self.components, StackedGaussiansWithNormalGammas) else 'quad'

    def _session(self, obs, comm=None):
        fam = self._family()
        idx = E.identity_map(self.dim, self.dim) if fam == 'quad' else None
        return Session(_as_obs(obs) if isinstance(obs, (np.ndarray, list)) else obs, self.size,
                       self.gating, [Part(self.components, idx, idx)], fam,
                       self.precision or self.components.likelihood.precision, comm=comm)

    def used_labels(self, obs):
        labels = np.argmax(self.expected_responsibilities(obs), axis=0)
        return np.where(np.bincount(labels, minlength=self.size) > 0)[0]

    # -- MAP-EM ------------------------------------------------------------------------------
    def max_aposteriori(self, obs, randomize=True, maxiter=250, progress_bar=True, process_id=0):
        """gmm.py:176-204."""
        s = self._session(obs)
        if randomize:
            s.stats_from_resp(random_responsibilities(self.size, s.N))
        else:
            s.sweep(s.operands_from_likelihood(self.likelihood._log_probs()), hard=False)
        log_prob = []
        with tqdm(total=maxiter, desc=f'MAP #{process_id + 1}', position=process_id, disable=not progress_bar) as pbar:
            for _ in range(maxiter):
                ops, outs = s.update_from_stats(MAP, want_lik=True)
                s.check(outs)
                s.store(outs, MAP)
                s.sweep(ops, hard=False)
                log_lik = float(s.lse_sum.item())
                log_prior = self.gating.prior.log_likelihood(self.gating.likelihood.params) \
                    + np.sum(self.components.prior.log_likelihood(self.components.likelihood.params))
                log_prob.append(log_lik + log_prior)
                pbar.update(1)
        return log_prob

    # -- Gibbs -------------------------------------------------------------------------------
    def resample(self, obs, init_labels='prior', maxiter=1, progress_bar=True, process_id=0, comm=None,
                 label_rng='numpy', param_rng='numpy'):
        """gmm.py:207-225.  Random variates come from the global numpy.random stream in the
        reference's order (components per k, gating, one uniform per point), so a seeded run
        reproduces the reference's chain; all arithmetic on them is on the device.
        Sharded (comm): rank 0's stream is the chain's stream -- parameter variates and label uniforms are drawn
        there and broadcast, so an identically seeded 1-rank run gives the same chain.
        label_rng='philox': label uniforms from the kernel's counter-based generator keyed by the global point
        index (no N host variates per sweep; not the reference's stream).
        param_rng='device': the parameter variates from a device generator seeded once from numpy.random (no host
        read of the statistics and no broadcast per sweep; not the reference's stream)."""
        s = self._session(obs, comm)
        if param_rng == 'device':
            s.seed_parameters(s.host_draw(lambda: int(npr.randint(1 << 30))))
        lo = comm.point_offset if comm is not None else 0
        n_glob = comm.N_global if (comm is not None and comm.N_global is not None) else s.N
        if init_labels == 'random':
            labels = s.host_draw(lambda: npr.choice(self.size, size=(n_glob,)))[lo:lo + s.N]
        elif init_labels == 'prior':
            labels = s.host_draw(lambda: self.gating.likelihood.rvs(n_glob))[lo:lo + s.N]
        elif init_labels == 'posterior':
            ops = s.operands_from_likelihood(self.likelihood._log_probs())
            u, seed = s.label_uniforms(label_rng)
            labels = s.sweep(ops, hard=True, uniforms=u, seed=seed).labels
        if init_labels != 'posterior':
            s.stats_from_labels(labels)
        with tqdm(total=maxiter, desc=f'Init #{process_id + 1}', position=process_id, disable=not progress_bar) as pbar:
            for _ in range(maxiter):
                var, gvar = s.draw_gibbs_variates(param_rng)
                ops, outs = s.update_from_stats(GIBBS, variates=var, gating_variates=gvar, want_lik=True)
                s.check(outs)
                u, seed = s.label_uniforms(label_rng)
                buf = s.sweep(ops, hard=True, uniforms=u, seed=seed)
                pbar.update(1)
        if maxiter > 0:
            s.store(outs, GIBBS)
            self.labels_ = E.to_host(buf.labels)

    def resample_labels(self, obs):
        log_prob = self.likelihood.log_complete_likelihood(obs)
        from ..utils.stats import sample_discrete_from_log
        labels = sample_discrete_from_log(log_prob, axis=0,
                                          precision=self.precision or self.components.likelihood.precision or E.default_precision())
        return log_prob, labels

    def resample_gating(self, labels):
        self.gating.resample(labels)

    def resample_components(self, obs, labels):
        s = self._session(obs)
        s.stats_from_labels(labels)
        counts = s.counts_host()
        stat_host = E.to_host(s.stat) if s.family == 'diag' else None
        v = self.components._draw_variates(counts, stat_host) if s.family == 'diag' else self.components._draw_variates(counts)
        out = self.components._update(s.stat, s.F, s.parts[0].layout(s.D + 1, 0), GIBBS, variates=v, want_lik=True)
        out['info'].check()
        self.components._store(out, GIBBS)

    # -- mean field --------------------------------------------------------------------------
    def expected_log_complete_likelihood(self, obs):
        """(K, N): E_q[log N] + E_q[log pi] (gmm.py:244-254)."""
        s = self._session(obs)
        return E.to_host(s.loglik(s.operands_from_posterior())).astype(np.float64)

    def expected_log_likelihood(self, obs):
        s = self._session(obs)
        a = s.loglik(s.operands_from_posterior())
        return E.to_host(E.softmax(a, s.precision, lse=True)['lse']).astype(np.float64)

    def expected_responsibilities(self, obs):
        s = self._session(obs)
        a = s.loglik(s.operands_from_posterior())
        E.softmax(a, s.precision, resp=True)
        return E.to_host(a).astype(np.float64)

    def meanfield_coordinate_descent(self, obs, randomize=True, maxiter=250, tol=1e-8,
                                     progress_bar=True, process_id=0, comm=None, rtol=0., sample_likelihood=False,
                                     graph=False):
        """gmm.py:261-287.  Per iteration: batched posterior kernels (statistics -> posterior,
        operands, lower-bound terms), then ONE fused E-step + statistics sweep.
        randomize=True draws the reference's npr.rand(K, N) on the host (seeded runs replay the reference);
        randomize='device' draws the random responsibilities on the device in point chunks, for N where a (K, N)
        host array cannot exist.  `obs` may be a resident device tensor (N, d) of the model's precision.
        tol is the reference's ABSOLUTE early-stop threshold; rtol adds a relative one (|delta| < rtol |vlb|): the
        FP64 statistics are accumulated with atomics in a run-dependent order, so at N ~ 1e7+ the bound jitters by
        ~1e-9 relative from sweep to sweep and an absolute 1e-8 can never fire.
        sample_likelihood=True also performs the reference's per-iteration likelihood.params = posterior.rvs()
        draws (components, then gating) so the numpy.random stream advances as in the reference.
        graph=True replays iterations 2.. from a CUDA graph of the first one (single process only): the same kernels on
        the same buffers, without their launch latencies -- what bounds an iteration at the reference's own example sizes."""
        s = self._session(obs, comm)
        if randomize == 'device':
            s.stats_from_random_resp(seed=s.host_draw(lambda: int(npr.randint(1 << 30))))
        elif randomize:
            s.stats_from_resp(random_responsibilities(self.size, s.N))
        else:
            s.sweep(s.operands_from_posterior(), hard=False)
        vlb = []
        outs = None
        with tqdm(total=maxiter, desc=f'VI #{process_id + 1}', position=process_id, disable=not progress_bar) as pbar:
            cuda_graph = None
            for it in range(maxiter):
                if graph and comm is None and it == 1:
                    cuda_graph, outs = s.capture_meanfield_step()
                if cuda_graph is not None:
                    cuda_graph.replay()
                else:
                    ops, outs = s.update_from_stats(MEANFIELD)
                    s.sweep(ops, hard=False)
                s.check(outs)
                vlb.append(s.lower_bound(outs))
                if sample_likelihood:
                    s.store(outs, MEANFIELD, set_probs=False)
                    self.components.likelihood.params = self.components.posterior.rvs()
                    self.gating.likelihood.params = self.gating.posterior.rvs()
                if len(vlb) > 1 and (abs(vlb[-1] - vlb[-2]) < tol or abs(vlb[-1] - vlb[-2]) < rtol * abs(vlb[-1])):
                    break
                pbar.update(1)
        if outs is not None:
            s.store(outs, MEANFIELD, set_probs=False)
        return vlb

    def meanfield_update_parameters(self, obs, resp):
        self.meanfield_update_components(obs, resp)
        self.meanfield_update_gating(resp)

    def meanfield_update_gating(self, resp):
        self.gating.meanfield_update(None, resp)

    def meanfield_update_components(self, obs, resp):
        self.components.meanfield_update(obs, resp)

    # -- SVI ---------------------------------------------------------------------------------
    def meanfield_stochastic_descent(self, obs, randomize=True, maxiter=500, step_size=1e-2,
                                     batch_size=128, progress_bar=True, procces_id=0, device=False, graph=False,
                                     lower_bound_every=1):
        """gmm.py:300-326.  device=True keeps the whole loop on the device (mixtures/_svi.py: resident data, minibatch
        gather, natural-parameter blend folded into the conjugate-update kernel, lower bounds read once at the end;
        Normal-Wishart components); graph=True also replays the iterations from a CUDA graph; lower_bound_every=k
        evaluates the full-data bound (the reference does it after EVERY minibatch, which costs a full sweep) only
        every k-th iteration.  The device path does not perform the reference's trailing likelihood.params =
        posterior.rvs() draws: they never feed back into the iterations."""
        if device:
            from . import _svi
            with tqdm(total=maxiter, desc=f'SVI #{procces_id + 1}', position=procces_id, disable=not progress_bar) as pbar:
                return _svi.run(self, self._session(obs), randomize, maxiter, step_size, batch_size, graph, lower_bound_every,
                                batches, random_responsibilities, pbar)
        obs = _as_obs(obs)
        vlb = []
        with tqdm(total=maxiter, desc=f'SVI #{procces_id + 1}', position=procces_id, disable=not progress_bar) as pbar:
            scale = batch_size / float(len(obs))
            for i in range(maxiter):
                for batch in batches(batch_size, len(obs)):
                    if i == 0 and randomize is True:
                        resp = random_responsibilities(self.size, len(batch))
                    else:
                        resp = self.expected_responsibilities(obs[batch, :])
                    self.meanfield_sgd_parameters(obs[batch, :], resp, scale, step_size)
                vlb.append(self._lower_bound_at_posterior(obs))
                pbar.update(1)
        return vlb

    def meanfield_sgd_parameters(self, obs, resp, scale, step_size):
        self.meanfield_sgd_components(obs, resp, scale, step_size)
        self.meanfield_sgd_gating(resp, scale, step_size)

    def meanfield_sgd_components(self, obs, resp, scale, step_size):
        self.components.meanfield_sgd(obs, resp, scale, step_size)

    def meanfield_sgd_gating(self, resp, scale, step_size):
        self.gating.meanfield_sgd(None, resp, scale, step_size)

    # -- lower bound -------------------------------------------------------------------------
    def _lower_bound_at_posterior(self, obs):
        """lower bound with responsibilities = E-step of the current posterior."""
        s = self._session(obs)
        s.sweep(s.operands_from_posterior(), hard=False)
        return float(self.gating.variational_lowerbound() + np.sum(self.components.variational_lowerbound())
                     + s.lse_sum.item())

    def variational_lowerbound_obs(self, obs, resp):
        return np.sum(resp * self.components.expected_log_likelihood(obs))

    def variational_lowerbound_labels(self, resp):
        vlb = 0.
        if isinstance(self.gating, CategoricalWithDirichlet):
            vlb += np.sum(resp * np.expand_dims(self.gating.expected_log_likelihood(), axis=1))
        else:
            acc = np.vstack((np.cumsum(resp[::-1, :], axis=0)[-2::-1, :], np.zeros((1, resp.shape[-1]))))
            e_stick, e_rest = self.gating.expected_log_likelihood()
            vlb += np.sum(resp * e_stick[:, None] + acc * e_rest[:, None])
        with np.errstate(invalid='ignore', divide='ignore'):
            vlb -= np.nansum(resp * np.log(resp))
        return vlb

    def variational_lowerbound(self, obs, resp):
        """gmm.py:358-364 for arbitrary responsibilities."""
        return self.gating.variational_lowerbound() + np.sum(self.components.variational_lowerbound()) \
            + self.variational_lowerbound_obs(obs, resp) + self.variational_lowerbound_labels(resp)

    def plot(self, *args, **kwargs):
        raise NotImplementedError('plotting is outside the scope of mimo_b200')
