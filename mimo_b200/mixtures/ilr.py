"""Mixtures of linear-Gaussian experts ("infinite local regression"): Gibbs, mean-field VI,
SVI drivers and the predictive path (API of mimo/mixtures/ilr.py; sweeps on the GPU).

A point is z = [x | y]; the input density ("basis", Normal-Wishart on x), the expert
("models", Matrix-Normal-Wishart on (x, y)) and the gating write their whitening rows into
ONE operand block per component, so basis + models + gating (ilr.py:71-75, 178-189) is a
single pass of the same fused sweep the GMM uses, and their sufficient statistics are
sub-blocks of one packed second-moment matrix of [x | y | 1]."""
import numpy as np
import numpy.random as npr
from scipy.special import logsumexp, gammaln
from tqdm import tqdm

from .. import _engine as E
from ..distributions.bayesian import CategoricalWithDirichlet, MEANFIELD, GIBBS
from ..utils.data import batches
from ._driver import Session, Part, random_responsibilities

eps = np.finfo(np.float64).tiny


class _Scaler:
    """mean / standard-deviation scaling (what sklearn's StandardScaler does at ilr.py:108-127)."""

    def __init__(self):
        self.mean_ = self.scale_ = self.var_ = None

    def fit(self, a):
        a = np.asarray(a, dtype=np.float64)
        self.mean_ = a.mean(axis=0)
        self.var_ = a.var(axis=0)
        self.scale_ = np.sqrt(self.var_)
        self.scale_[self.scale_ == 0.] = 1.
        return self

    def transform(self, a):
        return (np.asarray(a, dtype=np.float64) - self.mean_) / self.scale_

    def inverse_transform(self, a):
        return np.asarray(a) * self.scale_ + self.mean_


class _LikPart:
    """adapter: bare likelihoods of the non-Bayesian mixture as session parts."""

    def __init__(self, lik, rows, kind):
        self.lik, self.rows, self.kind = lik, rows, kind

    def _rows(self, mode):
        return self.rows

    def _prior_dev(self):
        return None

    def _likelihood_operands(self, ops, layout):
        if self.kind == 'basis':
            return E.operands_gauss(ops, E.to_dev(self.lik.mus), E.to_dev(self.lik.lmbdas),
                                    row_off=layout['row_off'], col_map=layout['col_map'])
        As = self.lik._affine_As() if hasattr(self.lik, '_affine_As') else self.lik.As      # slope and offset kept apart (hilr)
        return E.operands_lingauss(ops, E.to_dev(As), E.to_dev(self.lik.lmbdas),
                                   layout['row_off'], layout['col_map'])


def _stack_xy(x, y):
    z = np.hstack((np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)))
    if np.isnan(z).any():
        raise ValueError('mimo_b200 sweep drivers need finite inputs')
    return z


class MixtureOfLinearGaussians:

    def __init__(self, size, input_dim, output_dim, gating, basis, models):
        self.size = size
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.gating = gating
        self.basis = basis      # input density
        self.models = models    # output density

    @property
    def params(self):
        raise NotImplementedError

    @property
    def nb_params(self):
        raise NotImplementedError

    def _session(self, x, y, precision=None):
        lay = self.models.layout.dev()
        parts = [Part(_LikPart(self.basis, self.input_dim, 'basis'), lay['basis_idx'], lay['basis_idx']),
                 Part(_LikPart(self.models, self.output_dim, 'models'), lay['stat_idx'], lay['col_map'])]
        return Session(_stack_xy(x, y), self.size, None, parts, 'quad', precision or self.models.precision)

    def _log_probs(self):
        with np.errstate(divide='ignore'):
            return np.log(self.gating.probs)

    def used_labels(self, x, y):
        labels = np.argmax(self.responsibilities(x, y), axis=0)
        return np.where(np.bincount(labels, minlength=self.size) > 0)[0]

    def rvs(self, size=1):
        z = self.gating.rvs(size)
        counts = np.bincount(z, minlength=self.size)
        x = np.empty((size, self.input_dim))
        y = np.empty((size, self.output_dim))
        for idx, (b, m, count) in enumerate(zip(self.basis.dists, self.models.dists, counts)):
            if count > 0:
                x[z == idx, ...] = np.reshape(b.rvs(int(count)), (-1, self.input_dim))
                y[z == idx, ...] = np.reshape(m.rvs(x[z == idx, ...]), (-1, self.output_dim))
        perm = npr.permutation(size)
        return x[perm], y[perm], z[perm]

    def log_complete_likelihood(self, x, y):
        s = self._session(x, y)
        return E.to_host(s.loglik(s.operands_from_likelihood(self._log_probs()))).astype(np.float64)

    def log_likelihood(self, x, y):
        return logsumexp(self.log_complete_likelihood(x, y), axis=0)

    def responsibilities(self, x, y):
        s = self._session(x, y)
        a = s.loglik(s.operands_from_likelihood(self._log_probs()))
        E.softmax(a, s.precision, resp=True)
        return E.to_host(a).astype(np.float64)

    def max_likelihood(self, x, y, randomize=True, maxiter=250, progress_bar=True, process_id=0):
        raise NotImplementedError


class BayesianMixtureOfLinearGaussians:

    def __init__(self, size, input_dim, output_dim, gating, basis, models, scale=False, precision=None):
        self.size = size
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.gating = gating
        self.basis = basis      # input density
        self.models = models    # output density
        self.precision = precision
        self.likelihood = MixtureOfLinearGaussians(size, input_dim, output_dim, gating=self.gating.likelihood,
                                                   basis=self.basis.likelihood, models=self.models.likelihood)
        self.scale = scale
        self.input_transform = _Scaler()
        self.output_transform = _Scaler()
        self.labels_ = None

    def _scaled(self, x, y=None):
        x = np.reshape(x, (-1, self.input_dim))
        xx = self.input_transform.transform(x) if self.scale else np.asarray(x, dtype=np.float64)
        if y is None:
            return xx
        y = np.reshape(y, (-1, self.output_dim))
        return xx, (self.output_transform.transform(y) if self.scale else np.asarray(y, dtype=np.float64))

    def _session(self, xx, yy, comm=None):
        lay = self.models.layout.dev()
        parts = [Part(self.basis, lay['basis_idx'], lay['basis_idx']),
                 Part(self.models, lay['stat_idx'], lay['col_map'])]
        return Session(_stack_xy(xx, yy), self.size, self.gating, parts, 'quad',
                       self.precision or self.models.likelihood.precision, comm=comm)

    def used_labels(self, x, y):
        xx, yy = self._scaled(x, y)
        z = np.argmax(self.expected_responsibilities(xx, yy), axis=0)
        return np.where(np.bincount(z, minlength=self.size) > 0)[0]

    def init_transform(self, x, y):
        self.scale = True
        self.input_transform.fit(x)
        self.output_transform.fit(y)

    def max_aposteriori(self, x, y, randomize=True, maxiter=250, progress_bar=True, process_id=0):
        raise NotImplementedError

    # -- Gibbs -------------------------------------------------------------------------------
    def resample(self, x, y, init_labels='prior', maxiter=1, progress_bar=True, process_id=0):
        """ilr.py:134-159; variates in the reference's order: basis, models, gating, labels."""
        xx, yy = self._scaled(x, y)
        s = self._session(xx, yy)
        if init_labels == 'random':
            z = npr.choice(self.size, size=(s.N,))
        elif init_labels == 'prior':
            z = self.gating.likelihood.rvs(s.N)
        elif init_labels == 'posterior':
            ops = s.operands_from_likelihood(self.likelihood._log_probs())
            s.sweep(ops, hard=True, uniforms=npr.random(size=(1, s.N)))
        if init_labels != 'posterior':
            s.stats_from_labels(z)
        with tqdm(total=maxiter, desc=f'Init #{process_id + 1}', position=process_id, disable=not progress_bar) as pbar:
            for _ in range(maxiter):
                var, gvar = s.draw_gibbs_variates()
                ops, outs = s.update_from_stats(GIBBS, variates=var, gating_variates=gvar, want_lik=True)
                s.check(outs)
                buf = s.sweep(ops, hard=True, uniforms=npr.random(size=(1, s.N)))
                pbar.update(1)
        if maxiter > 0:
            s.store(outs, GIBBS)
            self.labels_ = E.to_host(buf.labels)

    def resample_labels(self, x, y):
        from ..utils.stats import sample_discrete_from_log
        log_prob = self.likelihood.log_complete_likelihood(x, y)
        labels = sample_discrete_from_log(log_prob, axis=0, precision=self.precision or E.default_precision())
        return log_prob, labels

    def resample_gating(self, z):
        self.gating.resample(z)

    def resample_basis(self, x, z):
        from ..utils.data import one_hot
        self.basis.resample(x, one_hot(z, K=self.size))

    def resample_models(self, x, y, z):
        from ..utils.data import one_hot
        self.models.resample(x, y, one_hot(z, K=self.size))

    # -- mean field --------------------------------------------------------------------------
    def expected_log_complete_likelihood(self, x, y):
        s = self._session(x, y)
        return E.to_host(s.loglik(s.operands_from_posterior())).astype(np.float64)

    def expected_responsibilities(self, x, y):
        s = self._session(x, y)
        a = s.loglik(s.operands_from_posterior())
        E.softmax(a, s.precision, resp=True)
        return E.to_host(a).astype(np.float64)

    def meanfield_coordinate_descent(self, x, y, randomize=True, maxiter=250, tol=1e-8,
                                     progress_bar=True, process_id=0, comm=None, rtol=0.):
        """ilr.py:196-228.  randomize='device' / rtol: see BayesianMixtureOfGaussians.meanfield_coordinate_descent."""
        xx, yy = self._scaled(x, y)
        s = self._session(xx, yy, comm)
        if randomize == 'device':
            s.stats_from_random_resp(seed=s.host_draw(lambda: int(npr.randint(1 << 30))))
        elif randomize:
            s.stats_from_resp(random_responsibilities(self.size, s.N))
        else:
            s.sweep(s.operands_from_posterior(), hard=False)
        vlb = []
        outs = None
        with tqdm(total=maxiter, desc=f'VI #{process_id + 1}', position=process_id, disable=not progress_bar) as pbar:
            for _ in range(maxiter):
                ops, outs = s.update_from_stats(MEANFIELD)
                s.sweep(ops, hard=False)
                s.check(outs)
                vlb.append(s.lower_bound(outs))
                if len(vlb) > 1 and (abs(vlb[-1] - vlb[-2]) < tol or abs(vlb[-1] - vlb[-2]) < rtol * abs(vlb[-1])):
                    break
                pbar.update(1)
        if outs is not None:
            s.store(outs, MEANFIELD, set_probs=False)
        return vlb

    def meanfield_update_parameters(self, x, y, resp):
        self.meanfield_update_basis(x, resp)
        self.meanfield_update_models(x, y, resp)
        self.meanfield_update_gating(resp)

    def meanfield_update_gating(self, resp):
        self.gating.meanfield_update(None, resp)

    def meanfield_update_basis(self, x, resp):
        self.basis.meanfield_update(x, resp)

    def meanfield_update_models(self, x, y, resp):
        self.models.meanfield_update(x, y, resp)

    # -- SVI ---------------------------------------------------------------------------------
    def meanfield_stochastic_descent(self, x, y, randomize=True, maxiter=500, step_size=1e-3,
                                     batch_size=128, progress_bar=True, procces_id=0, device=False, graph=False,
                                     lower_bound_every=1):
        """ilr.py:245-278.  device / graph / lower_bound_every: the device-resident route of mixtures/_svi.py, see
        BayesianMixtureOfGaussians.meanfield_stochastic_descent."""
        xx, yy = self._scaled(x, y)
        if device:
            from . import _svi
            with tqdm(total=maxiter, desc=f'SVI #{procces_id + 1}', position=procces_id, disable=not progress_bar) as pbar:
                return _svi.run(self, self._session(xx, yy), randomize, maxiter, step_size, batch_size, graph, lower_bound_every,
                                batches, random_responsibilities, pbar)
        vlb = []
        with tqdm(total=maxiter, desc=f'SVI #{procces_id + 1}', position=procces_id, disable=not progress_bar) as pbar:
            scale = batch_size / float(len(xx))
            for i in range(maxiter):
                for batch in batches(batch_size, len(xx)):
                    if i == 0 and randomize is True:
                        resp = random_responsibilities(self.size, len(batch))
                    else:
                        resp = self.expected_responsibilities(xx[batch, :], yy[batch, :])
                    self.meanfield_sgd_parameters(xx[batch, :], yy[batch, :], resp, scale, step_size)
                vlb.append(self._lower_bound_at_posterior(xx, yy))
                pbar.update(1)
        return vlb

    def meanfield_sgd_parameters(self, x, y, resp, scale, step_size):
        self.meanfield_sgd_basis(x, resp, scale, step_size)
        self.meanfield_sgd_models(x, y, resp, scale, step_size)
        self.meanfield_sgd_gating(resp, scale, step_size)

    def meanfield_sgd_gating(self, resp, scale, step_size):
        self.gating.meanfield_sgd(None, resp, scale, step_size)

    def meanfield_sgd_basis(self, x, resp, scale, step_size):
        self.basis.meanfield_sgd(x, resp, scale, step_size)

    def meanfield_sgd_models(self, x, y, resp, scale, step_size):
        self.models.meanfield_sgd(x, y, resp, scale, step_size)

    # -- lower bound -------------------------------------------------------------------------
    def _lower_bound_at_posterior(self, xx, yy):
        s = self._session(xx, yy)
        s.sweep(s.operands_from_posterior(), hard=False)
        return float(self.gating.variational_lowerbound() + np.sum(self.basis.variational_lowerbound())
                     + np.sum(self.models.variational_lowerbound()) + s.lse_sum.item())

    def variational_lowerbound_data(self, x, y, resp):
        return np.sum(resp * self.basis.expected_log_likelihood(x)) \
            + np.sum(resp * self.models.expected_log_likelihood(x, y))

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

    def variational_lowerbound(self, x, y, resp):
        return self.gating.variational_lowerbound() + np.sum(self.basis.variational_lowerbound()) \
            + np.sum(self.models.variational_lowerbound()) + self.variational_lowerbound_data(x, y, resp) \
            + self.variational_lowerbound_labels(resp)

    # -- prediction (ilr.py:325-430) -----------------------------------------------------------
    def _predictive_weights_dev(self, x, dist):
        """softmax_k( log E[pi_k] + log p(x; posterior-predictive basis_k) ) as a (K, N) device tensor: the Gaussian form
        through the E-step kernel, the Student-t form (utils/stats.py:53-79) by a device transform of its output."""
        if dist not in ('gaussian', 'studentt'):
            raise NotImplementedError(dist)
        mus, lmbdas = self.basis.posterior_predictive_gaussian()
        precision = self.precision or E.default_precision()
        d = self.input_dim
        ops = E.QuadOperands(self.size, d, d, precision)
        log_gating = np.log(self.gating.posterior.mean())
        E.set_log_weights(ops, log_gating if dist == 'gaussian' else np.zeros(self.size))
        E.operands_gauss(ops, E.to_dev(mus), E.to_dev(lmbdas)).check()
        a = E.loglik(E.to_dev(np.reshape(x, (-1, d)), E.tdtype(precision)), ops)
        if dist == 'studentt':
            dfs = self.basis.posterior_predictive_studentt()[2]
            half_logdet = 0.5 * np.linalg.slogdet(lmbdas)[1]
            c0 = half_logdet - 0.5 * d * np.log(2. * np.pi)                 # constant of the Gaussian form the kernel produced
            aux = gammaln((dfs + d) / 2.) - gammaln(dfs / 2.) + half_logdet - (d / 2.) * np.log(dfs * np.pi) - 0.5 * (dfs + d)
            E.studentt_from_quad(a, precision, c0, log_gating + aux, dfs)
        E.softmax(a, precision, resp=True)
        return a

    def meanfield_predictive_weights(self, x, dist='gaussian'):
        return E.to_host(self._predictive_weights_dev(x, dist)).astype(np.float64)

    def meanfield_predictive_activation(self, x, dist='gaussian'):
        return self.meanfield_predictive_weights(self._scaled(x), dist)

    def meanfield_predictive_moments(self, x, dist='gaussian'):
        """Per-expert moments (K, N, o), (K, N, o, o) as the reference returns them (API use at plotting sizes; the
        prediction itself combines them on the device without building these arrays)."""
        if dist == 'gaussian':
            mus, lmbdas = self.models.posterior_predictive_gaussian(x)
            return mus, np.linalg.inv(lmbdas)
        if dist != 'studentt':
            raise NotImplementedError(dist)
        mus, lmbdas, dfs = self.models.posterior_predictive_studentt(x)
        dfs = np.broadcast_to(dfs, (self.size,))
        return mus, np.einsum('kndl,k->kndl', np.linalg.inv(lmbdas), dfs / (dfs - 2))

    def meanfiled_log_predictive_likelihood(self, x, y, dist='gaussian'):
        mus, lmbdas = self.models.posterior_predictive_gaussian(x)[:2]
        diff = np.reshape(y, (-1, self.output_dim))[None, :, :] - mus
        delta = np.einsum('knd,kndl,knl->kn', diff, lmbdas, diff)
        if dist == 'gaussian':
            return -0.5 * delta + 0.5 * np.linalg.slogdet(lmbdas)[1] - 0.5 * self.output_dim * np.log(2. * np.pi)
        dfs = np.broadcast_to(self.models.posterior_predictive_studentt(x)[2], (self.size,))[:, None]
        o = self.output_dim
        aux = gammaln((dfs + o) / 2.) - gammaln(dfs / 2.) + 0.5 * np.linalg.slogdet(lmbdas)[1] \
            - (o / 2.) * np.log(dfs * np.pi) - 0.5 * (dfs + o)
        return aux + np.log1p(delta / dfs)

    @staticmethod
    def mixture_moments(mus, covars, weights):
        mu = np.einsum('knd,kn->nd', mus, weights)
        covar = np.einsum('kndl,kn->ndl', covars + np.einsum('knd,knl->kndl', mus, mus), weights) \
            - np.einsum('nd,nl->ndl', mu, mu)
        return mu, covar

    def meanfield_prediction(self, x, y=None, prediction='average', dist='gaussian',
                             incremental=False, variance='diagonal'):
        """ilr.py:384-430.  Weights (E-step kernel + softmax), expert moments, their mixture / mode and the negative log
        predictive density all stay on the device (mimo_predict_lingauss); only (N, o)-sized results come back."""
        if prediction not in ('mode', 'average'):
            raise NotImplementedError
        x = np.reshape(x, (-1, self.input_dim))
        xx = self._scaled(x)
        precision = self.precision or E.default_precision()
        W = self._predictive_weights_dev(xx, dist)
        Ms, Ks, psis, nus = self.models.posterior.params
        o = self.output_dim
        K = self.size
        psis = np.ascontiguousarray(np.broadcast_to(psis, (K, o, o)))
        dfs = np.broadcast_to(np.asarray(nus, dtype=np.float64) - self.models.likelihood.row_dim + 1, (K,))
        yy = None
        if y is not None:
            yy = np.reshape(y, (-1, self.output_dim))
            yy = self.output_transform.transform(yy) if self.scale else yy
        mu, covar, nlpd = E.predict_lingauss(
            E.to_dev(xx, E.tdtype(precision)), W, Ms, np.linalg.inv(Ks), np.linalg.inv(psis), psis, np.linalg.slogdet(psis)[1],
            dfs, self.models.likelihood.affine, 0 if prediction == 'average' else 1, dist == 'studentt', precision,
            Y=None if yy is None else E.to_dev(yy, E.tdtype(precision)), eps=eps)
        mu, covar = E.to_host(mu).astype(np.float64), E.to_host(covar).astype(np.float64)
        nlpd = None if nlpd is None else E.to_host(nlpd).astype(np.float64)
        if self.scale:
            mu = self.output_transform.inverse_transform(mu)
            mat = np.diag(np.sqrt(self.output_transform.var_))
            covar = np.einsum('kh,...hj,ji->...ki', mat, covar, mat.T)
        if incremental:
            mu += x[:, :self.output_dim]
        var = np.vstack(list(map(np.diag, covar)))
        out = (mu, var if variance == 'diagonal' else covar, np.sqrt(var))
        return out + (nlpd,) if y is not None else out
