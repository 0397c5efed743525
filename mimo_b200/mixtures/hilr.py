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
                                     progress_bar=True, process_id=0, lower_bound=False):
        """The reference returns an empty list (its lower-bound line is commented out, hilr.py:194); lower_bound=True
        returns the bound of every iteration instead (parameter terms + sum_n logsumexp from the fused sweep) and
        stops on tol like the other drivers."""
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
                if lower_bound:
                    vlb.append(float(lse.item()) + float(outs['gating']['vlb'].item())
                               + sum(float(o['vlb'].sum().item()) for o in outs['parts']))
                    if len(vlb) > 1 and abs(vlb[-1] - vlb[-2]) < tol:
                        break
                pbar.update(1)
        if outs is not None:
            s.store(outs, MEANFIELD, set_probs=False)
        return vlb

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
