"""Device-resident stochastic variational inference for Bayesian mixtures of Gaussians and of linear-Gaussian experts
(SURVEY 8 a9 / f3).

mixtures/gmm.py:300-336 and mixtures/ilr.py:245-291 of the reference: per iteration one minibatch -> E-step of the current posterior on it ->
weighted statistics -> natural-parameter blend  posterior <- (1 - rho) posterior + rho (prior + statistics / scale)
(distributions/bayesian.py:85-91, 161-171, 232-238) -> full-data lower bound.

The API path (gmm.meanfield_stochastic_descent(device=False)) runs the E-step and the statistics of a minibatch through
the kernels and blends on the host.  Here nothing but the minibatch indices crosses the boundary per iteration:

  * the data stay resident; a minibatch is a device gather of `batch_size` rows;
  * the blend is linear in natural parameters, so it is folded into the conjugate-update kernel: the PSEUDO-PRIOR
    (1 - rho) nat(posterior) + rho nat(prior) is formed on the device (K-sized tensor algebra, one batched d x d inverse),
    and `mimo_nw_posterior(pseudo-prior, (rho / scale) statistics)` (`mimo_mnw_posterior` for the experts) returns the
    blended posterior, its Cholesky factors and the E-step operands of the next iteration in one call -- tied covariances
    included (the kernel's mean over k acts on the blended natural parameters exactly like composite.py:275-283, 800-808);
  * the lower-bound terms of the parameters (entropy - cross-entropy of Normal-Wishart / Dirichlet / stick-breaking
    posteriors against their priors) are closed forms evaluated on the device, the data term is the fused sweep's
    sum_n logsumexp: the bound of every iteration lands in a device vector that is read once at the end;
  * with graph=True one iteration (gather ... blend ... full-data sweep ... bound) is captured in a CUDA graph and
    replayed: a 64-point step is launch-latency bound (~40 launches), the replay removes that latency.
"""
import math

import numpy as np
import torch

from .. import _engine as E
from ..distributions.bayesian import MEANFIELD, CategoricalWithDirichlet


def _outer(a):
    return a[:, :, None] * a[:, None, :]


def _logdet_spd(a):
    L, _ = torch.linalg.cholesky_ex(a)
    return 2. * torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)


def _inv(a):
    return torch.linalg.inv_ex(a)[0]


def nw_lower_bound(prior, post):
    """entropy(q) - cross_entropy(q, p) per component for Normal-Wisharts q = post, p = prior (composite.py:120-134 with
    the natural parameters / expected statistics of :50-65, :106-118 and the log-partition of :95-98, wishart.py:129-132),
    in torch FP64 on the device."""
    m, k, psi, nu = post
    d = m.shape[1]
    ar = torch.arange(d, dtype=torch.float64, device=m.device)
    e_lm = nu[:, None] * torch.einsum('kdl,kl->kd', psi, m)
    stats = (e_lm, -0.5 * (d / k + (m * e_lm).sum(1)), -0.5 * nu[:, None, None] * psi,
             0.5 * (torch.special.digamma((nu[:, None] - ar[None, :]) / 2.).sum(1) + d * math.log(2.) + _logdet_spd(psi)))

    def free(p):
        pm, pk, ppsi, pnu = p
        logz = -0.5 * d * torch.log(pk) + 0.5 * pnu * d * math.log(2.) + torch.special.multigammaln(pnu / 2., d) \
            + 0.5 * pnu * _logdet_spd(ppsi)
        c = _inv(ppsi) + pk[:, None, None] * _outer(pm)
        dot = (pk[:, None] * pm * stats[0]).sum(1) + pk * stats[1] + (c * stats[2]).sum((1, 2)) + (pnu - d) * stats[3]
        return logz - dot
    return free(post) - free(prior)


def dirichlet_lower_bound(a0, a):
    """bayesian.py:93-96 with dirichlet.py:78-97."""
    s = torch.special.digamma(a) - torch.special.digamma(a.sum())

    def free(x):
        return torch.lgamma(x).sum() - torch.lgamma(x.sum()) - ((x - 1.) * s).sum()
    return free(a) - free(a0)


def stick_lower_bound(prior, post):
    """bayesian.py:173-176 with dirichlet.py:195-214."""
    g, dl = post
    es = torch.special.digamma(g) - torch.special.digamma(g + dl)
    er = torch.special.digamma(dl) - torch.special.digamma(g + dl)

    def free(x, y):
        return (torch.lgamma(x) + torch.lgamma(y) - torch.lgamma(x + y)).sum() - ((x - 1.) * es + (y - 1.) * er).sum()
    return free(g, dl) - free(*prior)


def mnw_lower_bound(prior, post):
    """the same for Matrix-Normal-Wisharts (composite.py:577-599, 622-663)."""
    M, Kc, psi, nu = post
    o, c = M.shape[1], M.shape[2]
    ar = torch.arange(o, dtype=torch.float64, device=M.device)
    e_la = nu[:, None, None] * torch.einsum('kdl,klm->kdm', psi, M)
    stats = (e_la, -0.5 * (o * _inv(Kc) + torch.einsum('kdl,kdm->klm', M, e_la)), -0.5 * nu[:, None, None] * psi,
             0.5 * (torch.special.digamma((nu[:, None] - ar[None, :]) / 2.).sum(1) + o * math.log(2.) + _logdet_spd(psi)))

    def free(p):
        pM, pK, ppsi, pnu = p
        logz = -0.5 * o * _logdet_spd(pK) + 0.5 * pnu * o * math.log(2.) + torch.special.multigammaln(pnu / 2., o) \
            + 0.5 * pnu * _logdet_spd(ppsi)
        mk = torch.einsum('kdl,klm->kdm', pM, pK)
        cc = _inv(ppsi) + torch.einsum('kdm,khm->kdh', mk, pM)
        dot = (mk * stats[0]).sum((1, 2)) + (pK * stats[1]).sum((1, 2)) + (cc * stats[2]).sum((1, 2)) + (pnu - o - 1. + c) * stats[3]
        return logz - dot
    return free(post) - free(prior)


class _NWPart:
    """Normal-Wishart block (components of a GMM, input densities of an ILR): state and blend."""
    keys = ('m', 'kappa', 'psi', 'nu')
    bound = staticmethod(nw_lower_bound)

    def __init__(self, wrapper, prior_dev, layout):
        self.w, self.lay = wrapper, layout
        self.prior = [t.clone() for t in prior_dev]
        self.post = [E.to_dev(np.asarray(p, dtype=np.float64)) for p in wrapper.posterior.params]
        self.prior_psi_inv = _inv(self.prior[2])            # constant: taken once

    def pseudo_prior(self, rho):
        m, k, psi, nu = self.post
        m0, k0, psi0, nu0 = self.prior
        kq = (1. - rho) * k + rho * k0
        mq = ((1. - rho) * k[:, None] * m + rho * k0[:, None] * m0) / kq[:, None]
        cq = (1. - rho) * (_inv(psi) + k[:, None, None] * _outer(m)) + rho * (self.prior_psi_inv + k0[:, None, None] * _outer(m0))
        return [mq.contiguous(), kq.contiguous(), _inv(cq - kq[:, None, None] * _outer(mq)).contiguous(),
                ((1. - rho) * nu + rho * nu0).contiguous()]

    def store(self):
        self.w.posterior.params = tuple(E.to_host(t) for t in self.post)


class _MNWPart(_NWPart):
    """Matrix-Normal-Wishart block (the experts of an ILR)."""
    keys = ('M', 'K', 'psi', 'nu')
    bound = staticmethod(mnw_lower_bound)

    def pseudo_prior(self, rho):
        M, Kc, psi, nu = self.post
        M0, K0, psi0, nu0 = self.prior
        Kq = (1. - rho) * Kc + rho * K0
        mk, mk0 = torch.einsum('kdl,klm->kdm', M, Kc), torch.einsum('kdl,klm->kdm', M0, K0)
        Mq = torch.einsum('kdl,klm->kdm', (1. - rho) * mk + rho * mk0, _inv(Kq))
        cq = (1. - rho) * (_inv(psi) + torch.einsum('kdm,khm->kdh', mk, M)) + rho * (self.prior_psi_inv + torch.einsum('kdm,khm->kdh', mk0, M0))
        inner = cq - torch.einsum('kdl,klm,khm->kdh', Mq, Kq, Mq)
        return [Mq.contiguous(), Kq.contiguous(), _inv(inner).contiguous(), ((1. - rho) * nu + rho * nu0).contiguous()]


class DeviceSVI:
    """state and one iteration of device-resident SVI: a session (resident data, operand block), its gating and its
    conjugate blocks (one Normal-Wishart block for a GMM; input densities + experts for an ILR)."""

    def __init__(self, model, session, batch_size, step_size):
        from ..distributions.bayesian import StackedGaussiansWithNormalWisharts, StackedLinearGaussiansWithMatrixNormalWisharts
        s = self.s = session
        self.model = model
        self.K, self.B, self.rho = model.size, int(batch_size), float(step_size)
        self.scale = self.B / float(s.N)
        self.dirichlet = isinstance(model.gating, CategoricalWithDirichlet)
        self.parts = []
        row = 0
        for p, prior_dev in zip(s.parts, s.part_priors):
            lay = p.layout(s.D + 1, row)
            if isinstance(p.w, StackedLinearGaussiansWithMatrixNormalWisharts):
                self.parts.append(_MNWPart(p.w, prior_dev, lay))
            elif isinstance(p.w, StackedGaussiansWithNormalWisharts):
                self.parts.append(_NWPart(p.w, prior_dev, lay))
            else:
                raise NotImplementedError('device-resident SVI: %s blocks are not supported' % type(p.w).__name__)
            row += p.w._rows(MEANFIELD)
        ga, gb = model.gating._prior_arrays(model.gating.prior)
        pa, pb = model.gating._prior_arrays(model.gating.posterior)
        self.gprior = (E.to_dev(ga), E.to_dev(gb) if gb is not None else None)
        self.gpost = [E.to_dev(np.asarray(pa, dtype=np.float64)), E.to_dev(np.asarray(pb, dtype=np.float64)) if pb is not None else None]
        self.idx = torch.zeros((self.B,), dtype=torch.int64, device=s.Z.device)
        self.Zb = torch.empty((self.B, s.D), dtype=s.Z.dtype, device=s.Z.device)
        self.stat_b = E.zeros((self.K, s.F))
        self.zero_stat = E.zeros((self.K, s.F))
        self.bound = E.zeros((1,))
        self.ops = s.ops(MEANFIELD)
        self.ops_current = False            # the block holds the operands of the current posteriors
        self.infos = []

    # -- pieces ----------------------------------------------------------------------------
    def set_batch(self, batch):
        self.idx.copy_(torch.as_tensor(np.asarray(batch, dtype=np.int64)), non_blocking=True)

    def _gating_operands(self):
        g = self.model.gating._update(self.zero_stat, self.s.F, self.s.count_feature, MEANFIELD, ops=self.ops,
                                      prior_dev=(self.gpost[0], self.gpost[1]))
        self.infos.append(g['info'])

    def _write_operands(self):
        """operands of the CURRENT posteriors into the session's block: gating sets cst, the blocks add theirs."""
        self._gating_operands()
        for p in self.parts:
            c = p.w._update(self.zero_stat, self.s.F, p.lay, MEANFIELD, ops=self.ops, prior_dev=p.post, want_vlb=False)
            self.infos.append(c['info'])

    def stats_from_resp(self, resp):
        """statistics of the gathered minibatch for explicit responsibilities (K, B) (the randomised first iteration)."""
        torch.index_select(self.s.Z, 0, self.idx, out=self.Zb)
        self.stat_b.zero_()
        E.stats_soft(self.Zb, E.to_dev(resp, self.Zb.dtype), self.s.feats, self.s.precision, stat=self.stat_b)

    def estep_batch(self):
        """E-step of the current posterior on the gathered minibatch -> its weighted statistics."""
        torch.index_select(self.s.Z, 0, self.idx, out=self.Zb)
        if not self.ops_current:            # (after a blend they are already there)
            self._write_operands()
        a = E.loglik(self.Zb, self.ops)
        E.softmax(a, self.s.precision, resp=True)
        self.stat_b.zero_()
        E.stats_soft(self.Zb, a, self.s.feats, self.s.precision, stat=self.stat_b)

    def blend(self):
        """posterior <- (1 - rho) posterior + rho (prior + statistics / scale) in natural parameters; leaves the operands
        of the new posteriors in the session's block."""
        rho = self.rho
        counts = self.stat_b[:, self.s.count_feature]
        # (the state tensors are updated IN PLACE: a captured iteration reads and writes the same buffers on every replay)
        self.gpost[0].copy_((1. - rho) * self.gpost[0] + rho * (self.gprior[0] + counts / self.scale))
        if not self.dirichlet:
            tail = torch.flip(torch.cumsum(torch.flip(counts, [0]), 0), [0])
            acc = torch.cat((tail[1:], tail.new_zeros(1)))
            self.gpost[1].copy_((1. - rho) * self.gpost[1] + rho * (self.gprior[1] + acc / self.scale))
        self._gating_operands()
        # conjugate blocks: pseudo-prior in standard form, then the conjugate kernel on the scaled statistics
        scaled = self.stat_b * (rho / self.scale)
        for p in self.parts:
            out = p.w._update(scaled, self.s.F, p.lay, MEANFIELD, ops=self.ops, prior_dev=p.pseudo_prior(rho), want_vlb=False)
            for dst, key in zip(p.post, p.keys):
                dst.copy_(out[key])
            self.infos.append(out['info'])
        self.ops_current = True

    def lower_bound(self):
        """full-data bound at the current posteriors (their operands are in the block): one fused sweep + closed forms."""
        self.s.sweep(self.ops, hard=False)
        total = self.s.lse_sum.reshape(()) + (dirichlet_lower_bound(self.gprior[0], self.gpost[0]) if self.dirichlet
                                              else stick_lower_bound(self.gprior, (self.gpost[0], self.gpost[1])))
        for p in self.parts:
            total = total + p.bound(p.prior, p.post).sum()
        self.bound.copy_(total.reshape(1))

    def iteration(self, with_bound=True):
        self.estep_batch()
        self.blend()
        if with_bound:
            self.lower_bound()

    # -- results ---------------------------------------------------------------------------
    def check(self):
        for i in self.infos:
            i.check()
        self.infos = []

    def store(self):
        """download the posteriors into the model (the likelihood objects are left alone: the reference leaves SAMPLED
        parameters there)."""
        for p in self.parts:
            p.store()
        g = self.model.gating
        if self.dirichlet:
            g.posterior.alphas = E.to_host(self.gpost[0])
        else:
            g.posterior.gammas, g.posterior.deltas = E.to_host(self.gpost[0]), E.to_host(self.gpost[1])


def run(model, session, randomize, maxiter, step_size, batch_size, graph, lower_bound_every, batches, random_responsibilities, pbar=None):
    """the loop of gmm.py:300-326 / ilr.py:245-278 on the device.  Returns the list of lower bounds (one per iteration
    where it was asked for; read from the device once, at the end)."""
    s = session
    st = DeviceSVI(model, s, batch_size, step_size)
    every = max(1, int(lower_bound_every))
    bounds = E.zeros((maxiter,))
    asked = []
    graphs = {}                                         # one capture per kind of iteration: with / without the full-data bound
    for i in range(maxiter):
        for batch in batches(batch_size, s.N):
            st.set_batch(batch)
            want = (i + 1) % every == 0 or i + 1 == maxiter
            if i == 0 and randomize is True:
                st.stats_from_resp(random_responsibilities(model.size, len(batch)))
                st.blend()
                if want:
                    st.lower_bound()
            elif graph and want in graphs:
                graphs[want].replay()
            elif graph and i >= 1 and st.ops_current:
                st.iteration(want)                      # one eager iteration allocates every buffer ...
                st.check()
                torch.cuda.synchronize()
                graphs[want] = torch.cuda.CUDAGraph()   # ... the next ones replay this capture (capturing does not execute)
                with torch.cuda.graph(graphs[want]):
                    st.iteration(want)
                st.infos = []
            else:
                st.iteration(want)
            if want:
                bounds[i:i + 1].copy_(st.bound)
                asked.append(i)
        if pbar is not None:
            pbar.update(1)
    st.check()
    st.store()
    vals = E.to_host(bounds)
    return [float(vals[i]) for i in asked]
