"""Device-resident sweep session shared by the GMM and ILR drivers.

One session = data z = [x | y] uploaded once, one packed operand buffer (W | S,T and cst)
that every model part writes its block into, one packed FP64 statistics buffer, and the
fused sweep kernel between them:

    statistics --(batched posterior kernels)--> operands --(mimo_sweep)--> statistics

Mean-field: the data + label terms of the lower bound equal sum_n logsumexp_k (E log joint)
whenever the responsibilities are the E-step of the current posterior, so the
sweep returns that scalar and the reference's second E-step per iteration
(mixtures/gmm.py:338-339) is not needed.
"""
import numpy as np
import numpy.random as npr
import torch

from .. import _engine as E
from ..distributions.bayesian import MEANFIELD, GIBBS


class Part:
    """a conjugate wrapper plus where its variables sit in zt and its rows in W."""

    def __init__(self, wrapper, stat_idx=None, col_map=None):
        self.w = wrapper
        self.stat_idx, self.col_map = stat_idx, col_map

    def layout(self, Dp, row_off):
        return dict(stat_idx=self.stat_idx, col_map=self.col_map, Dp=Dp, row_off=row_off)


class Session:

    def __init__(self, z, K, gating, parts, family, precision=None, comm=None):
        self.precision = precision or E.default_precision()
        self.K, self.gating, self.parts, self.family = K, gating, parts, family
        self.Z = z if isinstance(z, torch.Tensor) else E.to_dev(np.ascontiguousarray(z), E.tdtype(self.precision))
        assert self.Z.is_cuda and self.Z.dtype == E.tdtype(self.precision), \
            'a device tensor passed as data must be %s (precision %r)' % (E.tdtype(self.precision), self.precision)
        self.N, self.D = self.Z.shape
        self.feats = E.quad_features(self.D) if family == 'quad' else E.diag_features(self.D)
        self.F = self.feats.F
        self.count_feature = self.F - 1
        self.comm = comm                     # optional sharded.Communicator: all-reduce of stat / lse
        self._ops = {}
        self._bufs = {}
        self.stat = E.zeros((K, self.F))
        self.lse_sum = E.zeros((1,))
        self.gating_prior = gating._prior_dev() if gating is not None else None
        self.part_priors = [p.w._prior_dev() for p in parts]
        self.last = None
        self._absmax = None

    # -- buffers ---------------------------------------------------------------------------
    def ops(self, mode):
        if mode not in self._ops:
            if self.family == 'quad':
                rows = sum(p.w._rows(mode) for p in self.parts)
                self._ops[mode] = E.QuadOperands(self.K, self.D, rows, self.precision)
            else:
                self._ops[mode] = E.DiagOperands(self.K, self.D, self.precision)
        return self._ops[mode]

    def buf(self, hard):
        if hard not in self._bufs:
            self._bufs[hard] = E.SweepBuffers(self.N, self.K, self.F, self.precision, hard)
        return self._bufs[hard]

    # -- statistics from explicit labels / responsibilities -------------------------------------
    def stats_from_labels(self, labels):
        lab = E.to_dev(np.asarray(labels, dtype=np.int32), torch.int32)
        self.stat = self._reduce(E.stats_hard(self.Z, lab, self.K, self.feats, self.precision))
        return self.stat

    def stats_from_resp(self, resp):
        R = resp if isinstance(resp, torch.Tensor) else E.to_dev(resp, E.tdtype(self.precision))
        self.stat = self._reduce(E.stats_soft(self.Z, R, self.feats, self.precision))
        return self.stat

    def stats_from_random_resp(self, seed=0, chunk=None):
        """statistics of random responsibilities (mixtures/gmm.py:265-267: uniform variates normalised over the
        components) drawn ON THE DEVICE in point chunks, so no (K, N) array ever exists on the host -- the start of
        `meanfield_coordinate_descent(randomize='device')` at sizes where npr.rand(K, N) cannot be built.  The
        generator is keyed by (seed, global chunk index): the draw does not depend on the shard count."""
        K, N = self.K, self.N
        tdt = E.tdtype(self.precision)
        chunk = chunk or max(256, min(1 << 20, (1 << 28) // max(K, 1)))
        offset = self.comm.point_offset if self.comm is not None else 0
        stat = E.zeros((K, self.F))
        g = torch.Generator(device=self.Z.device)
        lo = 0
        while lo < N:
            # global chunk grid so that shards reproduce the single-process draw
            gchunk = (offset + lo) // chunk
            hi = min(N, (gchunk + 1) * chunk - offset)
            g.manual_seed(int(seed) * 1000003 + gchunk)
            r = torch.rand((K, chunk), generator=g, device=self.Z.device, dtype=tdt)
            a0 = offset + lo - gchunk * chunk
            r = r[:, a0:a0 + (hi - lo)]
            r = (r / r.sum(0, keepdim=True)).contiguous()
            if self.precision == 'fp32' and self.family == 'quad' and 24 <= self.D <= 128:
                E.stats_soft_tc(self.Z[lo:hi], r, self.feats, stat=stat)      # tcgen05 feature GEMM (accumulates)
            else:
                E.stats_soft(self.Z[lo:hi], r, self.feats, self.precision, stat=stat)
            lo = hi
        self.stat = self._reduce(stat)
        return self.stat

    def _reduce(self, stat):
        if self.comm is not None:
            self.comm.allreduce(stat)
        return stat

    # -- operands --------------------------------------------------------------------------
    def update_from_stats(self, mode, variates=None, gating_variates=None, want_lik=False):
        """gating + every part: posterior update from self.stat, operands into ops(mode)."""
        ops = self.ops(mode)
        outs = {}
        if self.gating is not None:
            outs['gating'] = self.gating._update(self.stat, self.F, self.count_feature, mode, ops=ops,
                                                 variates=gating_variates, prior_dev=self.gating_prior)
        else:
            ops.cst.zero_()
        row = 0
        outs['parts'] = []
        k_range = self._component_shard()
        for i, p in enumerate(self.parts):
            lay = p.layout(self.D + 1, row)
            kw = dict(k_range=k_range) if k_range is not None else {}
            outs['parts'].append(p.w._update(self.stat, self.F, lay, mode, ops=ops,
                                             variates=None if variates is None else variates[i],
                                             prior_dev=self.part_priors[i], want_lik=want_lik, **kw))
            row += p.w._rows(mode) if self.family == 'quad' else 0
        if k_range is not None:
            self._gather_components(ops, outs, k_range)
        self.last = outs
        return ops, outs

    # -- posterior update sharded over ranks -------------------------------------------------
    def _component_shard(self):
        """[k0, k1) of this rank when the posterior update can be split over ranks: every rank holds the all-reduced
        statistics, the kernels are one CTA per component, so each rank updates K / world components and the operand
        blocks (what the next sweep reads) and lower-bound terms are all-gathered -- instead of every rank repeating
        all K Cholesky factorisations.  Needs untied components (tied ones average over k) and K divisible by world."""
        import os
        c = self.comm
        if c is None or c.world <= 1 or self.K % c.world or self.family != 'quad' or os.environ.get('MIMO_REPLICATED_POSTERIOR'):
            return None
        if not all(getattr(p.w, '_shardable', False) and not p.w._tied for p in self.parts):
            return None
        n = self.K // c.world
        return (c.rank * n, (c.rank + 1) * n)

    def _gather(self, t, k_range):
        """all-gather the component slices of a (K, ...) tensor in place."""
        import torch.distributed as dist
        lo, hi = k_range
        mine = t[lo:hi].clone()
        dist.all_gather_into_tensor(t.view(-1), mine.view(-1), group=self.comm.group)

    def _gather_components(self, ops, outs, k_range):
        self._gather(ops.W, k_range)
        self._gather(ops.cst, k_range)
        for o in outs['parts']:
            if o.get('vlb') is not None:
                self._gather(o['vlb'], k_range)
            o['gathered'] = False                    # posterior parameters still hold this rank's slice only

    def _gather_parameters(self, outs):
        """before downloading posterior parameters into the model (store): complete the per-rank slices."""
        for o in outs['parts']:
            if o.get('gathered') is False:
                for key in ('m', 'kappa', 'psi', 'nu', 'lik_mu', 'lik_lmbda'):
                    if o.get(key) is not None:
                        self._gather(o[key], o['k_range'])
                o['gathered'] = True

    def operands_from_posterior(self, gating_mode=MEANFIELD):
        """operands of the CURRENT posteriors (no new statistics)."""
        ops = self.ops(MEANFIELD)
        zero = E.zeros((self.K, self.F))
        infos = []
        if self.gating is not None:
            pa, pb = self.gating._prior_arrays(self.gating.posterior)
            g = self.gating._update(zero, self.F, self.count_feature, gating_mode, ops=ops,
                                    prior_dev=(E.to_dev(pa), E.to_dev(pb) if pb is not None else None))
            infos.append(g['info'])
        else:
            ops.cst.zero_()
        row = 0
        for p in self.parts:
            infos.append(p.w._posterior_operands(ops, p.layout(self.D + 1, row)))
            row += p.w._rows(MEANFIELD) if self.family == 'quad' else 0
        for i in infos:
            i.check()
        return ops

    def operands_from_likelihood(self, log_probs=None):
        """operands of the explicit likelihood parameters + log gating probabilities."""
        ops = self.ops(GIBBS)
        if log_probs is None:
            ops.cst.zero_()
        else:
            E.set_log_weights(ops, log_probs)
        row = 0
        for p in self.parts:
            p.w._likelihood_operands(ops, p.layout(self.D + 1, row)).check()
            row += p.w._rows(GIBBS) if self.family == 'quad' else 0
        return ops

    # -- the sweep -------------------------------------------------------------------------
    def sweep(self, ops, hard, uniforms=None, seed=0, ll_out=None, phase_ms=None):
        buf = self.buf(hard)
        u = None
        if uniforms is not None:
            u = uniforms if isinstance(uniforms, torch.Tensor) else E.to_dev(np.asarray(uniforms).reshape(-1))
        offset = self.comm.point_offset if self.comm is not None else 0
        if self._absmax is None:                 # the data are resident and never change: their scale is taken once
            self._absmax = float(self.Z.abs().max().item()) if (self.N > 0 and self.precision == 'fp32') else 0.0
        E.sweep(self.Z, ops, self.feats, buf, uniforms=u, seed=seed, offset=offset, ll_out=ll_out,
                phase_ms=phase_ms, absmax=self._absmax)
        if self.comm is not None:
            self.comm.allreduce(buf.flat)
        self.stat, self.lse_sum = buf.stat, buf.lse_sum
        return buf

    def loglik(self, ops):
        return E.loglik(self.Z, ops)

    def capture_meanfield_step(self):
        """One mean-field iteration (batched posterior kernels -> operands -> fused sweep) captured in a CUDA graph: at the
        sizes the reference ships (N ~ 1e3, K ~ 25) an iteration is ~30 kernel launches of microseconds each and the launch
        latency is the whole cost.  Call after at least one eager iteration (buffers allocated, self.stat aliasing the sweep
        buffer).  Returns (graph, outs): graph.replay() performs the iteration, `outs` are the tensors it refreshes."""
        assert self.comm is None or self.comm.world == 1, 'graph capture of a sharded iteration is not supported'
        buf = self.buf(False)
        assert self.stat is buf.stat, 'run one eager iteration first'
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            ops, outs = self.update_from_stats(MEANFIELD)
            self.sweep(ops, hard=False)
        return graph, outs

    # -- scalars ---------------------------------------------------------------------------
    def lower_bound(self, outs):
        """gating term + component terms (from the posterior kernels that produced the
        operands) + sum_n lse_n (from the sweep that used them).  One device->host read."""
        terms = [self.lse_sum.reshape(1)]
        if 'gating' in outs:
            terms.append(outs['gating']['vlb'].reshape(1))
        for o in outs['parts']:
            terms.append(o['vlb'].sum().reshape(1))
        return float(torch.cat(terms).sum().item())

    def check(self, outs):
        infos = ([outs['gating']['info']] if 'gating' in outs else []) + [o['info'] for o in outs['parts']]
        if self.comm is not None and self.comm.world > 1:
            # every rank must raise together (a rank that factorised only its K / world components would otherwise
            # fail alone and leave the others waiting in the next collective)
            import torch.distributed as dist
            word = torch.stack([i.t for i in infos]).view(-1)
            dist.all_reduce(word, op=dist.ReduceOp.MAX, group=self.comm.group)
            for n, i in enumerate(infos):
                i.t.copy_(word.view(-1, 2)[n])
        for i in infos:
            i.check()

    def store(self, outs, mode, set_probs=True):
        """download posterior (and sampled / mode likelihood) parameters into the model."""
        self._gather_parameters(outs)
        if 'gating' in outs:
            self.gating._store(outs['gating'], set_probs=set_probs)
        for p, o in zip(self.parts, outs['parts']):
            p.w._store(o, mode)

    def counts_host(self):
        return E.to_host(self.stat[:, self.count_feature])

    def host_draw(self, fn):
        """fn() on rank 0 (its numpy.random stream is THE stream of a sharded chain), the result broadcast to every
        rank: all ranks then update the same parameters from the same variates.  Without a communicator: fn()."""
        c = self.comm
        if c is None or c.world <= 1:
            return fn()
        import torch.distributed as dist
        box = [fn() if c.rank == 0 else None]
        src = dist.get_global_rank(c.group, 0) if c.group is not None else 0
        dist.broadcast_object_list(box, src=src, group=c.group)
        return box[0]

    def seed_parameters(self, seed):
        """(re)seed the device generator of draw_gibbs_variates('device'); every rank of a sharded run must pass the
        same seed (they then draw the same variates from the same all-reduced statistics: no broadcast)."""
        self._param_gen = torch.Generator(device=self.Z.device)
        self._param_gen.manual_seed(int(seed))

    def draw_gibbs_variates(self, rng='numpy'):
        """host draws from the global numpy.random stream in the reference's order:
        every part (per component), then the gating.  Sharded: drawn once, on rank 0.
        rng='device': the same variates (gamma / chi-square / normal, shaped by the current statistics) from a device
        generator, with no host read of the statistics and no broadcast -- what a sweep of a few milliseconds needs
        (cfg3: the host draw + broadcast cost as much as the sweep at 8 GPUs).  Not the reference's stream."""
        if rng == 'device':
            if getattr(self, '_param_gen', None) is None:
                self.seed_parameters(0)
            counts = self.stat[:, self.count_feature]
            var = [p.w._draw_variates_device(counts, self.stat, self._param_gen, self.part_priors[i])
                   for i, p in enumerate(self.parts)]
            gvar = self.gating._draw_variates_device(counts, self._param_gen, self.gating_prior) \
                if self.gating is not None else None
            return var, gvar
        counts = self.counts_host()
        stat_host = E.to_host(self.stat) if self.family == 'diag' else None

        def draw():
            var = []
            for p in self.parts:
                var.append(p.w._draw_variates(counts, stat_host) if self.family == 'diag' else p.w._draw_variates(counts))
            gvar = self.gating._draw_variates(counts) if self.gating is not None else None
            return var, gvar
        return self.host_draw(draw)

    def label_uniforms(self, rng='numpy'):
        """one uniform per point for the inverse-CDF label draw (utils/stats.py:14).
        'numpy': npr.random(size=(1, N)) -- the reference's single RNG call; sharded: rank 0 draws all N_global values
        and every rank keeps its slice, so the chain equals the single-process chain.
        'philox': no host variates at all -- returns (None, seed): the kernel draws Philox4x32-10 uniforms keyed by
        (seed, GLOBAL point index), which is what sizes like N = 1e8 need."""
        if rng == 'philox':
            return None, int(self.host_draw(lambda: int(npr.randint(1 << 30))))
        c = self.comm
        if c is None or c.world <= 1:
            return npr.random(size=(1, self.N)), 0
        n_global = c.N_global if c.N_global is not None else None
        assert n_global is not None, 'a sharded label draw from host uniforms needs Communicator(N_global=...)'
        u = self.host_draw(lambda: npr.random(size=(1, n_global)))
        return u[:, c.point_offset:c.point_offset + self.N], 0


def random_responsibilities(K, N):
    """npr.rand(K, N) normalised over components (mixtures/gmm.py:265-267)."""
    resp = npr.rand(K, N)
    resp /= np.sum(resp, axis=0)
    return resp
