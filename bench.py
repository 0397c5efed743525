#!/usr/bin/env python
"""bench.py -- points*components/s of one full inference sweep (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5] [--impl reference]

A "step" is one full sweep of the named workload over synthetic data of its shape:
  mean-field : batched posterior kernels (statistics -> posteriors, operands, lower-bound
               terms) -> fused E-step + statistics sweep -> [all-reduce] -> lower bound read
  Gibbs      : host-drawn parameter variates -> posterior kernels (draw) -> fused E-step +
               label draw + statistics sweep -> [all-reduce]

What the one JSON line holds (default workload cfg5: N=50M, d=128, K=1024, mean field):
  value / roofline : the DENSE regime -- every (point, component) pair goes through the 3-pass
               tensor-core E-step and the tensor-core statistics GEMM, whatever the data look like.
               `roofline.frac` = SURVEY 8(d) algorithmic flops of the dominant kernel / its measured
               device time / the sustained bf16 peak of MEASURED_PEAKS.json.  Timed for --steps.
  screened_path   : the library's default path on the same data and model (screening pass + exact
               refinement of the candidate pairs; falls back to the dense kernels per chunk on the
               device).  Its bound is the TMEM read port, not the algorithmic roofline; candidate
               fraction and tier are totals over ALL chunks of the last sweep, per rank.
  overlap_regime  : a second data set whose components overlap (centre spread 1.5 sigma), started
               from random responsibilities, every sweep timed from the first one.
  parity_subsample: the CUDA results on a sub-sample of the resident data against oracle/ (log-joint,
               responsibilities, log-normalisers, statistics, posterior parameters), max scaled errors.
  e2e             : the full step through host buffers: posterior update from the previous
               statistics, mimo_sweep_host (pinned host data -> device, sweep, statistics back),
               lower bound, every step.
  other_configs   : short legs of cfg1-cfg4 (BASELINE.json's other shapes) with their own roofline.
`--impl reference` times the CPU port of the reference's algorithm (oracle/) on the box's
host cores on a bounded sample of the same workload.
Multi-GPU (torchrun): the N points are split across ranks (strong scaling); one all-reduce
of the packed FP64 statistics per sweep.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault('NCCL_DEBUG', 'WARN')      # keep stdout to the one JSON line (NCCL prints its version banner there)

WORKLOADS = {
    # name: family, N, d (or d_in), o, K, mode
    'cfg1': dict(kind='gmm', N=2500, d=2, K=25, mode='vi', desc='examples/gmm sine-shaped Bayesian GMM, NW + Dirichlet'),
    'cfg2': dict(kind='ilr', N=10_000_000, d=8, o=1, K=128, mode='vi',
                 desc='stick-breaking mixture of linear-Gaussian experts (tied MNW), N=10M d_in=8 d_out=1 K=128'),
    'cfg3': dict(kind='dgmm', N=100_000_000, d=64, K=256, mode='gibbs',
                 desc='diagonal-covariance GMM (Normal-Gamma), Gibbs, N=100M d=64 K=256'),
    'cfg4': dict(kind='gmm', N=10_000_000, d=16, K=64, mode='vi', stick=True,
                 desc='DP-GMM full covariance, mean-field, N=10M d=16 K=64'),
    'cfg5': dict(kind='gmm', N=50_000_000, d=128, K=1024, mode='vi', stick=True,
                 desc='DP-GMM full covariance, mean-field, N=50M d=128 K=1024'),
}
OVERLAP_N = 8_000_000          # points of the overlapping-components regime (same d, K as the workload)
OVERLAP_SPREAD = 1.5


def read_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return dict(hbm=j['hbm_gbs'], tf_burst=j['bf16_tflops'], tf_sus=j.get('bf16_tflops_sustained', j['bf16_tflops']),
                        src='measured')
        except Exception:
            pass
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src='fallback')


def ncu_traffic(kernel):
    """DRAM bytes (read + write) per launch of a kernel, from the committed `ncu --set full` capture of the same
    kernel on one point chunk of cfg5 (profiles/ncu_traffic.json); None when not captured."""
    for fn in ('ncu_traffic.json', 'r01_ncu_traffic.json'):
        try:
            t = json.load(open(os.path.join(ROOT, 'profiles', fn))).get(kernel)
            if t:
                return t
        except Exception:
            pass
    return None


def algorithmic_work(w):
    """flops per (point, component) pair and bytes per point of one sweep."""
    d, K = w['d'], w['K']
    if w['kind'] == 'gmm':
        e = 2 * d * (d + 1) + 2 * d
        s = 2 * (d + 1) ** 2
        bytes_pt = 4 * d
    elif w['kind'] == 'dgmm':
        e = 4 * d + 1
        s = 4 * d + 1
        bytes_pt = 4 * d
    else:
        o = w['o']
        c = d + 1
        e = (2 * d * (d + 1) + 2 * d) + (2 * o * c + 2 * o * o + 2 * c * c + 2 * o + 2 * c)
        s = 2 * c * c + 2 * (c + o) ** 2
        bytes_pt = 4 * (d + o)
    if w['mode'] == 'gibbs':
        return dict(e_flops_pair=e, s_flops_pair=s / K, bytes_pt=bytes_pt + 4)   # hard stats: per point, + labels out
    return dict(e_flops_pair=e, s_flops_pair=s, bytes_pt=bytes_pt)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '100'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[4:8]):
                    if v.strip().lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out = dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------
# synthetic data (on the device, never timed)
# ---------------------------------------------------------------------------------------
def make_data(w, n_local, lo, seed, dev, spread=None):
    """points [lo, lo+n_local) of the workload's synthetic data set, as FP32 on `dev`.
    Blobs: centres ~ N(0, spread^2 I); full-covariance blobs have random SPD covariances
    (Wishart(I, d+2)/d), diagonal blobs per-dimension sigmas in [0.5, 1.5]; labels from
    stick-breaking (alpha=5) or uniform weights.  Generation is chunked and keyed by the
    global chunk index, so every shard count sees the same data set."""
    import torch
    d, K = w['d'], w['K']
    D = d + w.get('o', 0)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    if spread is None:
        spread = 4.0 if w['kind'] != 'ilr' else 3.0
    centres = spread * torch.randn(K, d, generator=g, device=dev)
    if w['kind'] == 'dgmm':
        chol = None
        sig = 0.5 + torch.rand(K, d, generator=g, device=dev)
    else:
        a = torch.randn(K, d, d + 2, generator=g, device=dev)
        cov = a @ a.transpose(1, 2) / d + 0.05 * torch.eye(d, device=dev)
        chol = torch.linalg.cholesky(cov)
        sig = None
    if w.get('stick') or w['kind'] == 'ilr':
        # seeded on the host: every rank and every run sees the same mixture weights (the global torch RNG is not seeded)
        v = torch.from_numpy(np.random.default_rng(seed).beta(1.0, 5.0 * K / 16, size=K)).to(device=dev, dtype=torch.float32)
        v[-1] = 1.0
        pi = v * torch.cumprod(torch.cat([torch.ones(1, device=dev), 1 - v[:-1]]), 0)
        pi = (pi + 0.2 / K)
        pi = pi / pi.sum()
    else:
        pi = torch.full((K,), 1.0 / K, device=dev)
    wvec = torch.randn(d, 1, generator=g, device=dev) if w['kind'] == 'ilr' else None
    Z = torch.empty(n_local, D, dtype=torch.float32, device=dev)
    labels = torch.empty(n_local, dtype=torch.int32, device=dev)
    CH = 1 << 20
    first = lo // CH
    pos = 0
    ci = first
    while pos < n_local:
        c_lo = ci * CH
        gg = torch.Generator(device=dev)
        gg.manual_seed(seed * 1000003 + ci)
        z = torch.multinomial(pi, CH, replacement=True, generator=gg)
        eps = torch.randn(CH, d, generator=gg, device=dev)
        if chol is not None:
            order = torch.argsort(z)
            zs = z[order]
            counts = torch.bincount(zs, minlength=K).tolist()
            x = torch.empty(CH, d, device=dev)
            start = 0
            for k, cnt in enumerate(counts):
                if cnt:
                    idx = order[start:start + cnt]
                    x[idx] = centres[k] + eps[idx] @ chol[k].T
                    start += cnt
        else:
            x = centres[z] + eps * sig[z]
        if w['kind'] == 'ilr':
            y = torch.sin(x @ wvec) + 0.3 * torch.randn(CH, 1, generator=gg, device=dev)
            x = torch.cat([x, y], 1)
        a0 = max(lo, c_lo) - c_lo
        a1 = min(lo + n_local, c_lo + CH) - c_lo
        n = a1 - a0
        Z[pos:pos + n] = x[a0:a1]
        labels[pos:pos + n] = z[a0:a1].to(torch.int32)
        pos += n
        ci += 1
    if w['kind'] == 'ilr':   # standard-scaled, as BayesianMixtureOfLinearGaussians.init_transform does
        mean = Z[: min(n_local, 1 << 20)].mean(0)
        std = Z[: min(n_local, 1 << 20)].std(0)
        Z = (Z - mean) / std
    return Z, labels


def build_model(w, precision='fp32'):
    """priors of SURVEY 8(d) (from the example scripts) with explicit likelihoods (no RNG)."""
    import mimo_b200.distributions as D
    from mimo_b200.mixtures import BayesianMixtureOfGaussians, BayesianMixtureOfLinearGaussians
    K, d = w['K'], w['d']
    if w.get('stick') or w['kind'] == 'ilr':
        gating = D.CategoricalWithStickBreaking(K, D.TruncatedStickBreaking(K, np.ones(K), 5.0 * np.ones(K)),
                                                likelihood=D.Categorical(K))
    else:
        gating = D.CategoricalWithDirichlet(K, D.Dirichlet(K, np.ones(K)), likelihood=D.Categorical(K))
    if w['kind'] == 'gmm':
        prior = D.StackedNormalWisharts(K, d, np.zeros((K, d)), 1e-2 * np.ones(K), np.stack(K * [np.eye(d)]),
                                        (d + 1.0) * np.ones(K) + 1e-8)
        lik = D.StackedGaussiansWithPrecision(K, d, np.zeros((K, d)), np.stack(K * [np.eye(d)]), precision=precision)
        comp = D.StackedGaussiansWithNormalWisharts(K, d, prior=prior, likelihood=lik)
        return BayesianMixtureOfGaussians(gating, comp, precision=precision)
    if w['kind'] == 'dgmm':
        prior = D.StackedNormalGammas(K, d, np.zeros((K, d)), 1e-2 * np.ones((K, d)),
                                      (3.0 + 1e-8) / 2 * np.ones((K, d)), 0.5 * np.ones((K, d)))
        lik = D.StackedGaussiansWithDiagonalPrecision(K, d, np.zeros((K, d)), np.ones((K, d)), precision=precision)
        comp = D.StackedGaussiansWithNormalGammas(K, d, prior=prior, likelihood=lik)
        return BayesianMixtureOfGaussians(gating, comp, precision=precision)
    o, c = w['o'], d + 1
    bprior = D.StackedNormalWisharts(K, d, np.zeros((K, d)), 1e-2 * np.ones(K), np.stack(K * [1e2 * np.eye(d)]),
                                     (d + 1.0) * np.ones(K) + 1e-16)
    basis = D.StackedGaussiansWithNormalWisharts(
        K, d, prior=bprior, likelihood=D.StackedGaussiansWithPrecision(K, d, np.zeros((K, d)), np.stack(K * [np.eye(d)]),
                                                                      precision=precision))
    mprior = D.TiedMatrixNormalWisharts(K, c, o, np.zeros((K, o, c)), np.stack(K * [1e-2 * np.eye(c)]),
                                        np.stack(K * [1e1 * np.eye(o)]), (o + 1.0) * np.ones(K) + 1e-16)
    models = D.TiedLinearGaussiansWithMatrixNormalWisharts(
        K, c, o, mprior, likelihood=D.TiedLinearGaussiansWithPrecision(K, c, o, np.zeros((K, o, c)), np.stack(K * [np.eye(o)]),
                                                                      precision=precision))
    return BayesianMixtureOfLinearGaussians(K, d, o, gating, basis, models, precision=precision)


def make_session(model, w, Z, comm):
    if w['kind'] == 'ilr':
        from mimo_b200.mixtures._driver import Session, Part
        lay = model.models.layout.dev()
        parts = [Part(model.basis, lay['basis_idx'], lay['basis_idx']), Part(model.models, lay['stat_idx'], lay['col_map'])]
        return Session(Z, model.size, model.gating, parts, 'quad', 'fp32', comm=comm)
    return model._session(Z, comm)


# ---------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's sweep on a bounded sample
# ---------------------------------------------------------------------------------------
def cpu_sweep_fn(w, n_sample, seed=0):
    """returns step_fn: one reference-algorithm sweep on n_sample points."""
    from oracle import mimo_oracle as orc
    rng = np.random.default_rng(seed)
    K, d = w['K'], w['d']
    x = rng.standard_normal((n_sample, d)) * 2.0
    if w['kind'] == 'gmm':
        prior = (np.zeros((K, d)), 1e-2 * np.ones(K), np.stack(K * [np.eye(d)]), (d + 1.0) * np.ones(K) + 1e-8)
        state = {'resp': rng.dirichlet(np.ones(K), size=n_sample).T}

        def step():
            post = orc.nw_nat_to_std(orc.add_stats(orc.nw_std_to_nat(*prior), orc.gauss_full_wstats(x, state['resp'])))
            gp, dp = orc.stick_posterior(np.ones(K), 5.0 * np.ones(K), orc.categorical_wstats(state['resp']))
            ell = orc.nw_expected_loglik(x, *post) + orc.stick_expected_log(gp, dp)[0][:, None]
            state['resp'], lse = orc.responsibilities(ell)
            return orc.stick_vlb((np.ones(K), 5.0 * np.ones(K)), (gp, dp)) + np.sum(orc.nw_vlb(prior, post)) + lse.sum()
        return step
    if w['kind'] == 'dgmm':
        prior = (np.zeros((K, d)), 1e-2 * np.ones((K, d)), (3.0 + 1e-8) / 2 * np.ones((K, d)), 0.5 * np.ones((K, d)))
        state = {'labels': rng.integers(0, K, size=n_sample)}

        def step():
            wts = orc.one_hot(state['labels'], K)                        # utils/data.py:160-169
            post = orc.ng_nat_to_std(orc.add_stats(orc.ng_std_to_nat(*prior), orc.gauss_diag_wstats(x, wts)))
            g = rng.gamma(post[2], 1.0 / post[3])
            mu, lam = orc.ng_rvs_from_variates(post[0], post[1], post[2], post[3], g, rng.standard_normal((K, d)))
            probs = orc.dirichlet_probs_from_gammas(rng.standard_gamma(1.0 + orc.categorical_stats(state['labels'], K)))
            lp = orc.gauss_diag_loglik(x, mu, lam) + np.log(probs)[:, None]
            state['labels'] = orc.sample_discrete_from_log(lp, rng.random(n_sample))
            return float(lp.sum())
        return step
    o, c = w['o'], d + 1
    y = np.sin(x @ rng.standard_normal((d, o))) + 0.3 * rng.standard_normal((n_sample, o))
    bprior = (np.zeros((K, d)), 1e-2 * np.ones(K), np.stack(K * [1e2 * np.eye(d)]), (d + 1.0) * np.ones(K) + 1e-16)
    mprior = (np.zeros((K, o, c)), np.stack(K * [1e-2 * np.eye(c)]), np.stack(K * [1e1 * np.eye(o)]), (o + 1.0) * np.ones(K) + 1e-16)
    state = {'resp': rng.dirichlet(np.ones(K), size=n_sample).T}

    def step():
        r = state['resp']
        bpost = orc.nw_nat_to_std(orc.add_stats(orc.nw_std_to_nat(*bprior), orc.gauss_full_wstats(x, r)))
        mpost = orc.mnw_nat_to_std(orc.add_stats(orc.mnw_std_to_nat(*mprior), orc.lingauss_wstats(x, y, r)), tied=True)
        gp, dp = orc.stick_posterior(np.ones(K), 5.0 * np.ones(K), orc.categorical_wstats(r))
        ell = orc.nw_expected_loglik(x, *bpost) + orc.mnw_expected_loglik(x, y, *mpost) + orc.stick_expected_log(gp, dp)[0][:, None]
        state['resp'], lse = orc.responsibilities(ell)
        return float(lse.sum())
    return step


CPU_SAMPLE = {'cfg1': 2500, 'cfg2': 20000, 'cfg3': 2000, 'cfg4': 20000, 'cfg5': 512}


def time_cpu(w, name, steps, warmup):
    """the oracle port on the box's host cores: ALL of them -- torch.distributed.run exports OMP_NUM_THREADS=1 to its
    workers, so the BLAS pools are widened explicitly and the thread count that was really in use is what is reported."""
    n_s = min(CPU_SAMPLE[name], w['N'])
    fn = cpu_sweep_fn(w, n_s)
    cores = len(os.sched_getaffinity(0))
    used = cores
    try:
        from threadpoolctl import threadpool_limits, threadpool_info
        ctx = threadpool_limits(limits=cores)
    except Exception:                                        # no threadpoolctl: whatever the environment gives
        ctx, threadpool_info = None, None
    try:
        if threadpool_info is not None:
            pools = [p.get('num_threads', 1) for p in threadpool_info() if p.get('user_api') in ('blas', 'openmp')]
            used = max(pools) if pools else 1
        for _ in range(warmup):
            fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        dt = (time.perf_counter() - t0) / steps
    finally:
        if ctx is not None:
            ctx.restore_original_limits()
    return dict(value=n_s * w['K'] / dt, unit='points*components/s', cores=used, kind='port',
                sample='%d of %d points, all K=%d components, %d step(s), %.2f s/step; NumPy/OpenBLAS on %d threads (%d cores visible)'
                       % (n_s, w['N'], w['K'], steps, dt, used, cores)), dt


# ---------------------------------------------------------------------------------------
# one timed leg: W warm-up + K timed full steps of a session
# ---------------------------------------------------------------------------------------
class Ctx:
    """what every leg needs: torch, the engine, rank / world, the communicator."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        from mimo_b200 import _engine as E, _lib
        self.torch, self.dist, self.E, self._lib = torch, dist, E, _lib
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.dev = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, x):
        if self.world > 1:
            t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def gather_rows(self, row):
        """(world, len(row)) float64 array on every rank."""
        t = self.torch.tensor(row, device=self.dev, dtype=self.torch.float64)
        if self.world == 1:
            return t.cpu().numpy()[None]
        out = self.torch.empty((self.world, t.numel()), device=self.dev, dtype=self.torch.float64)
        self.dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy()


def one_step(s, hard, rng, phase=None, vlbs=None):
    from mimo_b200.distributions.bayesian import MEANFIELD, GIBBS
    if hard:
        var, gvar = s.draw_gibbs_variates('device')      # parameter variates from the device generator: no host read per sweep
        ops, outs = s.update_from_stats(GIBBS, variates=var, gating_variates=gvar)
        s.sweep(ops, hard=True, seed=int(rng.integers(1 << 30)), phase_ms=phase)
    else:
        ops, outs = s.update_from_stats(MEANFIELD)
        s.sweep(ops, hard=False, phase_ms=phase)
        v = s.lower_bound(outs)
        if vlbs is not None:
            vlbs.append(v)
    return outs


def time_leg(cx, s, hard, steps, warmup, tc_mode=None, sample_clocks=False, per_step=False):
    """W untimed + K timed full steps; CUDA events on the launching (current) stream, barrier + synchronize on both
    sides, max over ranks.  Returns ms per step, the per-phase device times of the timed steps and the lower bounds."""
    torch, E = cx.torch, cx.E
    old = E.set_tensor_cores(tc_mode) if tc_mode is not None else None
    rng = np.random.default_rng(7)
    try:
        outs = None
        for _ in range(warmup):
            outs = one_step(s, hard, rng)
        if outs is not None:
            s.check(outs)
        torch.cuda.synchronize()
        cx.barrier()
        sampler = ClockSampler(torch.cuda.current_device()) if (sample_clocks and cx.rank == 0) else None
        phase = np.zeros(6)
        vlbs = []
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        torch.cuda.synchronize()
        evs[0].record()
        for i in range(steps):
            outs = one_step(s, hard, rng, phase, vlbs)
            if per_step:
                evs[i + 1].record()
        if not per_step:
            evs[steps].record()
        torch.cuda.synchronize()
        cx.barrier()
        ms = cx.max_over_ranks(evs[0].elapsed_time(evs[steps]) / steps)
        clocks = sampler.stop() if sampler else None
        s.check(outs)
        out = dict(ms=ms, phase=phase, vlbs=vlbs, clocks=clocks, steps=steps, warmup=warmup)
        if per_step:
            out['ms_each'] = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        return out
    finally:
        if old is not None:
            E.set_tensor_cores(old)


def leg_roofline(w, n_local, leg, peaks, kernel_names):
    """roofline of the dominant phase of a leg from SURVEY 8(d)'s algorithmic work and the per-phase CUDA-event times."""
    work = algorithmic_work(w)
    steps = leg['steps']
    phase = leg['phase']
    phase_ms = phase[:3] / steps
    chunks = max(1.0, phase[4])
    pairs = n_local * w['K']
    flops = [work['e_flops_pair'] * pairs, 0.0, work['s_flops_pair'] * pairs]
    t_hbm = n_local * work['bytes_pt'] / (peaks['hbm'] * 1e9)
    t_tensor = sum(flops) / (peaks['tf_sus'] * 1e12)
    bound = 'tensor' if t_tensor >= t_hbm else 'hbm'
    dom = int(np.argmax(phase_ms))
    hard = w['mode'] == 'gibbs'
    launches = 1.0 if (hard and dom == 2) else chunks / steps
    ms = leg['ms']
    if bound == 'tensor':
        if flops[dom] <= 0:                       # the softmax phase dominates a tensor-bound config: rate the sweep
            ach = sum(flops) / (ms * 1e-3) / 1e12
        else:
            ach = flops[dom] / (phase_ms[dom] * 1e-3) / 1e12 if phase_ms[dom] > 0 else 0.0
        roof = dict(bound='tensor', achieved=ach, peak=peaks['tf_sus'], unit='TFLOP/s', frac=ach / peaks['tf_sus'], traffic=None)
        roof['whole_sweep_frac'] = (sum(flops) / (ms * 1e-3) / 1e12) / peaks['tf_sus']
        roof['per_kernel_frac'] = {nm: (f / (p * 1e-3) / 1e12) / peaks['tf_sus'] for nm, f, p in
                                   zip(('estep', 'softmax', 'stats'), flops, phase_ms) if f > 0 and p > 0}
    else:
        ach = n_local * work['bytes_pt'] / (phase_ms[dom] * 1e-3) / 1e9 if phase_ms[dom] > 0 else 0.0
        roof = dict(bound='hbm', achieved=ach, peak=peaks['hbm'], unit='GB/s', frac=ach / peaks['hbm'], traffic=None)
        roof['whole_sweep_frac'] = (n_local * work['bytes_pt'] / (ms * 1e-3) / 1e9) / peaks['hbm']
    roof.update(kernel=kernel_names[dom], launches_per_step=launches,
                ms_per_launch=phase_ms[dom] / max(launches, 1.0),
                phase_ms_per_step=dict(estep=phase_ms[0], softmax=phase_ms[1], stats=phase_ms[2]),
                t_tensor_ms=t_tensor * 1e3, t_hbm_ms=t_hbm * 1e3,
                peak_source='%s (sustained bf16 / copy bandwidth of MEASURED_PEAKS.json)' % peaks['src'],
                work='SURVEY 8(d): E-step %d flop/pair, statistics %.4g flop/pair, %d B/point'
                     % (work['e_flops_pair'], work['s_flops_pair'], work['bytes_pt']))
    return roof


def init_from_labels(cx, s, labels, K, comm):
    s.stat = cx.E.stats_hard(s.Z, labels, K, s.feats, 'fp32')
    if comm is not None:
        comm.allreduce(s.stat)


def screen_totals(cx, n_local, K):
    """totals of the last screened sweep over ALL its chunks, one row per rank."""
    t = cx.E.screen_totals()
    rows = cx.gather_rows([float(v) for v in t])
    out = []
    for r in rows:
        cand, pts, dense_chunks, chunks, level = r
        out.append(dict(chunks=int(chunks), dense_chunks=int(dense_chunks), tier_at_end=int(level),
                        candidate_pairs_per_point=(cand / pts) if pts > 0 else None,
                        candidate_fraction=(cand / (pts * K)) if pts > 0 else None))
    return out


# ---------------------------------------------------------------------------------------
# parity on a sub-sample of the resident data, against oracle/
# ---------------------------------------------------------------------------------------
def parity_subsample(cx, s, w, n_sub):
    """CUDA results on the first n_sub resident points (dense kernels and the screened default) against the oracle,
    with the model the benchmark is in (posterior from the session's all-reduced statistics).  max scaled errors:
    |gpu - ref|_max / max(1, |ref|_max) per quantity."""
    from oracle import mimo_oracle as orc
    from mimo_b200.distributions.bayesian import MEANFIELD
    torch, E = cx.torch, cx.E
    K, d = w['K'], w['d']
    n_sub = min(n_sub, s.N)
    t0 = time.perf_counter()
    stat_in = s.stat.clone()
    ops, outs = s.update_from_stats(MEANFIELD)
    s._gather_parameters(outs)
    s.check(outs)
    if cx.rank != 0:                       # the collective part is done; the check itself runs on rank 0's shard
        s.stat = stat_in
        return None
    Zs = s.Z[:n_sub].contiguous()
    x = Zs.double().cpu().numpy()

    def err(a, b):
        a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, float)
        return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))
    # ---- oracle: posterior from the statistics, expected log-joint, responsibilities, statistics of the sub-sample
    S = stat_in.cpu().numpy()
    il = np.tril_indices(d + 1)
    M = np.zeros((K, d + 1, d + 1))
    M[:, il[0], il[1]] = S
    M[:, il[1], il[0]] = S
    stats = [M[:, d, :d], M[:, d, d], M[:, :d, :d], M[:, d, d]]
    prior = (np.zeros((K, d)), 1e-2 * np.ones(K), np.stack(K * [np.eye(d)]), (d + 1.0) * np.ones(K) + 1e-8)
    post = orc.nw_nat_to_std(orc.add_stats(orc.nw_std_to_nat(*prior), stats))
    gp, dp = orc.stick_posterior(np.ones(K), 5.0 * np.ones(K), stats[1])
    ell = orc.nw_expected_loglik(x, *post) + orc.stick_expected_log(gp, dp)[0][:, None]
    resp, lse = orc.responsibilities(ell)
    st = orc.gauss_full_wstats(x, resp)
    ref_stat = np.zeros((K, d + 1, d + 1))
    ref_stat[:, :d, :d] = st[2]
    ref_stat[:, d, :d] = st[0]
    ref_stat[:, d, d] = st[1]
    ref_packed = ref_stat[:, il[0], il[1]]
    out = dict(points=n_sub, K=K, d=d, tolerance='rel 1e-4 (FP32 compute, FP64 accumulation); matrices norm-wise')
    po = outs['parts'][0]
    out['posterior'] = {k: err(po[k], r) for k, r in zip(('m', 'kappa', 'psi', 'nu'), post)}
    # ---- dense kernels: log-joint (K, n), log-normalisers, statistics
    r_t = E.empty((K, n_sub), torch.float32)            # a soft sweep's (K, n) output is the responsibilities
    lse_t = E.empty((n_sub,), torch.float32)
    old = E.set_tensor_cores(3)
    try:
        buf = E.SweepBuffers(n_sub, K, s.F, 'fp32', False)
        E.sweep(Zs, ops, s.feats, buf, ll_out=r_t, lse_out=lse_t)
        g_ll = E.loglik_tc(Zs, ops).double().cpu().numpy()          # the same dense 3-pass E-step kernel, stand-alone
        g_lse = lse_t.double().cpu().numpy()
        out['dense'] = dict(log_joint=err(g_ll, ell), lse=err(g_lse, lse), lse_sum_rel=abs(buf.lse_sum.item() - lse.sum()) / abs(lse.sum()),
                            resp=float(np.max(np.abs(r_t.double().cpu().numpy() - resp))), stats=err(buf.stat, ref_packed))
    finally:
        E.set_tensor_cores(old)
    # ---- the default (screened) path: statistics and the lower-bound data term (no (K, n) output on this path)
    buf = E.SweepBuffers(n_sub, K, s.F, 'fp32', False)
    E.sweep(Zs, ops, s.feats, buf)
    tot = E.screen_totals()
    out['screened'] = dict(lse_sum_rel=abs(buf.lse_sum.item() - lse.sum()) / abs(lse.sum()), stats=err(buf.stat, ref_packed),
                           dense_chunks=int(tot[2]), chunks=int(tot[3]))
    worst = max([out['dense'][k] for k in ('log_joint', 'lse', 'resp', 'stats')] + [out['screened']['stats']] + list(out['posterior'].values()))
    out['max_scaled_error'] = worst
    out['ok'] = bool(worst <= 1e-4)
    out['seconds'] = time.perf_counter() - t0
    s.stat = stat_in
    return out


# ---------------------------------------------------------------------------------------
# end to end: the full step through host buffers
# ---------------------------------------------------------------------------------------
def time_e2e(cx, w, s, hard, steps, comm, tc_mode=None, cache=None):
    """Every step: posterior update from the statistics of the previous step (batched posterior kernels on the device),
    operands device -> host, mimo_sweep_host on every rank (pinned host shard of Z -> device, one sweep, statistics +
    lower-bound scalar (+ labels) back to pinned host memory), the all-reduce of the statistics when sharded,
    statistics host -> device for the next update and the lower bound (mean field).  All inside the timed region; the
    step time is the max over ranks."""
    from mimo_b200.distributions.bayesian import MEANFIELD, GIBBS
    torch, E, _lib = cx.torch, cx.E, cx._lib
    N, D = s.Z.shape
    world = cx.world
    old = E.set_tensor_cores(tc_mode) if tc_mode is not None else None
    cache = cache if cache is not None else {}
    if 'zh' not in cache:                   # the pinned host copy of this rank's shard (made once, outside the timed region)
        cache['zh'] = torch.empty((N, D), dtype=torch.float32, pin_memory=True)
        cache['zh'].copy_(s.Z)
    zh = cache['zh']
    ops = s.ops(GIBBS if hard else MEANFIELD)
    a, b = (ops.W, None) if ops.family == 0 else (ops.S, ops.T)
    ah = torch.empty(a.shape, dtype=a.dtype, pin_memory=True)
    bh = torch.empty(b.shape, dtype=b.dtype, pin_memory=True) if b is not None else None
    ch = torch.empty(ops.cst.shape, dtype=ops.cst.dtype, pin_memory=True)
    flat_h = torch.zeros((s.K * s.F + 1,), dtype=torch.float64, pin_memory=True)      # statistics | sum of lse
    stat_h, lse_h = flat_h[:s.K * s.F], flat_h[s.K * s.F:]
    flat_d = torch.empty_like(flat_h, device=s.Z.device)
    lab_h = torch.empty((N,), dtype=torch.int32, pin_memory=True) if hard else None
    fi, fj = s.feats.fi_host, s.feats.fj_host
    rng = np.random.default_rng(11)
    stat_keep = s.stat.clone()
    torch.cuda.synchronize()
    times, vlbs = [], []
    try:
        for it in range(steps + 1):
            cx.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if hard:
                var, gvar = s.draw_gibbs_variates('device')
                _, outs = s.update_from_stats(GIBBS, variates=var, gating_variates=gvar)
            else:
                _, outs = s.update_from_stats(MEANFIELD)
            ah.copy_(a, non_blocking=True)
            if bh is not None:
                bh.copy_(b, non_blocking=True)
            ch.copy_(ops.cst, non_blocking=True)
            torch.cuda.synchronize()
            _lib.call('mimo_sweep_host', 0, ops.family, 1 if hard else 0, zh.data_ptr(), N, D,
                      ah.data_ptr(), bh.data_ptr() if bh is not None else None, ch.data_ptr(), ops.K, ops.Rp, ops.Dpp,
                      fi.ctypes.data, fj.ctypes.data, s.F, None, int(rng.integers(1 << 30)),
                      stat_h.data_ptr(), lse_h.data_ptr(), lab_h.data_ptr() if lab_h is not None else None)
            flat_d.copy_(flat_h, non_blocking=True)
            if world > 1:                                   # close the sweep: sum the shard statistics
                comm.allreduce(flat_d)
            s.stat.copy_(flat_d[:s.K * s.F].view(s.K, s.F))
            s.lse_sum = flat_d[s.K * s.F:]
            if not hard:
                vlbs.append(s.lower_bound(outs))            # device -> host read of the scalar
            else:
                torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        s.check(outs)
    finally:
        _lib.call('mimo_sweep_host_release')             # give the call's cached device buffers back
        if old is not None:
            E.set_tensor_cores(old)
        s.stat = stat_keep
    dt = cx.max_over_ranks(float(np.mean(times[1:])))
    ops_b = ah.numel() * 4 + (bh.numel() * 4 if bh is not None else 0) + ch.numel() * 4 + 2 * s.F * 4
    h2d = w['N'] * D * 4 + world * (ops_b + s.K * s.F * 8 + 8)
    d2h = world * (s.K * s.F * 8 + 8 + ops_b + 8) + (w['N'] * 4 if hard else 0)
    return dict(value=w['N'] * s.K / dt, unit='points*components/s', h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                ms_per_step=dt * 1e3, steps=steps, warmup=1, lower_bound=vlbs[-2:] if vlbs else None,
                call='posterior kernels -> operands to host -> mimo_sweep_host (C-ABI, pinned host buffers) on every rank -> '
                     'statistics back%s -> lower bound' % (' -> all-reduce' if world > 1 else ''))


# ---------------------------------------------------------------------------------------
# overlapping-components regime from a random start
# ---------------------------------------------------------------------------------------
def overlap_leg(cx, w, steps, comm_factory, peaks):
    """Same d, K, priors; OVERLAP_N points whose blob centres are OVERLAP_SPREAD sigma apart, model started from random
    responsibilities (device-side, mixtures/gmm.py:265-267): every sweep from the first one is timed, on the library's
    default path (which decides per chunk, on the device, between screening and the dense kernels)."""
    from mimo_b200.sharded import shard_bounds
    torch, E = cx.torch, cx.E
    wo = dict(w)
    wo['N'] = min(OVERLAP_N, w['N'])
    lo, hi = shard_bounds(wo['N'], cx.rank, cx.world)
    Z, _ = make_data(wo, hi - lo, lo, 4242, cx.dev, spread=OVERLAP_SPREAD)
    comm = comm_factory(wo['N'])
    model = build_model(wo)
    s = make_session(model, wo, Z, comm)
    s.stats_from_random_resp(seed=99)
    steps = max(2, min(steps, 6))
    leg = time_leg(cx, s, False, steps, 0, tc_mode=None, per_step=True)
    tot = screen_totals(cx, hi - lo, wo['K'])
    work = algorithmic_work(wo)
    pairs = (hi - lo) * wo['K']
    fl = (work['e_flops_pair'] + work['s_flops_pair']) * pairs
    out = dict(N=wo['N'], centre_spread_sigma=OVERLAP_SPREAD, init='random responsibilities (device)', steps=steps, warmup=0,
               ms_per_step=leg['ms'], ms_each_sweep=leg.get('ms_each'), value=wo['N'] * wo['K'] / (leg['ms'] * 1e-3),
               unit='points*components/s', tensor_frac=(fl / (leg['ms'] * 1e-3) / 1e12) / peaks['tf_sus'],
               screen_last_sweep=tot, lower_bound=leg['vlbs'])
    del s, Z
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------
def kernel_names(w, tcu, D):
    """names of the kernels behind the three phases of a leg (what the library dispatches to for this shape) and the bound
    that holds for the E-step kernel where it is not the algorithmic roofline."""
    d, K = w['d'], w['K']
    if w['kind'] == 'dgmm' and d <= 64 and K <= 256:
        return (['E-step + label draw: tc_diag_kernel (tcgen05 feature GEMM on [z, z^2], CTA pairs, labels in the epilogue)',
                 'softmax / label draw (fused into the E-step kernel)', 'hard statistics: block-local counting sort + diag_hard_stats_kernel'],
                'epilogue instruction issue (max / exp / cumulative sum per pair: 22 warp instructions per column at 2.1 per clock); '
                'the MMAs alone need 3072 clk per 256-point tile')
    if tcu and d > 64:
        return (['E-step: tc_estep2_kernel<2,128,3> (3 x FP16 split, CTA pairs, fused log-normaliser)', 'softmax / label draw',
                 'statistics: tc_fstats_kernel (feature GEMM over the folded triangle)'], None)
    if tcu and D < 24:
        return (['E-step: tc_estep2_kernel<1,Rp,3> (tcgen05, one 16-wide K step per 256 columns, 16 epilogue warps, fused log-normaliser)',
                 'softmax / label draw', 'statistics: tc_sstats_kernel (tcgen05 feature GEMM, whole packed triangle in one accumulator)'],
                'tensor-memory read port: K x Rp accumulator columns x 4 B per point at 64 B/clk/SM')
    return (['E-step (%s)' % ('tcgen05' if tcu else 'CUDA cores'), 'softmax / label draw', 'sufficient statistics'], None)


def run_config(cx, name, steps, warmup, comm_factory, peaks, n_override=0):
    """a short leg of one of the other BASELINE.json shapes on the library's default path."""
    from mimo_b200.sharded import shard_bounds
    torch, E = cx.torch, cx.E
    w = dict(WORKLOADS[name])
    if n_override:
        w['N'] = min(w['N'], n_override)
    lo, hi = shard_bounds(w['N'], cx.rank, cx.world)
    Z, labels = make_data(w, hi - lo, lo, 1337, cx.dev)
    comm = comm_factory(w['N'])
    model = build_model(w)
    s = make_session(model, w, Z, comm)
    init_from_labels(cx, s, labels, w['K'], comm)
    del labels
    hard = w['mode'] == 'gibbs'
    leg = time_leg(cx, s, hard, steps, warmup)
    tcu = E.sweep_uses_tensor_cores(s.ops(1 if hard else 0), Z.shape[1])
    names, true_bound = kernel_names(w, tcu, Z.shape[1])
    roof = leg_roofline(w, hi - lo, leg, peaks, names)
    if true_bound:
        roof['estep_bound'] = true_bound
    out = dict(workload='%s: %s' % (name, w['desc']), N=w['N'], K=w['K'], d=w['d'], sweep=w['mode'], steps=steps, warmup=warmup,
               ms_per_step=leg['ms'], value=w['N'] * w['K'] / (leg['ms'] * 1e-3), unit='points*components/s', roofline=roof,
               lower_bound=leg['vlbs'][-2:] if leg['vlbs'] else None, gpu_launches=int(leg['phase'][3]))
    if name == 'cfg1' and not hard and cx.world == 1:
        # at this size an iteration is ~30 launches of microseconds each: the same iteration replayed from a CUDA graph
        # (Session.capture_meanfield_step, what meanfield_coordinate_descent(graph=True) uses), lower bound read every step
        graph, outs = s.capture_meanfield_step()
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(steps, 20)
        e0.record()
        for _ in range(reps):
            graph.replay()
            v = s.lower_bound(outs)
        e1.record()
        torch.cuda.synchronize()
        gms = e0.elapsed_time(e1) / reps
        out['cuda_graph'] = dict(ms_per_step=gms, value=w['N'] * w['K'] / (gms * 1e-3), steps=reps, lower_bound=v)
    del s, Z
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--workload', default=os.environ.get('MIMO_BENCH_WORKLOAD', 'cfg5'))
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--n-override', type=int, default=0, help='debug only: smaller N (marks the line invalid)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-screened', action='store_true', help='skip the screened-path leg')
    ap.add_argument('--no-overlap', action='store_true', help='skip the overlapping-components regime')
    ap.add_argument('--no-others', action='store_true', help='skip the short cfg1-cfg4 legs')
    ap.add_argument('--no-parity', action='store_true', help='skip the oracle check on a sub-sample')
    ap.add_argument('--parity-points', type=int, default=2048)
    ap.add_argument('--tc-mode', type=int, default=-1, help='A/B only: force a tensor-core mode for the headline leg')
    args = ap.parse_args()
    name = args.workload
    w = dict(WORKLOADS[name])
    if args.n_override:
        w['N'] = args.n_override
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    warmup = max(3, args.warmup)
    config = dict(workload='%s: %s' % (name, w['desc']), N=w['N'], K=w['K'], d=w['d'], sweep=w['mode'],
                  inputs='resident FP32 data %.1f GB per sweep >> 126 MB L2 (no flush needed)' % (w['N'] * (w['d'] + w.get('o', 0)) * 4 / 1e9))
    if w['mode'] == 'gibbs':
        config['rng'] = 'labels: Philox4x32-10 keyed by the global point index; parameter variates: device generator (draw_gibbs_variates(\'device\'))'
    if args.n_override:
        config['INVALID'] = 'N overridden for debugging'

    if args.impl == 'reference':
        if rank != 0:
            return
        cpu, dt = time_cpu(w, name, max(1, args.steps), max(0, min(args.warmup, 1)))
        line = dict(impl='reference', metric='points*components/s per full sweep', value=cpu['value'], unit=cpu['unit'],
                    n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=dt * 1e3, higher_is_better=True,
                    scaling='strong', vs_baseline=None, dtype='f64', data='synthetic', config=config, cpu_baseline=cpu,
                    e2e=dict(value=cpu['value'], unit=cpu['unit'], h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import torch
    import __graft_entry__ as ge
    ge.build()
    from mimo_b200.sharded import Communicator, init_from_env, shard_bounds
    init_from_env()
    cx = Ctx()
    E = cx.E
    if os.environ.get('MIMO_TC_TRIANGULAR'):                 # A/B only: rows per step of the triangular skip (0 = kernel off)
        E.set_triangular(int(os.environ['MIMO_TC_TRIANGULAR']))
        config['tc_triangular'] = int(os.environ['MIMO_TC_TRIANGULAR'])
    if os.environ.get('MIMO_TC_QUAD_GEN'):                   # A/B only: 0 = the plain CTA-pair dense E-step
        E.set_quad_generations(int(os.environ['MIMO_TC_QUAD_GEN']))
        config['tc_quad_generations'] = int(os.environ['MIMO_TC_QUAD_GEN'])
    if os.environ.get('MIMO_TC_FLUSH_TILES'):                # A/B only: 128-point tiles between FP64 drains of the statistics
        E._lib.load().mimo_tc_set_flush_tiles(int(os.environ['MIMO_TC_FLUSH_TILES']))
        config['tc_flush_tiles'] = int(os.environ['MIMO_TC_FLUSH_TILES'])
    if world == 1:
        torch.cuda.set_device(0)
    cx.dev = dev = torch.device('cuda', torch.cuda.current_device())
    peaks = read_peaks()
    t_start = time.perf_counter()

    def comm_factory(n_global):
        return Communicator(N_global=n_global) if world > 1 else None

    lo, hi = shard_bounds(w['N'], rank, world)
    n_local = hi - lo
    Z, true_labels = make_data(w, n_local, lo, 1337, dev)
    comm = comm_factory(w['N'])
    model = build_model(w)
    s = make_session(model, w, Z, comm)
    K = w['K']
    hard = w['mode'] == 'gibbs'
    # the model the timed sweeps run in: posterior of the statistics of the generating labels (a converged model)
    init_from_labels(cx, s, true_labels, K, comm)
    del true_labels
    tcu = E.sweep_uses_tensor_cores(s.ops(1 if hard else 0), Z.shape[1])
    screenable = tcu and w['kind'] == 'gmm' and w['d'] >= 24 and K >= 32 and not hard
    head_mode = args.tc_mode if args.tc_mode >= 0 else (3 if screenable else None)
    if head_mode is not None:
        config['tc_mode'] = head_mode
    config['regime'] = ('dense: every pair through the 3-pass tcgen05 E-step and the tcgen05 statistics GEMM (mimo_set_tensor_cores(3))'
                        if head_mode == 3 else 'library default path')
    config['model'] = 'posterior of the statistics of the generating labels; blob centres %g sigma apart' % (3.0 if w['kind'] == 'ilr' else 4.0)

    # ---- headline leg ------------------------------------------------------------------------------------------
    stat0 = s.stat.clone()
    leg = time_leg(cx, s, hard, args.steps, warmup, tc_mode=head_mode, sample_clocks=True)
    ms = leg['ms']
    value = w['N'] * K / (ms * 1e-3)
    knames, true_bound = kernel_names(w, tcu, Z.shape[1])
    roof = leg_roofline(w, n_local, leg, peaks, knames)
    if true_bound:
        roof['estep_bound'] = true_bound
    if name == 'cfg5' and not args.n_override:
        dom = int(np.argmax(leg['phase'][:3]))
        tr = ncu_traffic(['tc_estep2_kernel', 'softmax_kernel', 'tc_fstats_kernel'][dom])
        if tr:
            roof['traffic'] = tr['bytes']
            roof['traffic_source'] = tr['source']
    launches = int(leg['phase'][3] + {'gmm': 4, 'dgmm': 3, 'ilr': 7}[w['kind']] * args.steps)

    # ---- the library's default (screened) path on the same data and model --------------------------------------
    screened = None
    if screenable and head_mode == 3 and not args.no_screened:
        s.stat = stat0.clone()
        sl = time_leg(cx, s, hard, args.steps, warmup, tc_mode=1)
        tot = screen_totals(cx, n_local, K)
        kms = sl['phase'][5] / sl['steps']                         # the screening kernel alone
        chunks = max(1.0, sl['phase'][4]) / sl['steps']
        rp = 1 << (max(w['d'], 8) - 1).bit_length()
        pairs = n_local * K
        mhz = (leg['clocks'] or {}).get('sm_mhz') or 1900.0
        floor_ms = pairs * min(rp, 32) * 4 / (64.0 * 148 * mhz * 1e6) * 1e3       # accumulator read-back, 64 B/clk/SM
        screened = dict(value=w['N'] * K / (sl['ms'] * 1e-3), unit='points*components/s', ms_per_step=sl['ms'], steps=sl['steps'],
                        warmup=sl['warmup'], per_rank=tot,
                        dominant_kernel=dict(kernel='tc_estep2_kernel<2,32,1> screening pass', ms_per_step=kms, launches_per_step=chunks,
                                             bound='TMEM read port (64 B/clk/SM)', floor_ms_per_step=floor_ms,
                                             frac_of_floor=floor_ms / kms if kms > 0 else None,
                                             executed_mma_tflops=2.0 * min(rp, 32) * w['d'] * pairs / (kms * 1e-3) / 1e12 if kms > 0 else None),
                        phase_ms_per_step=dict(estep=sl['phase'][0] / sl['steps'], softmax=sl['phase'][1] / sl['steps'],
                                               stats=sl['phase'][2] / sl['steps']),
                        lower_bound=sl['vlbs'][-2:], gpu_launches=int(sl['phase'][3]),
                        note='does not execute the SURVEY 8(d) flops: one FP16 pass over a 32-row orthogonal projection bounds '
                             'every log-joint, pairs within 40 nats of a point\'s best component are recomputed in FP32, statistics '
                             'are summed over those pairs; chunks with > 4 % candidates take the dense kernels (device-side)')

    # ---- parity of what was just timed, on a sub-sample against the oracle -------------------------------------
    parity = None
    if name == 'cfg5' and w['kind'] == 'gmm' and not args.no_parity:
        s.stat = stat0.clone()
        try:
            parity = parity_subsample(cx, s, w, args.parity_points)
        except Exception as ex:            # report, never fake
            if world > 1:
                raise
            parity = dict(error=str(ex)[:300])

    # ---- end to end --------------------------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        s.stat = stat0.clone()
        try:
            pinned = {}
            e2e = time_e2e(cx, w, s, hard, 3 if (ms > 2000) else min(max(args.steps, 3), 5), comm, tc_mode=head_mode, cache=pinned)
            if screened is not None:
                s.stat = stat0.clone()
                screened['e2e'] = time_e2e(cx, w, s, hard, min(max(args.steps, 3), 5), comm, tc_mode=1, cache=pinned)
            del pinned
        except Exception as ex:
            if world > 1:
                raise
            e2e = dict(value=None, unit='points*components/s', error=str(ex)[:300])
    del s, Z, stat0
    torch.cuda.empty_cache()

    # ---- overlapping components, random start ------------------------------------------------------------------
    overlap = None
    if screenable and not args.no_overlap:
        overlap = overlap_leg(cx, w, args.steps, comm_factory, peaks)

    # ---- the other BASELINE.json shapes, short legs ------------------------------------------------------------
    others = None
    if name == 'cfg5' and not args.no_others and not args.n_override:
        others = {}
        for nm in ('cfg1', 'cfg2', 'cfg3', 'cfg4'):
            try:
                others[nm] = run_config(cx, nm, min(args.steps, 5), 3, comm_factory, peaks)
            except Exception as ex:
                if world > 1:
                    raise
                others[nm] = dict(error=str(ex)[:300])
    if rank != 0:
        return

    cpu = None
    if not args.no_cpu and world == 1:
        cpu, _ = time_cpu(w, name, 3, 1 if CPU_SAMPLE[name] * K < 5e6 else 0)      # ~10-20 s of host work on the bounded sample
    line = dict(metric='points*components/s per full sweep', value=value, unit='points*components/s', n_gpus=world,
                steps=args.steps, warmup=warmup, ms_per_step=ms, higher_is_better=True, scaling='strong',
                vs_baseline=None, dtype='f32', data='synthetic', config=config, roofline=roof, cpu_baseline=cpu, e2e=e2e,
                gpu_launches=launches, clocks=leg['clocks'], lower_bound=leg['vlbs'][-3:] if leg['vlbs'] else None,
                parity_subsample=parity, screened_path=screened, overlap_regime=overlap, other_configs=others,
                comm=dict(messages=comm.messages, bytes_per_message=comm.bytes // max(comm.messages, 1)) if comm else None,
                bench_seconds=time.perf_counter() - t_start)
    print(json.dumps(line))


if __name__ == '__main__':
    try:
        main()
    finally:
        try:
            import torch.distributed as _d
            if _d.is_available() and _d.is_initialized():
                _d.destroy_process_group()
        except Exception:
            pass
