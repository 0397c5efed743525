#!/usr/bin/env python
"""bench.py -- points*components/s of one full inference sweep (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5] [--impl reference]

A "step" is one full sweep of the named workload over synthetic data of its shape:
  mean-field : batched posterior kernels (statistics -> posteriors, operands, lower-bound
               terms) -> fused E-step + statistics sweep -> [all-reduce] -> lower bound read
  Gibbs      : host-drawn parameter variates -> posterior kernels (draw) -> fused E-step +
               label draw + statistics sweep -> [all-reduce]
`value` = N*K / (device time per step), data resident in HBM (inputs larger than L2).
`e2e`   = the same metric through the host-buffer C-ABI call (mimo_sweep_host): pinned host
          data -> device, one sweep, statistics + lower-bound scalar back, every step.
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
import threading
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
    """DRAM bytes (read + write) per launch of the dominant kernel, from the committed `ncu --set full` capture of
    the same kernel on one point chunk of this workload (profiles/r01_ncu_traffic.json); None when not captured."""
    p = os.path.join(ROOT, 'profiles', 'r01_ncu_traffic.json')
    try:
        return json.load(open(p)).get(kernel)
    except Exception:
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
def make_data(w, n_local, lo, seed, dev):
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
    """returns (step_fn, N_sample, K): one reference-algorithm sweep on n_sample points."""
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
    n_s = min(CPU_SAMPLE[name], w['N'])
    fn = cpu_sweep_fn(w, n_s)
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps
    cores = len(os.sched_getaffinity(0))
    return dict(value=n_s * w['K'] / dt, unit='points*components/s', cores=cores, kind='port',
                sample='%d of %d points, all K=%d components, %d step(s), %.2f s/step; NumPy/OpenBLAS on %d threads'
                       % (n_s, w['N'], w['K'], steps, dt, cores)), dt


# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--workload', default=os.environ.get('MIMO_BENCH_WORKLOAD', 'cfg5'))
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--n-override', type=int, default=0, help='debug only: smaller N (marks the line invalid)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-dense', action='store_true', help='skip the dense-path comparison sweep')
    ap.add_argument('--tc-mode', type=int, default=-1, help='A/B only: 3 = dense 3-pass E-step (no screening), 0 = CUDA cores')
    args = ap.parse_args()
    name = args.workload
    w = dict(WORKLOADS[name])
    if args.n_override:
        w['N'] = args.n_override
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    config = dict(workload='%s: %s' % (name, w['desc']), N=w['N'], K=w['K'], d=w['d'], sweep=w['mode'],
                  inputs='resident FP32 data %.1f GB per sweep >> 126 MB L2 (no flush needed)' % (w['N'] * (w['d'] + w.get('o', 0)) * 4 / 1e9))
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
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    from mimo_b200 import _engine as E, _lib
    from mimo_b200.sharded import Communicator, init_from_env, shard_bounds
    from mimo_b200.distributions.bayesian import MEANFIELD, GIBBS
    init_from_env()
    if args.tc_mode >= 0:
        E.set_tensor_cores(args.tc_mode)
        config['tc_mode'] = args.tc_mode
    if world == 1:
        torch.cuda.set_device(0)
    dev = torch.device('cuda', torch.cuda.current_device())
    lo, hi = shard_bounds(w['N'], rank, world)
    n_local = hi - lo
    Z, true_labels = make_data(w, n_local, lo, 1337, dev)
    comm = Communicator(N_global=w['N']) if world > 1 else None
    model = build_model(w)
    s = make_session(model, w, Z, comm)
    K = w['K']
    hard = w['mode'] == 'gibbs'
    rng = np.random.default_rng(7)

    # initial statistics from the generating labels (one-hot), then the sweep loop
    s.stat = E.stats_hard(s.Z, true_labels, K, s.feats, 'fp32')
    if comm is not None:
        comm.allreduce(s.stat)
    del true_labels
    phase = np.zeros(6)
    vlbs = []

    def step(timed):
        if hard:
            var, gvar = s.draw_gibbs_variates()
            ops, outs = s.update_from_stats(GIBBS, variates=var, gating_variates=gvar)
            s.sweep(ops, hard=True, seed=int(rng.integers(1 << 30)), phase_ms=phase if timed else None)
        else:
            ops, outs = s.update_from_stats(MEANFIELD)
            s.sweep(ops, hard=False, phase_ms=phase if timed else None)
            vlbs.append(s.lower_bound(outs))
        return outs

    for _ in range(args.warmup):
        outs = step(False)
    s.check(outs)
    torch.cuda.synchronize()
    if os.environ.get('MIMO_BENCH_DEBUG') and not hard:
        for rep in range(2):
            t = [time.perf_counter()]
            ops, outs = s.update_from_stats(MEANFIELD); torch.cuda.synchronize(); t.append(time.perf_counter())
            s.sweep(ops, hard=False); torch.cuda.synchronize(); t.append(time.perf_counter())
            s.lower_bound(outs); t.append(time.perf_counter())
            ph = np.zeros(6)
            s.sweep(ops, hard=False, phase_ms=ph); torch.cuda.synchronize(); t.append(time.perf_counter())
            sys.stderr.write('DEBUG update %.2f ms | sweep %.2f ms | vlb %.2f ms | timed sweep %.2f ms phases %s\n'
                             % tuple([1e3 * (t[i + 1] - t[i]) for i in range(4)] + [ph.tolist()]))
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        outs = step(True)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if sampler else None
    s.check(outs)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    screened = (args.tc_mode in (-1, 1, 4)) and w['kind'] == 'gmm' and w['d'] >= 24 and K >= 32 and not hard
    screen = None
    if rank == 0 and screened and E.sweep_uses_tensor_cores(s.ops(0), w['d']):
        cands, fell_back = E.screen_last()
        screen = dict(last_chunk_candidate_pairs=cands, last_chunk_dense_fallback=bool(fell_back))
    # end-to-end through the host-buffer C-ABI call on every rank: pinned host shard in, statistics out (+ the
    # all-reduce of the statistics when sharded), every step
    e2e = None
    if not args.no_e2e:
        try:
            e2e = time_e2e(w, s, E, _lib, hard, min(args.steps, 2), comm)
        except Exception as ex:   # report, never fake
            if world > 1:
                raise
            e2e = dict(value=None, unit='points*components/s', error=str(ex)[:200])
    if rank != 0:
        return

    peaks = read_peaks()
    work = algorithmic_work(w)
    value = w['N'] * K / (ms * 1e-3)
    # roofline of the dominant kernel, from the per-phase CUDA-event times of the timed steps
    chunks = max(1.0, phase[4])          # point chunks over the timed steps = launches of each per-chunk kernel
    phase_ms = phase[:3] / args.steps
    dom = int(np.argmax(phase_ms))
    pairs_local = n_local * K
    t_hbm = n_local * work['bytes_pt'] / (peaks['hbm'] * 1e9)
    flops = [work['e_flops_pair'] * pairs_local, 0.0, work['s_flops_pair'] * pairs_local]
    t_tensor = sum(flops) / (peaks['tf_sus'] * 1e12)
    bound = 'tensor' if t_tensor >= t_hbm else 'hbm'
    launches_per_step = 1.0 if (hard and dom == 2) else chunks / args.steps
    if bound == 'tensor':
        ach = flops[dom] / (phase_ms[dom] * 1e-3) / 1e12 if phase_ms[dom] > 0 else 0.0
        roof = dict(bound='tensor', achieved=ach, peak=peaks['tf_sus'], unit='TFLOP/s', frac=ach / peaks['tf_sus'], traffic=None)
    else:
        ach = n_local * work['bytes_pt'] / (phase_ms[dom] * 1e-3) / 1e9 if phase_ms[dom] > 0 else 0.0
        roof = dict(bound='hbm', achieved=ach, peak=peaks['hbm'], unit='GB/s', frac=ach / peaks['hbm'], traffic=None)
    if name == 'cfg5' and not args.n_override:
        tr = ncu_traffic([('tc_estep2_screen_kernel' if screened else 'tc_estep2_kernel'), 'softmax_kernel',
                          ('pair_stats_kernel' if screened else 'tc_fstats_kernel')][dom])
        if tr:
            roof['traffic'] = tr['bytes']                     # DRAM bytes per launch (one ~1M-point chunk)
            roof['traffic_source'] = tr['source']
    if bound == 'tensor' and w['kind'] == 'gmm' and w['d'] >= 24:
        # what the tensor pipe really executes for the E-step: 3 FP16 passes over all Rp operand rows on the dense path;
        # 1 pass over the 32 projected rows on the screened path (plus FP32 CUDA-core work on the candidate lists)
        rp = 1 << (max(w['d'], 8) - 1).bit_length()
        mma_pair = (2.0 * min(rp, 32) * w['d']) if screened else (3 * 2.0 * rp * w['d'])
        if phase_ms[0] > 0:
            roof['estep_executed_tflops'] = mma_pair * pairs_local / (phase_ms[0] * 1e-3) / 1e12
        roof['note'] = ('achieved = ALGORITHMIC flops of the dense formulation (SURVEY 8d) / time; the default path is SCREENED: one '
                        'FP16 pass over a 32-row orthogonal projection of every operand bounds all N*K log-joints, pairs within 40 '
                        'nats of a point\'s best component are recomputed in FP32, statistics are summed over those pairs only, so '
                        'frac can exceed 1; `dense_path` is the same sweep with screening off (3-pass dense tensor-core kernels)'
                        if screened else 'dense 3-pass tensor-core path')
    if dom == 0 and bound == 'tensor' and phase[5] > 0:
        # the dominant KERNEL of the E-step phase (on the screened path: the screening pass; the rest of the phase is list
        # building and refinement): algorithmic E-step flops / its own launch time
        kms = phase[5] / args.steps
        roof['achieved'] = flops[0] / (kms * 1e-3) / 1e12
        roof['frac'] = roof['achieved'] / peaks['tf_sus']
        roof['kernel_ms_per_step'] = kms
    roof.update(kernel=[('E-step: tc_estep2_kernel<.,32,1> screening pass' if screened else 'E-step (log-likelihood)'),
                        'softmax / label draw', 'sufficient statistics'][dom],
                launches_per_step=launches_per_step,
                ms_per_launch=(phase[5] / args.steps if (dom == 0 and phase[5] > 0) else phase_ms[dom]) / max(launches_per_step, 1),
                phase_ms_per_step=dict(estep=phase_ms[0], softmax=phase_ms[1], stats=phase_ms[2]),
                peak_source='%s (sustained bf16 / copy bandwidth of MEASURED_PEAKS.json)' % peaks['src'],
                whole_sweep_frac=(sum(flops) / (ms * 1e-3) / 1e12) / peaks['tf_sus'] if bound == 'tensor'
                else (n_local * work['bytes_pt'] / (ms * 1e-3) / 1e9) / peaks['hbm'])

    cpu = None
    if not args.no_cpu and world == 1:
        cpu, _ = time_cpu(w, name, 1, 1 if CPU_SAMPLE[name] * K < 5e6 else 0)
    # the same sweep with the screening off (dense 3-pass E-step + dense tensor-core statistics): what overlapping
    # components would cost; one warm-up + one timed sweep
    dense = None
    vlb_tail = vlbs[-3:] if vlbs else None
    if world == 1 and screened and E.sweep_uses_tensor_cores(s.ops(0), w['d']):
        if not args.no_dense:
            old = E.set_tensor_cores(3)
            try:
                step(False)
                torch.cuda.synchronize()
                d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                d0.record()
                step(False)
                d1.record()
                torch.cuda.synchronize()
                dms = d0.elapsed_time(d1)
                dense = dict(value=w['N'] * K / (dms * 1e-3), unit='points*components/s', ms_per_step=dms, steps=1, warmup=1,
                             tensor_frac=(sum(flops) / (dms * 1e-3) / 1e12) / peaks['tf_sus'],
                             what='mimo_set_tensor_cores(3): dense 3-pass E-step and dense statistics on every chunk')
            finally:
                E.set_tensor_cores(old)
    posterior_launches = {'gmm': 4, 'dgmm': 3, 'ilr': 7}[w['kind']]
    line = dict(metric='points*components/s per full sweep', value=value, unit='points*components/s', n_gpus=world,
                steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling='strong',
                vs_baseline=None, dtype='f32', data='synthetic', config=config, roofline=roof, cpu_baseline=cpu, e2e=e2e,
                gpu_launches=int(phase[3] + posterior_launches * args.steps), clocks=clocks,
                lower_bound=vlb_tail, screen=screen, dense_path=dense,
                comm=dict(messages=comm.messages, bytes_per_message=comm.bytes // max(comm.messages, 1)) if comm else None)
    print(json.dumps(line))


def time_e2e(w, s, E, _lib, hard, steps, comm=None):
    """mimo_sweep_host on every rank: pinned host shard of Z -> device, one sweep with the current operands,
    statistics + lower-bound scalar (+ labels) back to the host; when sharded, the all-reduce of the statistics
    closes the step.  Everything inside the timed region; the step time is the max over ranks."""
    import torch
    import torch.distributed as dist
    N, D = s.Z.shape
    world = comm.world if comm is not None else 1
    ops = s.ops(1 if hard else 0)
    zh = torch.empty((N, D), dtype=torch.float32, pin_memory=True)
    zh.copy_(s.Z)
    a, b = (ops.W, None) if ops.family == 0 else (ops.S, ops.T)
    ah = a.cpu().contiguous()
    bh = b.cpu().contiguous() if b is not None else None
    ch = ops.cst.cpu().contiguous()
    flat_h = torch.zeros((s.K * s.F + 1,), dtype=torch.float64, pin_memory=True)      # statistics | sum of lse
    stat_h, lse_h = flat_h[:s.K * s.F], flat_h[s.K * s.F:]
    flat_d = torch.empty_like(flat_h, device=s.Z.device) if world > 1 else None
    lab_h = torch.empty((N,), dtype=torch.int32, pin_memory=True) if hard else None
    fi, fj = s.feats.fi_host, s.feats.fj_host
    torch.cuda.synchronize()
    times = []
    for it in range(steps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.call('mimo_sweep_host', 0, ops.family, 1 if hard else 0, zh.data_ptr(), N, D,
                  ah.data_ptr(), bh.data_ptr() if bh is not None else None, ch.data_ptr(), ops.K, ops.Rp, ops.Dpp,
                  fi.ctypes.data, fj.ctypes.data, s.F, None, 12345 + it,
                  stat_h.data_ptr(), lse_h.data_ptr(), lab_h.data_ptr() if lab_h is not None else None)
        if world > 1:                                   # close the sweep: sum the shard statistics
            flat_d.copy_(flat_h, non_blocking=True)
            comm.allreduce(flat_d)
            flat_h.copy_(flat_d, non_blocking=True)
            torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    _lib.call('mimo_sweep_host_release')             # give the call's cached device buffers back
    dt = float(np.mean(times[1:]))
    if world > 1:
        t = torch.tensor([dt], device=s.Z.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    ops_b = ah.numel() * 4 + (bh.numel() * 4 if bh is not None else 0) + ch.numel() * 4 + 2 * s.F * 4
    h2d = w['N'] * D * 4 + world * ops_b + (world * s.K * s.F * 8 if world > 1 else 0)
    d2h = world * (s.K * s.F * 8 + 8) * (2 if world > 1 else 1) + (w['N'] * 4 if hard else 0)
    return dict(value=w['N'] * s.K / dt, unit='points*components/s', h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                ms_per_step=dt * 1e3, call='mimo_sweep_host (C-ABI, pinned host buffers) on every rank'
                                           + (' + all-reduce of the statistics' if world > 1 else ''))


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
