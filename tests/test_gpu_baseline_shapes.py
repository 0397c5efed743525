"""Parity at the BASELINE.json shapes (K, d of cfg2-cfg5; N cut to what the CPU oracle finishes in seconds):
the CUDA path through the C-ABI against oracle/ on the same seeded inputs.  GPU only.

  cfg5  K=1024 d=128  full covariance, mean field: dense 3-pass kernels, screening tier 0 (32-row projection)
                      and tier 1 (all rows, one pass)
  cfg3  K=256  d=64   diagonal covariance, Gibbs labels from supplied uniforms + hard statistics + NG posterior
  cfg2  K=128  d_in=8 o=1  tied matrix-normal-Wishart experts, stick-breaking, two mean-field iterations
  cfg4  K=64   d=16   stick-breaking DP-GMM, two mean-field iterations
plus direct calls of the public per-phase methods on the toy fixtures made from the reference.

Tolerances (north_star): rel 1e-4 for FP32 compute / FP64 accumulation (matrices norm-wise); labels bit-exact away
from a CDF boundary (band 1e-3 in FP32 mode).
"""
import os

import numpy as np
import numpy.random as npr
import pytest
import torch

from oracle import mimo_oracle as orc
from test_gpu_kernels import close, eng, spd, unpack_quad

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def load(name):
    return dict(np.load(os.path.join(GOLD, name + '.npz')))


def _packed(st, d):
    """oracle statistics (sum r x, sum r, sum r xx^T, sum r) -> packed lower triangle of zt zt^T (K, F)."""
    K = st[1].shape[0]
    S = np.zeros((K, d + 1, d + 1))
    S[:, :d, :d] = st[2]
    S[:, d, :d] = st[0]
    S[:, d, d] = st[1]
    il = np.tril_indices(d + 1)
    return S[:, il[0], il[1]]


# ------------------------------------------------------------------------------------------------------------
# cfg5 shape
# ------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def cfg5_case():
    """K = 1024 Normal-Wishart posteriors in d = 128 and 3072 points drawn around their means: tight, well separated
    components (what a converged model of the benchmark data looks like), so the screening tiers have something to
    screen.  The oracle results are computed once."""
    K, d, N = 1024, 128, 3072
    rng = np.random.default_rng(2024)
    centres = 4.0 * rng.standard_normal((K, d))
    z = rng.integers(0, K, size=N)
    x = centres[z] + rng.standard_normal((N, d))
    mus = centres + 0.02 * rng.standard_normal((K, d))
    kappas = 50.0 + 10 * rng.random(K)
    nus = d + 50.0 + 10 * rng.random(K)
    base = spd(rng, d)
    base = base / np.trace(base) * d
    psis = np.stack([(base * (0.8 + 0.4 * rng.random()) + 0.2 * np.diag(rng.random(d))) / nus[k] for k in range(K)])
    gam, dlt = 1.0 + 3.0 * rng.random(K), 5.0 + 3.0 * rng.random(K)
    x32 = x.astype(np.float32).astype(np.float64)
    ell = orc.nw_expected_loglik_blas(x32, mus, kappas, psis, nus) + orc.stick_expected_log(gam, dlt)[0][:, None]
    resp, lse = orc.responsibilities(ell)
    st = orc.gauss_full_wstats_blas(x32, resp)
    return dict(K=K, d=d, N=N, x=x32, post=(mus, kappas, psis, nus), stick=(gam, dlt), ell=ell, resp=resp, lse=lse,
                stat=_packed(st, d))


def _cfg5_operands(E, c):
    K, d = c['K'], c['d']
    feats = E.quad_features(d)
    ops = E.QuadOperands(K, d, d, 'fp32')
    idx = E.identity_map(d, d)
    zero = E.zeros((K, feats.F))
    gam, dlt = c['stick']
    E.gating_posterior(1, E.to_dev(gam), E.to_dev(dlt), zero, feats.F, feats.F - 1, mode=0, ops=ops)['info'].check()
    E.nw_posterior([E.to_dev(p) for p in c['post']], zero, feats.F, idx, d + 1, mode=0, ops=ops, col_map=idx)['info'].check()
    return ops, feats


@pytest.mark.parametrize('mode,what', [(3, 'dense'), (1, 'tier0'), (5, 'tier1')])
def test_cfg5_shape_sweep_against_oracle(cfg5_case, mode, what):
    E = eng()
    c = cfg5_case
    K, d, N = c['K'], c['d'], c['N']
    ops, feats = _cfg5_operands(E, c)
    Z = E.to_dev(c['x'], torch.float32)
    old = E.set_tensor_cores(mode)
    try:
        buf = E.SweepBuffers(N, K, feats.F, 'fp32', False)
        if mode == 3:
            close(E.loglik_tc(Z, ops), c['ell'], 1e-4, 'expected log-joint (dense 3-pass kernel, K=1024 d=128)')
            r = E.empty((K, N), torch.float32)          # a soft sweep's (K, N) output is the responsibilities
            lse_t = E.empty((N,), torch.float32)
            E.sweep(Z, ops, feats, buf, ll_out=r, lse_out=lse_t)
            close(lse_t, c['lse'], 1e-4, 'log-normalisers (dense)')
            err = float(np.max(np.abs(r.double().cpu().numpy() - c['resp'])))
            assert err <= 1e-4, 'responsibilities (dense): max abs err %.2e' % err
        else:
            E.sweep(Z, ops, feats, buf)
            cand, pts, dense_chunks, chunks, level = E.screen_totals()
            print('%s: %.2f candidate pairs per point, %d of %d chunks dense, ended on tier %d'
                  % (what, cand / max(pts, 1), dense_chunks, chunks, level))
            assert chunks >= 1
            if mode == 5:
                assert dense_chunks == 0 and level == 1, 'the all-rows screening pass should have handled this chunk'
            else:
                # the 32-row projection is too loose a bound for components 4 sigma apart in d = 128: the chunk takes the
                # dense kernels (device-side decision) and the sweep moves one tier up; either way the results must hold
                assert (dense_chunks == 0 and level == 0) or (dense_chunks == chunks and level == 1)
        close(buf.stat, c['stat'], 1e-4, 'statistics (%s)' % what)
        assert abs(buf.lse_sum.item() - c['lse'].sum()) <= 1e-6 * abs(c['lse'].sum()), 'sum of log-normalisers (%s)' % what
        assert abs(buf.stat.cpu().numpy()[:, -1].sum() - N) <= 1e-6 * N
    finally:
        E.set_tensor_cores(old)


def test_cfg5_shape_posterior_against_oracle(cfg5_case):
    """statistics -> Normal-Wishart posterior at K = 1024, d = 128 (FP64 kernel, one CTA per component)."""
    E = eng()
    c = cfg5_case
    K, d = c['K'], c['d']
    feats = E.quad_features(d)
    idx = E.identity_map(d, d)
    prior = (np.zeros((K, d)), 1e-2 * np.ones(K), np.stack(K * [np.eye(d)]), (d + 1.0) * np.ones(K) + 1e-8)
    S = unpack_quad(c['stat'], d)
    st = [S[:, d, :d], S[:, d, d], S[:, :d, :d], S[:, d, d]]
    ref = orc.nw_nat_to_std(orc.add_stats(orc.nw_std_to_nat(*prior), st))
    out = E.nw_posterior([E.to_dev(p) for p in prior], E.to_dev(c['stat']), feats.F, idx, d + 1, mode=0)
    out['info'].check()
    for key, r in zip(('m', 'kappa', 'psi', 'nu'), ref):
        close(out[key], r, 1e-9, 'NW posterior ' + key)
    close(out['vlb'], orc.nw_vlb(prior, ref), 1e-8, 'NW lower-bound terms')


# ------------------------------------------------------------------------------------------------------------
# cfg3 shape: diagonal Gibbs
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
def test_cfg3_shape_diag_gibbs(precision):
    E = eng()
    K, d, N = 256, 64, 20000
    rng = np.random.default_rng(33)
    centres = 4.0 * rng.standard_normal((K, d))
    sig = 0.5 + rng.random((K, d))
    z = rng.integers(0, K, size=N)
    x = centres[z] + rng.standard_normal((N, d)) * sig[z]
    mus = centres + 0.05 * rng.standard_normal((K, d))
    lam = 1.0 / sig ** 2 * (0.9 + 0.2 * rng.random((K, d)))
    logw = np.log(rng.dirichlet(np.ones(K)))
    tol = {'fp32': 1e-4, 'fp64': 1e-9}[precision]
    Z = E.to_dev(x, E.tdtype(precision))
    xr = Z.double().cpu().numpy()
    ops = E.DiagOperands(K, d, precision)
    E.operands_gauss_diag(ops, E.to_dev(mus), E.to_dev(lam))
    ops.cst += E.to_dev(logw, ops.cst.dtype)
    feats = E.diag_features(d)
    u = rng.random(N)
    buf = E.SweepBuffers(N, K, feats.F, precision, True)
    ll = E.empty((K, N), E.tdtype(precision))
    E.sweep(Z, ops, feats, buf, uniforms=E.to_dev(u), ll_out=ll)
    ref_ll = orc.gauss_diag_loglik(xr, mus, lam) + logw[:, None]
    close(ll, ref_ll, tol, 'diag log-joint K=256 d=64')
    lab_ref = orc.sample_discrete_from_log(ref_ll, u)
    safe = orc.label_boundary_distance(ref_ll, u) > (1e-6 if precision == 'fp64' else 1e-3)
    lab = buf.labels.cpu().numpy()
    assert safe.mean() > 0.99
    assert np.array_equal(lab[safe], lab_ref[safe]), 'labels (%d mismatches away from a CDF boundary)' % int((lab[safe] != lab_ref[safe]).sum())
    # hard statistics of the labels the kernel drew (reference: one_hot + weighted_statistics, gaussian.py:819-832)
    st = orc.gauss_diag_wstats(xr, orc.one_hot(lab, K))
    got = buf.stat.cpu().numpy()
    close(got[:, :d], st[0], tol, 'sum x by label')
    close(got[:, d:2 * d], st[3], tol, 'sum x^2 by label')
    assert np.array_equal(got[:, 2 * d], np.bincount(lab, minlength=K))
    # Normal-Gamma posterior from them
    prior = (np.zeros((K, d)), 1e-2 * np.ones((K, d)), (3.0 + 1e-8) / 2 * np.ones((K, d)), 0.5 * np.ones((K, d)))
    ref = orc.ng_nat_to_std(orc.add_stats(orc.ng_std_to_nat(*prior), st))
    out = E.ng_posterior([E.to_dev(p) for p in prior], buf.stat, feats.F, mode=0)
    out['info'].check()
    for key, r in zip(('m', 'kappa', 'alpha', 'beta'), ref):
        close(out[key], r, 10 * tol, 'NG posterior ' + key)


# ------------------------------------------------------------------------------------------------------------
# cfg4 / cfg2 shapes: two mean-field iterations through the session the public drivers use
# ------------------------------------------------------------------------------------------------------------
def _bench():
    import bench
    return bench


@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
def test_cfg4_shape_two_vi_iterations(precision):
    """K = 64, d = 16, stick-breaking: random responsibilities -> posterior -> E-step -> posterior, against the
    oracle chain (mixtures/gmm.py:261-297)."""
    import mimo_b200
    from mimo_b200.distributions.bayesian import MEANFIELD
    b = _bench()
    E = eng()
    K, d, N = 64, 16, 20000
    w = dict(b.WORKLOADS['cfg4'], N=N)
    rng = np.random.default_rng(4)
    centres = 3.0 * rng.standard_normal((K, d))
    z = rng.integers(0, K, size=N)
    x = centres[z] + rng.standard_normal((N, d)) @ np.linalg.cholesky(spd(rng, d)).T
    tol = {'fp32': 1e-4, 'fp64': 1e-9}[precision]
    mimo_b200.set_default_precision(precision)
    try:
        model = b.build_model(w, precision=precision)
        s = model._session(x)
        xr = s.Z.double().cpu().numpy()
        resp = rng.dirichlet(np.ones(K), size=N).T
        prior = (np.zeros((K, d)), 1e-2 * np.ones(K), np.stack(K * [np.eye(d)]), (d + 1.0) * np.ones(K) + 1e-8)
        g0 = (np.ones(K), 5.0 * np.ones(K))
        s.stats_from_resp(resp)
        for it in range(2):
            post = orc.nw_nat_to_std(orc.add_stats(orc.nw_std_to_nat(*prior), orc.gauss_full_wstats(xr, resp)))
            gp = orc.stick_posterior(g0[0], g0[1], orc.categorical_wstats(resp))
            ell = orc.nw_expected_loglik(xr, *post) + orc.stick_expected_log(*gp)[0][:, None]
            resp, lse = orc.responsibilities(ell)
            vlb = orc.stick_vlb(g0, gp) + np.sum(orc.nw_vlb(prior, post)) + lse.sum()
            ops, outs = s.update_from_stats(MEANFIELD)
            s.check(outs)
            for key, r in zip(('m', 'kappa', 'psi', 'nu'), post):
                close(outs['parts'][0][key], r, 10 * tol, 'iteration %d NW posterior %s' % (it, key))
            close(outs['gating']['a'], gp[0], 10 * tol, 'stick gammas')
            close(outs['gating']['b'], gp[1], 10 * tol, 'stick deltas')
            s.sweep(ops, hard=False)
            close(s.stat, _packed(orc.gauss_full_wstats(xr, resp), d), 10 * tol, 'iteration %d statistics' % it)
            got = s.lower_bound(outs)
            assert abs(got - vlb) <= 10 * tol * abs(vlb), 'iteration %d lower bound %.10g vs %.10g' % (it, got, vlb)
    finally:
        mimo_b200.set_default_precision('fp32')


@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
def test_cfg2_shape_ilr_two_vi_iterations(precision):
    """K = 128, d_in = 8, d_out = 1, TIED matrix-normal-Wishart experts + Normal-Wishart basis + stick-breaking
    (examples/ilr/evaluate_sine.py:88-119): posteriors and lower-bound data term of two mean-field iterations
    against the oracle chain (mixtures/ilr.py:196-242)."""
    import mimo_b200
    from mimo_b200.distributions.bayesian import MEANFIELD
    b = _bench()
    E = eng()
    K, d, o, N = 128, 8, 1, 6000
    c = d + 1
    w = dict(b.WORKLOADS['cfg2'], N=N)
    rng = np.random.default_rng(2)
    centres = 3.0 * rng.standard_normal((K, d))
    z = rng.integers(0, K, size=N)
    x = centres[z] + rng.standard_normal((N, d))
    y = np.sin(x @ rng.standard_normal((d, o))) + 0.3 * rng.standard_normal((N, o))
    x = (x - x.mean(0)) / x.std(0)
    y = (y - y.mean(0)) / y.std(0)
    tol = {'fp32': 1e-4, 'fp64': 1e-9}[precision]
    mimo_b200.set_default_precision(precision)
    try:
        model = b.build_model(w, precision=precision)
        s = model._session(x, y)
        zr = s.Z.double().cpu().numpy()
        xr, yr = zr[:, :d], zr[:, d:]
        resp = rng.dirichlet(np.ones(K), size=N).T
        bprior = (np.zeros((K, d)), 1e-2 * np.ones(K), np.stack(K * [1e2 * np.eye(d)]), (d + 1.0) * np.ones(K) + 1e-16)
        mprior = (np.zeros((K, o, c)), np.stack(K * [1e-2 * np.eye(c)]), np.stack(K * [1e1 * np.eye(o)]), (o + 1.0) * np.ones(K) + 1e-16)
        g0 = (np.ones(K), 5.0 * np.ones(K))
        s.stats_from_resp(resp)
        for it in range(2):
            bpost = orc.nw_nat_to_std(orc.add_stats(orc.nw_std_to_nat(*bprior), orc.gauss_full_wstats(xr, resp)))
            mpost = orc.mnw_nat_to_std(orc.add_stats(orc.mnw_std_to_nat(*mprior), orc.lingauss_wstats(xr, yr, resp)), tied=True)
            gp = orc.stick_posterior(g0[0], g0[1], orc.categorical_wstats(resp))
            ell = orc.nw_expected_loglik(xr, *bpost) + orc.mnw_expected_loglik(xr, yr, *mpost) + orc.stick_expected_log(*gp)[0][:, None]
            resp, lse = orc.responsibilities(ell)
            ops, outs = s.update_from_stats(MEANFIELD)
            s.check(outs)
            for key, r in zip(('m', 'kappa', 'psi', 'nu'), bpost):
                close(outs['parts'][0][key], r, 10 * tol, 'iteration %d basis posterior %s' % (it, key))
            for key, r in zip(('M', 'K', 'psi', 'nu'), mpost):
                close(outs['parts'][1][key], r, 10 * tol, 'iteration %d expert posterior %s' % (it, key))
            s.sweep(ops, hard=False)
            assert abs(s.lse_sum.item() - lse.sum()) <= 10 * tol * abs(lse.sum()), 'iteration %d sum of log-normalisers' % it
            close(s.counts_host(), orc.categorical_wstats(resp), 10 * tol, 'iteration %d soft counts' % it)
    finally:
        mimo_b200.set_default_precision('fp32')


# ------------------------------------------------------------------------------------------------------------
# public per-phase methods, directly (fixtures from the unmodified reference)
# ------------------------------------------------------------------------------------------------------------
def test_resample_components_and_gating_direct():
    """BayesianMixtureOfGaussians.resample_components / resample_gating (mixtures/gmm.py:232-237) called on their own,
    from the reference's seeded state: posterior and the SAMPLED likelihood parameters replay the reference."""
    import mimo_b200
    from test_gpu_api import make_gmm
    mimo_b200.set_default_precision('fp64')
    try:
        g = load('gmm_toy_gibbs')
        model = make_gmm(g)
        K = int(g['K'])
        npr.seed(int(g['seed']))
        labels = npr.choice(K, size=len(g['obs']))
        assert np.array_equal(labels, g['labels_init'])
        model.resample_components(g['obs'], labels)
        for key, p in zip(('mus', 'kappas', 'psis', 'nus'), model.components.posterior.params):
            close(p, g['post_%s_0' % key], 1e-9, 'resample_components posterior ' + key)
        close(model.components.likelihood.mus, g['lik_mus_0'], 1e-8, 'sampled mus')
        close(model.components.likelihood.lmbdas, g['lik_lmbdas_0'], 1e-8, 'sampled lmbdas')
        model.resample_gating(labels)
        close(model.gating.posterior.alphas, g['gate_alphas_0'], 1e-12, 'gating posterior')
        close(model.gating.likelihood.probs, g['probs_0'], 1e-9, 'sampled gating probabilities')
        log_prob, lab = model.resample_labels(g['obs'])
        close(log_prob, g['log_prob_0'], 1e-9, 'log_prob of resample_labels')
        assert np.array_equal(lab, g['labels_0'])
    finally:
        mimo_b200.set_default_precision('fp32')


@pytest.mark.parametrize('name', ['gmm_toy_vi', 'gmm_toy_vi_stick'])
def test_meanfield_update_parameters_direct(name):
    """meanfield_update_parameters (mixtures/gmm.py:289-297) on explicit responsibilities; afterwards the likelihood
    parameters are a posterior DRAW (valid probabilities), as in the reference (bayesian.py:83, 230)."""
    import mimo_b200
    from test_gpu_api import make_gmm
    mimo_b200.set_default_precision('fp64')
    try:
        g = load(name)
        model = make_gmm(g)
        npr.seed(3)
        model.meanfield_update_parameters(g['obs'], g['resp_init'])
        for key, p in zip(('mus', 'kappas', 'psis', 'nus'), model.components.posterior.params):
            close(p, g['post_%s_0' % key], 1e-9, 'meanfield_update_parameters posterior ' + key)
        if 'gate_alphas_0' in g:
            close(model.gating.posterior.alphas, g['gate_alphas_0'], 1e-12, 'gating posterior')
        else:
            close(model.gating.posterior.gammas, g['gate_gammas_0'], 1e-12, 'stick gammas')
            close(model.gating.posterior.deltas, g['gate_deltas_0'], 1e-12, 'stick deltas')
        probs = model.gating.likelihood.probs
        assert abs(probs.sum() - 1.0) < 1e-12 and np.all(probs >= 0), 'gating likelihood must hold a probability vector'
        assert model.gating.likelihood.rvs(5).shape == (5,)
        close(model.expected_responsibilities(g['obs']), g['resp_0'], 1e-8, 'responsibilities after the update')
    finally:
        mimo_b200.set_default_precision('fp32')


def test_soft_sweep_with_permuted_feature_table():
    """mimo_sweep accepts ANY (fi, fj) table: a permuted table of the same length must give the permuted statistics
    (the list / tensor-core kernels only produce the canonical order; the sweep has to notice).  D = 32: tensor-core
    path; D = 16: responsibility-list path."""
    E = eng()
    for d, K, N in ((32, 40, 6000), (16, 12, 5000)):
        rng = np.random.default_rng(d)
        centres = 3.0 * rng.standard_normal((K, d))
        x = centres[rng.integers(0, K, size=N)] + rng.standard_normal((N, d))
        ops = E.QuadOperands(K, d, d, 'fp32')
        E.set_log_weights(ops, np.log(rng.dirichlet(np.ones(K))))
        E.operands_gauss(ops, E.to_dev(centres), E.to_dev(np.stack(K * [np.eye(d)]))).check()
        Z = E.to_dev(x, torch.float32)
        canon = E.quad_features(d)
        ref = E.SweepBuffers(N, K, canon.F, 'fp32', False)
        E.sweep(Z, ops, canon, ref)
        perm = rng.permutation(canon.F)
        other = E.Features(canon.fi_host[perm], canon.fj_host[perm], d)
        buf = E.SweepBuffers(N, K, other.F, 'fp32', False)
        E.sweep(Z, ops, other, buf)
        close(buf.stat, ref.stat.cpu().numpy()[:, perm], 2e-5, 'permuted feature table, d=%d' % d)
        assert abs(buf.lse_sum.item() - ref.lse_sum.item()) <= 1e-6 * abs(ref.lse_sum.item())
