"""Hierarchical mixtures of Gaussians (mimo/mixtures/hgmm.py, distributions/bayesian.py:595-793) on the
GPU against fixtures made from the unmodified reference (oracle/make_golden.py hier): seeded constructors, mean-field
trajectories, a Gibbs chain, natural-gradient steps, a mixture of mixtures.  Tolerances: 1e-8 in FP64 mode, 2e-4 FP32."""

import numpy as np
import numpy.random as npr
import pytest

from oracle import mimo_oracle as orc
from test_hier_cpu import load, make_hgmm

pytestmark = pytest.mark.gpu
TOL = {'fp32': 2e-4, 'fp64': 1e-8}


def close(a, b, tol, what=''):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(1.0, float(np.max(np.abs(b))))
    assert np.allclose(a, b, rtol=tol, atol=tol * scale), \
        '%s: max abs err %.3e (scale %.3e, tol %.1e)' % (what, float(np.max(np.abs(a - b))), scale, tol)


@pytest.fixture(params=['fp64', 'fp32'])
def precision(request):
    import mimo_b200
    mimo_b200.set_default_precision(request.param)
    yield request.param
    mimo_b200.set_default_precision('fp32')


@pytest.fixture
def fp64():
    import mimo_b200
    mimo_b200.set_default_precision('fp64')
    yield 'fp64'
    mimo_b200.set_default_precision('fp32')


def hyper_close(comp, g, tol):
    for a, n in zip(comp.hyper_posterior.params, ('rho', 'kappa', 'psi', 'nu')):
        close(a, g[f'hyper_{n}_end'], tol, 'hyper-posterior ' + n)


@pytest.mark.parametrize('name', ['hgmm_vi', 'hgmm_d8_vi_stick', 'hgmm_gibbs'])
def test_constructor_replays_reference_stream(name, fp64):
    """the constructors consume numpy.random like the reference's (bayesian.py:597-617): a seeded script starts from the
    reference's state."""
    g = load(name)
    model = make_hgmm(g)
    comp = model.components
    close(comp.prior.mus, g['init_prior_mus'], 1e-9, 'prior taus')
    close(comp.prior.lmbdas, g['init_prior_lmbdas'], 1e-9, 'prior lmbdas')
    close(comp.posterior.lmbdas, g['init_prior_lmbdas'], 1e-9)
    close(comp.likelihood.mus, g['init_lik_mus'], 1e-9, 'likelihood mus')
    close(model.gating.likelihood.probs, g['init_probs'], 1e-12, 'gating probs')
    assert comp.hyper_posterior is not comp.hyper_prior and comp.posterior is not comp.prior


@pytest.mark.parametrize('name', ['hgmm_vi', 'hgmm_d8_vi_stick'])
def test_hgmm_meanfield_trajectory(name, precision):
    g = load(name)
    tol = TOL[precision]
    model = make_hgmm(g)
    comp = model.components
    npr.seed(int(g['seed']))
    vlb = model.meanfield_coordinate_descent(g['obs'], randomize=True, maxiter=int(g['iters']), maxsubiter=int(g['subiters']),
                                             tol=0., progress_bar=False)
    close(vlb, g['vlb'], tol, 'lower bound')
    close(comp.posterior.mus, g['post_mus_end'], 10 * tol, 'posterior means')
    close(comp.posterior.kappas, g['post_kappas_end'], 10 * tol, 'posterior kappas')
    hyper_close(comp, g, 10 * tol)
    close(comp.likelihood.mus, g['lik_mus_end'], 10 * tol, 'mode means')
    close(comp.likelihood.lmbdas, g['lik_lmbdas_end'], 10 * tol, 'mode precision')
    if int(g['stick']):
        close(model.gating.posterior.gammas, g['gate_gammas_end'], 10 * tol, 'gammas')
        close(model.gating.posterior.deltas, g['gate_deltas_end'], 10 * tol, 'deltas')
    else:
        close(model.gating.posterior.alphas, g['gate_alphas_end'], 10 * tol, 'alphas')
    close(model.expected_log_complete_likelihood(g['obs']), g['ell_end'], 10 * tol, 'E log joint')
    close(model.expected_responsibilities(g['obs']), g['resp_end'], 50 * tol, 'responsibilities')
    close(comp.variational_lowerbound(), g['comp_vlb_end'], 10 * tol, 'component terms of the bound')
    # the public bound for explicit responsibilities agrees with the fused one
    close(model.variational_lowerbound(g['obs'], g['resp_end']), g['vlb'][-1], 10 * tol, 'public vlb')
    # the wrapper on its own (bayesian.py:662-689, 731-749) against the oracle's restatement
    hyper = (g['hyper_mu0'], float(g['hyper_kappa0']), g['hyper_psi0'], float(g['hyper_nu0']))
    w = orc.responsibilities(np.random.default_rng(0).standard_normal((int(g['K']), len(g['obs']))))[0]
    hq0 = tuple(comp.hyper_posterior.params)
    comp.meanfield_update(g['obs'], w, nb_iter=3)
    xk, nk, xxk, _ = orc.gauss_full_wstats(g['obs'], w)
    mus, kap, hq = orc.hnw_meanfield_update(hyper, hq0, g['kappas0'], xk, nk, xxk, 3)
    close(comp.posterior.mus, mus, 10 * tol, 'meanfield_update: means')
    for a, b in zip(comp.hyper_posterior.params, hq):
        close(a, b, 10 * tol, 'meanfield_update: hyper-posterior')
    close(comp.expected_log_likelihood(g['obs']),
          orc.hnw_expected_loglik(g['obs'], hq, mus, kap[:, None, None] * comp.posterior.lmbdas), 10 * tol, 'expected_log_likelihood')


def test_hgmm_weighted_meanfield_matches_oracle(fp64):
    """weights (N,) multiply the responsibilities in the parameter update (hgmm.py:202): the non-fused path."""
    g = load('hgmm_vi')
    model = make_hgmm(g)
    comp = model.components
    lm0 = comp.posterior.lmbdas.copy()
    wts = np.random.default_rng(5).random(len(g['obs']))
    npr.seed(11)
    vlb = model.meanfield_coordinate_descent(g['obs'], randomize=True, weights=wts, maxiter=3, maxsubiter=2, tol=0., progress_bar=False)
    # oracle: the same loop with resp * weights in the update and the unweighted E-step in the bound
    npr.seed(11)
    K, N = int(g['K']), len(g['obs'])
    resp = npr.rand(K, N)
    resp /= resp.sum(0)
    hyper = (g['hyper_mu0'], float(g['hyper_kappa0']), g['hyper_psi0'], float(g['hyper_nu0']))
    hq = hyper
    ent = None
    ref = []
    for _ in range(3):
        rw = resp * wts[None, :]
        xk, nk, xxk, _ = orc.gauss_full_wstats(g['obs'], rw)
        mus, kap, hq = orc.hnw_meanfield_update(hyper, hq, g['kappas0'], xk, nk, xxk, 2)
        om = kap[:, None, None] * lm0
        ent = om if ent is None else ent
        alphas = orc.dirichlet_posterior(g['gate_alphas0'], rw.sum(1))
        ell = orc.hnw_expected_loglik(g['obs'], hq, mus, om)
        resp, lse = orc.responsibilities(ell + orc.dirichlet_expected_log(alphas)[:, None])
        ref.append(orc.dirichlet_vlb(g['gate_alphas0'], alphas) + orc.hnw_vlb(hyper, hq, g['kappas0'], mus, om, ent) + lse.sum())
    close(vlb, ref, 1e-8, 'weighted lower bound')
    close(comp.posterior.mus, mus, 1e-8)


def test_hgmm_gibbs_chain_replays_reference(fp64):
    g = load('hgmm_gibbs')
    model = make_hgmm(g)
    comp = model.components
    npr.seed(int(g['seed']))
    model.resample(g['obs'], maxiter=int(g['sweeps']), maxsubiter=int(g['subiters']), progress_bar=False)
    close(comp.likelihood.mus, g['lik_mus_end'], 1e-7, 'sampled means')
    close(comp.likelihood.lmbdas, g['lik_lmbdas_end'], 1e-7, 'sampled precisions')
    close(model.gating.likelihood.probs, g['probs_end'], 1e-9, 'sampled probabilities')
    close(comp.posterior.mus, g['post_mus_end'], 1e-7, 'posterior means')
    hyper_close(comp, g, 1e-7)
    npr.seed(int(g['seed']) + 1)
    log_prob, labels = model.resample_labels(g['obs'])
    close(log_prob, g['log_prob_end'], 1e-7, 'log_prob')
    assert np.array_equal(labels, g['labels_next'])


def test_hgmm_natural_gradient_steps(precision):
    g = load('hgmm_svi')
    tol = TOL[precision]
    model = make_hgmm(g)
    npr.seed(int(g['seed']))
    out = model.meanfield_stochastic_descent(g['obs'], randomize=True, maxiter=int(g['iters']), maxsubiter=int(g['subiters']),
                                             step_size=float(g['step_size']), progress_bar=False)
    assert out == []
    comp = model.components
    close(comp.posterior.mus, g['post_mus_end'], 20 * tol, 'posterior means')
    close(comp.posterior.kappas, g['post_kappas_end'], 20 * tol, 'posterior kappas')
    hyper_close(comp, g, 20 * tol)
    close(model.gating.posterior.alphas, g['gate_alphas_end'], 20 * tol, 'alphas')
    close(model.expected_responsibilities(g['obs']), g['resp_end'], 100 * tol, 'responsibilities')


def test_mixture_of_mixtures_meanfield(fp64):
    """hgmm.py:298-431: nested mean field, clusters trained with per-point weights."""
    from mimo_b200.distributions import Dirichlet, CategoricalWithDirichlet
    from mimo_b200.mixtures import BayesianMixtureOfMixtureOfGaussians
    g = load('hmom_vi')
    M_, K, d = int(g['M']), int(g['K']), int(g['d'])
    npr.seed(int(g['ctor_seed']))
    gating = CategoricalWithDirichlet(M_, Dirichlet(M_, np.ones(M_)))
    sub_g = dict(K=K, d=d, stick=0, gate_alphas0=np.ones(K), hyper_mu0=np.zeros(d), hyper_kappa0=1e-2, hyper_psi0=np.eye(d),
                 hyper_nu0=d + 1 + 1e-8, kappas0=1e-2 * (1. + np.arange(K)))
    subs = [make_hgmm(sub_g, ctor_seed=int(g['ctor_seed']) + 1 + m) for m in range(M_)]
    model = BayesianMixtureOfMixtureOfGaussians(M_, K, d, gating=gating, components=subs)
    npr.seed(int(g['seed']))
    model.meanfield_coordinate_descent(g['obs'], randomize=True, maxiter=int(g['iters']), maxsubiter=int(g['subiters']),
                                       maxsubsubiter=int(g['subsubiters']), progress_bar=False)
    close(gating.posterior.alphas, g['gate_alphas_end'], 1e-7, 'cluster alphas')
    for m, sub in enumerate(subs):
        close(sub.components.posterior.mus, g[f'sub{m}_post_mus'], 1e-7, 'sub-mixture means')
        close(sub.components.hyper_posterior.params[2], g[f'sub{m}_hyper_psi'], 1e-7, 'sub-mixture psi')
    close(model.expected_responsibilities(g['obs']), g['resp_end'], 1e-6, 'cluster responsibilities')
    # the bare mixture of mixtures built from the same parts
    lp = model.likelihood.log_complete_likelihood(g['obs'])
    assert lp.shape == (M_, len(g['obs'])) and np.all(np.isfinite(lp))
    close(model.likelihood.responsibilities(g['obs']).sum(0), np.ones(len(g['obs'])), 1e-9)


def test_single_gaussian_hierarchical_wrapper(fp64):
    """bayesian.py:503-592 (examples/hgauss): mean-field and Gibbs updates of one Gaussian against the closed form."""
    from mimo_b200.distributions import NormalWishart, GaussianWithScaledPrecision, GaussianWithHierarchicalNormalWishart
    rng = np.random.default_rng(2)
    d = 3
    x = rng.standard_normal((300, d)) @ np.diag([1., 2., .5]) + np.array([1., -2., .5])
    npr.seed(4)
    hp = NormalWishart(dim=d, mu=np.zeros(d), kappa=1e-2, psi=np.eye(d), nu=d + 1 + 1e-8)
    w = GaussianWithHierarchicalNormalWishart(d, hp, GaussianWithScaledPrecision(d, kappa=1e-2))
    w.meanfield_update(x, nb_iter=10)
    hyper = tuple(hp.params)
    xk, nk, xxk = x.sum(0)[None], np.array([300.]), (x.T @ x)[None]
    mus, kap, hq = orc.hnw_meanfield_update(hyper, hyper, np.array([1e-2]), xk, nk, xxk, 10)
    close(w.posterior.mu, mus[0], 1e-9, 'posterior mean')
    for a, b in zip(w.hyper_posterior.params, hq):
        close(a, b, 1e-8, 'hyper-posterior')
    close(np.linalg.inv(w.likelihood.lmbda), np.cov(x.T), 0.2, 'precision near the sample covariance')
    w.resample(x, nb_iter=3)
    assert np.all(np.isfinite(w.likelihood.mu)) and np.all(np.linalg.eigvalsh(w.likelihood.lmbda) > 0)


# ---- hierarchical mixtures of linear experts (mixtures/hilr.py:79-291) ---------------------------------------------------
from test_hier_cpu import make_hilr  # noqa: E402


def hilr_close(model, g, tol):
    b, m = model.basis, model.models
    close(m.slope_posterior.M, g['slope_M'], tol, 'slope M')
    close(m.slope_posterior.K, g['slope_K'], tol, 'slope K')
    close(m.precision_posterior.psi, g['prec_psi'], tol, 'precision psi')
    close(m.precision_posterior.nu, g['prec_nu'], tol, 'precision nu')
    close(m.offset_posterior.mus, g['off_mus'], tol, 'offsets')
    close(m.offset_posterior.kappas, g['off_kappas'], tol, 'offset kappas')
    close(b.posterior.mus, g['basis_post_mus'], tol, 'basis means')
    for a, n in zip(b.hyper_posterior.params, ('rho', 'kappa', 'psi', 'nu')):
        close(a, g[f'basis_hyper_{n}'], tol, 'basis hyper-posterior ' + n)
    close(m.likelihood.As, g['lik_As'], tol, 'likelihood slopes')
    close(m.likelihood.cs, g['lik_cs'], tol, 'likelihood offsets')
    close(m.likelihood.lmbdas, g['lik_lmbdas'], tol, 'likelihood precisions')
    close(b.likelihood.mus, g['basis_lik_mus'], tol, 'basis likelihood means')


@pytest.mark.parametrize('name', ['hilr_vi', 'hilr_d2_vi_stick'])
def test_hilr_constructor_and_meanfield(name, precision):
    g = load(name)
    tol = TOL[precision]
    model = make_hilr(g)
    close(model.basis.prior.lmbdas, g['init_basis_lmbdas'], 1e-9, 'constructor: basis precisions')
    close(model.basis.likelihood.mus, g['init_basis_lik_mus'], 1e-9, 'constructor: basis means')
    close(model.models.likelihood.As, g['init_As'], 1e-9, 'constructor: slopes')
    close(model.models.likelihood.cs, g['init_cs'], 1e-9, 'constructor: offsets')
    close(model.models.likelihood.lmbdas, g['init_lmbdas'], 1e-9, 'constructor: precisions')
    close(model.gating.likelihood.probs, g['init_probs'], 1e-12, 'constructor: probabilities')
    npr.seed(int(g['seed']))
    out = model.meanfield_coordinate_descent(g['x'], g['y'], randomize=True, maxiter=int(g['iters']), maxsubiter=int(g['subiters']),
                                             progress_bar=False)
    assert out == []
    hilr_close(model, g, 20 * tol)
    close(model.expected_log_complete_likelihood(g['x'], g['y']), g['ell_end'], 20 * tol, 'E log joint')
    resp = model.expected_responsibilities(g['x'], g['y'])
    close(resp, g['resp_end'], 100 * tol, 'responsibilities')
    close(model.models.variational_lowerbound(), g['models_vlb'], 20 * tol, 'experts: lower-bound terms')
    close(model.variational_lowerbound(g['x'], g['y'], g['resp_end']), g['vlb_end'], 20 * tol, 'public lower bound')
    # the wrapper on its own against the oracle's restatement (bayesian.py:1298-1383 from raw data)
    m = model.models
    w = orc.responsibilities(np.random.default_rng(1).standard_normal((int(g['K']), len(g['x']))))[0]
    off0 = m.offset_posterior.mus.copy()
    m.meanfield_update(g['x'], g['y'], w, nb_iter=3)
    slope, prec, offs = orc.tam_meanfield_update((m.slope_prior.M, m.slope_prior.K), (m.precision_prior.psi, m.precision_prior.nu),
                                                 (m.offset_prior.mus, m.offset_prior.kappas), off0, g['x'], g['y'], w, 3)
    close(m.slope_posterior.M, slope[0], 20 * tol, 'meanfield_update: slope')
    close(m.precision_posterior.psi, prec[0], 20 * tol, 'meanfield_update: psi')
    close(m.offset_posterior.mus, offs[0], 20 * tol, 'meanfield_update: offsets')
    close(m.expected_log_likelihood(g['x'], g['y']), orc.tam_expected_loglik(g['x'], g['y'], slope, offs, prec), 20 * tol,
          'expected_log_likelihood')


def test_hilr_fused_lower_bound_is_the_public_one(fp64):
    """lower_bound=True (an extension: the reference's loop has its bound commented out): parameter terms + sum_n
    logsumexp from the fused sweep equal the public bound evaluated with the E-step responsibilities."""
    g = load('hilr_vi')
    model = make_hilr(g)
    npr.seed(int(g['seed']))
    vlb = model.meanfield_coordinate_descent(g['x'], g['y'], randomize=True, maxiter=3, maxsubiter=3, progress_bar=False, lower_bound=True)
    resp = model.expected_responsibilities(g['x'], g['y'])
    close(vlb[-1], model.variational_lowerbound(g['x'], g['y'], resp), 1e-9, 'fused bound')
    assert len(vlb) == 3


def test_hilr_gibbs_chain_replays_reference(fp64):
    g = load('hilr_gibbs')
    model = make_hilr(g)
    npr.seed(int(g['seed']))
    model.resample(g['x'], g['y'], maxiter=int(g['sweeps']), maxsubiter=int(g['subiters']), progress_bar=False)
    hilr_close(model, g, 1e-7)
    close(model.gating.likelihood.probs, g['probs'], 1e-9, 'sampled probabilities')
    npr.seed(int(g['seed']) + 1)
    log_prob, labels = model.resample_labels(g['x'], g['y'])
    close(log_prob, g['log_prob_end'], 1e-7, 'log_prob')
    assert np.array_equal(labels, g['labels_next'])
    with pytest.raises(NotImplementedError):
        model.meanfield_stochastic_descent(g['x'], g['y'], maxiter=1, maxsubiter=1, progress_bar=False)


def test_mixture_of_mixtures_of_linear_experts(fp64):
    """hilr.py:293-609: nested mean field of tied-activation mixtures trained with per-point weights, then the predictive
    path (weights over clusters x experts, posterior-predictive moments, mixture / mode)."""
    from mimo_b200.distributions import Dirichlet, CategoricalWithDirichlet
    from mimo_b200.mixtures import BayesianMixtureOfMixtureOfLinearGaussians
    g = load('hmoilr_vi')
    M_, K = int(g['M']), int(g['K'])
    npr.seed(int(g['ctor_seed']))
    gating = CategoricalWithDirichlet(M_, Dirichlet(M_, np.ones(M_)))
    subs = []
    for m in range(M_):
        sg = dict(g)
        sg['ctor_seed'] = int(g['ctor_seed']) + 1 + m
        subs.append(make_hilr(sg))
    model = BayesianMixtureOfMixtureOfLinearGaussians(M_, K, int(g['din']), int(g['o']), gating=gating, components=subs)
    npr.seed(int(g['seed']))
    out = model.meanfield_coordinate_descent(g['x'], g['y'], randomize=True, maxiter=int(g['iters']), maxsubiter=int(g['subiters']),
                                             maxsubsubiter=int(g['subsubiters']), progress_bar=False)
    assert out == []
    close(gating.posterior.alphas, g['gate_alphas_end'], 1e-7, 'cluster alphas')
    for m, sub in enumerate(subs):
        close(sub.models.slope_posterior.M, g[f'sub{m}_slope_M'], 1e-7, 'slopes')
        close(sub.models.offset_posterior.mus, g[f'sub{m}_off_mus'], 1e-7, 'offsets')
        close(sub.basis.posterior.mus, g[f'sub{m}_basis_mus'], 1e-7, 'input-density means')
    close(model.expected_responsibilities(g['x'], g['y']), g['resp_end'], 1e-6, 'cluster responsibilities')
    close(model.meanfield_predictive_weights(g['x']), g['weights'], 1e-6, 'predictive weights')
    for pred in ('average', 'mode'):
        mean, var, std = model.meanfield_prediction(g['x'], prediction=pred)
        close(mean, g[f'pred_{pred}_mean'], 1e-6, pred + ' prediction')
        close(var, g[f'pred_{pred}_var'], 1e-6, pred + ' predictive variance')
    lp = model.likelihood.log_complete_likelihood(g['x'], g['y'])
    assert lp.shape == (M_, len(g['x'])) and np.all(np.isfinite(lp))


def test_hilr_weighted_meanfield_matches_oracle(fp64):
    """weights (N,) multiply the responsibilities in every parameter update (hilr.py:191): the non-fused path a mixture of
    mixtures drives, against the oracle's loop from raw data."""
    g = load('hilr_vi')
    model = make_hilr(g)
    b, m = model.basis, model.models
    x, y, K = g['x'], g['y'], int(g['K'])
    lm0 = b.posterior.lmbdas.copy()
    hyper = tuple(b.hyper_prior.params)
    sp, pp, op = (m.slope_prior.M, m.slope_prior.K), (m.precision_prior.psi, m.precision_prior.nu), (m.offset_prior.mus, m.offset_prior.kappas)
    off = m.offset_posterior.mus.copy()
    wts = np.random.default_rng(6).random(len(x))
    npr.seed(12)
    model.meanfield_coordinate_descent(x, y, randomize=True, weights=wts, maxiter=3, maxsubiter=2, progress_bar=False)
    npr.seed(12)
    resp = npr.rand(K, len(x))
    resp /= resp.sum(0)
    hq = hyper
    for _ in range(3):
        rw = resp * wts[None, :]
        xk, nk, xxk, _ = orc.gauss_full_wstats(x, rw)
        bm, bk, hq = orc.hnw_meanfield_update(hyper, hq, b.prior.kappas, xk, nk, xxk, 2)
        slope, prec, offs = orc.tam_meanfield_update(sp, pp, op, off, x, y, rw, 2)
        off = offs[0]
        alphas = orc.dirichlet_posterior(g['gate_alphas0'], rw.sum(1))
        joint = orc.hnw_expected_loglik(x, hq, bm, bk[:, None, None] * lm0) + orc.tam_expected_loglik(x, y, slope, offs, prec) \
            + orc.dirichlet_expected_log(alphas)[:, None]
        resp = orc.responsibilities(joint)[0]
    close(b.posterior.mus, bm, 1e-8, 'basis means')
    close(m.slope_posterior.M, slope[0], 1e-8, 'slope')
    close(m.offset_posterior.mus, offs[0], 1e-8, 'offsets')
    close(model.gating.posterior.alphas, alphas, 1e-8, 'alphas')
    close(model.expected_responsibilities(x, y), resp, 1e-7, 'responsibilities')
    close(model.expected_log_likelihood(x, y), orc.responsibilities(joint)[1], 1e-8, 'expected log-likelihood')


def test_single_affine_expert_gibbs_replays_reference(fp64):
    """bayesian.py:1137-1219 (examples/lingauss): slope, offset and precision of one expert, seeded chain."""
    from mimo_b200.distributions import (MatrixNormalWithPrecision, GaussianWithScaledPrecision, Wishart,
                                         AffineLinearGaussianWithMatrixNormalWishart)
    g = load('affine_expert_gibbs')
    din, o = g['x'].shape[1], g['y'].shape[1]
    npr.seed(int(g['ctor_seed']))
    w = AffineLinearGaussianWithMatrixNormalWishart(din, o, slope_prior=MatrixNormalWithPrecision(column_dim=din, row_dim=o, M=np.zeros((o, din)), K=1e-2 * np.eye(din)),
                                                    offset_prior=GaussianWithScaledPrecision(dim=o, kappa=1e-2, mu=np.zeros(o)),
                                                    precision_prior=Wishart(dim=o, psi=np.eye(o), nu=o + 1 + 1e-8))
    close(w.likelihood.A, g['init_A'], 1e-10, 'constructor: slope')
    close(w.likelihood.c, g['init_c'], 1e-10, 'constructor: offset')
    npr.seed(int(g['seed']))
    w.resample(g['x'], g['y'], nb_iter=int(g['iters']))
    for key, val in (('A', w.likelihood.A), ('c', w.likelihood.c), ('lmbda', w.likelihood.lmbda), ('slope_M', w.slope_posterior.M),
                     ('slope_K', w.slope_posterior.K), ('psi', w.precision_posterior.psi), ('nu', w.precision_posterior.nu),
                     ('off_mu', w.offset_posterior.mu), ('off_kappa', w.offset_posterior.kappa)):
        close(val, g[key], 1e-8, key)
    ll = w.likelihood.log_likelihood(g['x'], g['y'])
    ref = orc.lingauss_loglik(g['x'], g['y'], np.hstack((g['A'], g['c'][:, None]))[None], g['lmbda'][None], affine=True)[0]
    close(ll, ref, 1e-9, 'log-likelihood of the sampled expert')


def test_mixture_of_mixtures_em(fp64):
    """hgmm.py:59-89 (examples/hgmm/em_hgmm.py): EM of a mixture of mixtures of tied Gaussians; the clusters are trained
    by weighted EM (gmm.py:77-103 with weights)."""
    from mimo_b200.distributions import Categorical, TiedGaussiansWithPrecision
    from mimo_b200.mixtures import MixtureOfGaussians, MixtureOfMixtureOfGaussians
    g = load('hmom_em')
    M_, K, d = int(g['M']), int(g['K']), int(g['d'])
    comps = [MixtureOfGaussians(gating=Categorical(dim=K, probs=g[f'probs{m}']),
                                components=TiedGaussiansWithPrecision(size=K, dim=d, mus=g[f'mus{m}'], lmbdas=np.stack(K * [0.5 * np.eye(d)])))
             for m in range(M_)]
    model = MixtureOfMixtureOfGaussians(cluster_size=M_, mixture_size=K, dim=d, gating=Categorical(dim=M_), components=comps)
    npr.seed(int(g['seed']))
    ll = model.max_likelihood(g['obs'], maxiter=int(g['iters']), maxsubiter=int(g['subiters']), progress_bar=False)
    close(ll, g['ll'], 1e-8, 'log-likelihood trajectory')
    assert np.all(np.diff(ll) >= -1e-8)                        # "ll monoton?" of the example
    close(model.gating.probs, g['gate_probs'], 1e-8, 'cluster probabilities')
    for m, c in enumerate(comps):
        close(c.components.mus, g[f'end_mus{m}'], 1e-7, 'means')
        close(c.components.lmbdas, g[f'end_lmbdas{m}'], 1e-7, 'tied precisions')
        close(c.gating.probs, g[f'end_probs{m}'], 1e-8, 'local probabilities')
    close(model.responsibilities(g['obs']), g['resp_end'], 1e-7, 'cluster responsibilities')
