"""Hierarchical mixtures, host side: the scaled-precision prior classes (no GPU needed) and the model
builder shared with tests/test_gpu_hier.py."""
import os

import numpy as np
import numpy.random as npr

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def load(name):
    return dict(np.load(os.path.join(GOLD, name + '.npz')))


def make_hgmm(g, ctor_seed=None):
    from mimo_b200.distributions import (Dirichlet, TruncatedStickBreaking, CategoricalWithDirichlet, CategoricalWithStickBreaking,
                                         NormalWishart, TiedGaussiansWithScaledPrecision,
                                         TiedGaussiansWithHierarchicalNormalWisharts)
    from mimo_b200.mixtures import BayesianMixtureOfGaussiansWithHierarchicalPrior
    K, d = int(g['K']), int(g['d'])
    npr.seed(int(g['ctor_seed']) if ctor_seed is None else ctor_seed)
    if int(g.get('stick', 0)):
        gating = CategoricalWithStickBreaking(K, TruncatedStickBreaking(K, g['gate_gammas0'], g['gate_deltas0']))
    else:
        gating = CategoricalWithDirichlet(K, Dirichlet(K, g['gate_alphas0']))
    hp = NormalWishart(dim=d, mu=g['hyper_mu0'], kappa=float(g['hyper_kappa0']), psi=g['hyper_psi0'], nu=float(g['hyper_nu0']))
    pr = TiedGaussiansWithScaledPrecision(size=K, dim=d, kappas=g['kappas0'])
    comp = TiedGaussiansWithHierarchicalNormalWisharts(size=K, dim=d, hyper_prior=hp, prior=pr)
    return BayesianMixtureOfGaussiansWithHierarchicalPrior(size=K, dim=d, gating=gating, components=comp)


def test_scaled_precision_prior_algebra():
    from mimo_b200.distributions import TiedGaussiansWithScaledPrecision, GaussianWithScaledPrecision
    rng = np.random.default_rng(0)
    K, d = 3, 4
    A = rng.standard_normal((K, d, d))
    lm = A @ A.transpose(0, 2, 1) + np.eye(d)
    p = TiedGaussiansWithScaledPrecision(K, d, kappas=np.array([.5, 1., 2.]), mus=rng.standard_normal((K, d)), lmbdas=lm)
    np.testing.assert_allclose(p.omegas, p.kappas[:, None, None] * lm)
    np.testing.assert_allclose(p.sigmas, np.linalg.inv(p.omegas), rtol=1e-10)
    nat = p.nat_param
    q = TiedGaussiansWithScaledPrecision(K, d, kappas=np.ones(K), mus=np.zeros((K, d)), lmbdas=lm)
    q.nat_param = nat
    np.testing.assert_allclose(q.mus, p.mus)
    np.testing.assert_allclose(q.kappas, p.kappas)
    # quirk q11: the cached Cholesky factor survives a new kappa, a new lmbda resets it (gaussian.py:947-961)
    one = GaussianWithScaledPrecision(d, 2., mu=np.zeros(d), lmbda=lm[0])
    e0 = one.entropy()
    one.kappa = 8.
    assert one.entropy() == e0 and not np.allclose(one.omega, 2. * lm[0])
    one.lmbda = lm[0]
    assert abs(one.entropy() - (e0 - 0.5 * d * np.log(4.))) < 1e-12
    x = rng.standard_normal((5, d))
    w = rng.random((K, 5))
    xk, nk = p.weighted_statistics(x, w)
    np.testing.assert_allclose(xk, w @ x)
    np.testing.assert_allclose(nk, w.sum(1))
    # the reference's log-partition uses mu^T lmbda mu, not mu^T omega mu (gaussian.py:1011-1013): a density only for
    # kappa = 1 (component 1 here); kept as the reference has it
    from scipy.stats import multivariate_normal
    ll = p.log_likelihood(x)
    np.testing.assert_allclose(ll[1], multivariate_normal(p.mus[1], np.linalg.inv(p.omegas[1])).logpdf(x), rtol=1e-9)
    for k in range(K):
        ref = x @ (p.omegas[k] @ p.mus[k]) - 0.5 * np.einsum('nd,dl,nl->n', x, p.omegas[k], x) - 0.5 * d * np.log(2 * np.pi) \
            - (0.5 * p.mus[k] @ lm[k] @ p.mus[k] - 0.5 * np.linalg.slogdet(p.omegas[k])[1])
        np.testing.assert_allclose(ll[k], ref, rtol=1e-9)


def make_hilr(g):
    from mimo_b200.distributions import (Dirichlet, TruncatedStickBreaking, CategoricalWithDirichlet, CategoricalWithStickBreaking,
                                         NormalWishart, Wishart, MatrixNormalWithPrecision, TiedGaussiansWithScaledPrecision,
                                         TiedGaussiansWithHierarchicalNormalWisharts,
                                         TiedAffineLinearGaussiansWithMatrixNormalWisharts)
    from mimo_b200.mixtures import BayesianMixtureOfLinearGaussiansWithTiedActivation
    K, din, o = int(g['K']), int(g['din']), int(g['o'])
    npr.seed(int(g['ctor_seed']))
    if int(g['stick']):
        gating = CategoricalWithStickBreaking(K, TruncatedStickBreaking(K, g['gate_gammas0'], g['gate_deltas0']))
    else:
        gating = CategoricalWithDirichlet(K, Dirichlet(K, g['gate_alphas0']))
    bh = NormalWishart(dim=din, mu=np.zeros(din), kappa=1e-2, psi=np.eye(din), nu=din + 1 + 1e-8)
    bp = TiedGaussiansWithScaledPrecision(size=K, dim=din, kappas=1e-2 * np.ones(K))
    basis = TiedGaussiansWithHierarchicalNormalWisharts(size=K, dim=din, hyper_prior=bh, prior=bp)
    sp = MatrixNormalWithPrecision(column_dim=din, row_dim=o, M=np.zeros((o, din)), K=1e-2 * np.eye(din))
    op = TiedGaussiansWithScaledPrecision(size=K, dim=o, mus=np.zeros((K, o)), kappas=g['off_kappas0'])
    pp = Wishart(dim=o, psi=np.eye(o), nu=o + 1 + 1e-8)
    models = TiedAffineLinearGaussiansWithMatrixNormalWisharts(size=K, column_dim=din, row_dim=o, slope_prior=sp,
                                                               offset_prior=op, precision_prior=pp)
    return BayesianMixtureOfLinearGaussiansWithTiedActivation(size=K, input_dim=din, output_dim=o, gating=gating,
                                                              basis=basis, models=models)


# ---- the host-side algebra of the hierarchical wrappers against the oracle's restatements (no GPU: the statistics are
#      formed with NumPy here; on the device path they come from the statistics kernels) -------------------------------
def _bare(cls, **attrs):
    obj = object.__new__(cls)                # the constructors sample from their priors on the device: bypassed
    for k, v in attrs.items():
        setattr(obj, k, v)
    return obj


def test_hyper_posterior_update_matches_oracle():
    import copy
    from oracle import mimo_oracle as orc
    from mimo_b200.distributions import NormalWishart, TiedGaussiansWithScaledPrecision, TiedGaussiansWithHierarchicalNormalWisharts
    from mimo_b200.distributions.bayesian import _hyper_nw_update
    rng = np.random.default_rng(1)
    K, d, N = 5, 3, 200
    x = rng.standard_normal((N, d)) @ np.diag([1., 2., .5]) + 3. * rng.standard_normal((K, d))[rng.integers(0, K, N)]
    w = orc.responsibilities(rng.standard_normal((K, N)))[0]
    xk, nk, xxk, _ = orc.gauss_full_wstats(x, w)
    hyper = (0.1 * rng.standard_normal(d), 0.3, np.eye(d) + 0.1, d + 2.5)
    k0 = 1e-2 * (1. + np.arange(K))
    mus = rng.standard_normal((K, d))
    for a, b in zip(_hyper_nw_update(hyper, k0, mus, xk, nk, xxk.sum(0)), orc.hnw_hyper_update(hyper, k0, mus, xk, nk, xxk)):
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12)
    # the wrapper's mean-field and natural-gradient alternations on given reduced statistics
    A = rng.standard_normal((K, d, d))
    lm = A @ A.transpose(0, 2, 1) + np.eye(d)
    prior = TiedGaussiansWithScaledPrecision(K, d, kappas=k0, mus=np.zeros((K, d)), lmbdas=lm)
    hp = NormalWishart(d, *hyper)
    comp = _bare(TiedGaussiansWithHierarchicalNormalWisharts, size=K, dim=d, hyper_prior=hp, hyper_posterior=copy.deepcopy(hp),
                 prior=prior, posterior=copy.deepcopy(prior))
    comp._meanfield(xk, nk, xxk.sum(0), 4)
    mus_o, kap_o, hq_o = orc.hnw_meanfield_update(hyper, hyper, k0, xk, nk, xxk, 4)
    np.testing.assert_allclose(comp.posterior.mus, mus_o, rtol=1e-10)
    np.testing.assert_allclose(comp.posterior.kappas, kap_o, rtol=1e-12)
    for a, b in zip(comp.hyper_posterior.params, hq_o):
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12)
    # lower-bound terms: first evaluation = fresh factors; after kappa moves the entropy keeps the cached ones (q11)
    om = comp.posterior.omegas
    v0 = comp.variational_lowerbound()
    np.testing.assert_allclose(v0, orc.hnw_vlb(hyper, hq_o, k0, mus_o, om), rtol=1e-10)
    comp.posterior.kappas = comp.posterior.kappas * 2.
    np.testing.assert_allclose(comp.variational_lowerbound(), orc.hnw_vlb(hyper, hq_o, k0, mus_o, comp.posterior.omegas, om), rtol=1e-10)
    assert abs(comp.variational_lowerbound() - orc.hnw_vlb(hyper, hq_o, k0, mus_o, comp.posterior.omegas)) > 1e-6


def test_tied_affine_update_matches_oracle():
    import copy
    from oracle import mimo_oracle as orc
    from mimo_b200.distributions import (Wishart, MatrixNormalWithPrecision, TiedGaussiansWithScaledPrecision,
                                         TiedAffineLinearGaussiansWithMatrixNormalWisharts)
    rng = np.random.default_rng(2)
    K, c, o, N = 4, 2, 2, 150
    x = rng.uniform(-1.5, 1.5, (N, c))
    y = x @ rng.standard_normal((c, o)) + 0.5 * rng.integers(-1, 2, (N, 1)) + 0.05 * rng.standard_normal((N, o))
    w = orc.responsibilities(rng.standard_normal((K, N)))[0]
    sp = MatrixNormalWithPrecision(column_dim=c, row_dim=o, M=0.1 * rng.standard_normal((o, c)), K=1e-2 * np.eye(c))
    op = TiedGaussiansWithScaledPrecision(K, o, kappas=1e-2 * (1. + np.arange(K)), mus=0.1 * rng.standard_normal((K, o)),
                                          lmbdas=np.stack(K * [np.eye(o)]))
    pp = Wishart(dim=o, psi=np.eye(o) + 0.2, nu=o + 1.5)
    m = _bare(TiedAffineLinearGaussiansWithMatrixNormalWisharts, size=K, column_dim=c, row_dim=o, slope_prior=sp, offset_prior=op,
              precision_prior=pp, slope_posterior=copy.deepcopy(sp), offset_posterior=copy.deepcopy(op), precision_posterior=copy.deepcopy(pp))
    mom = dict(xm=w @ x, ym=w @ y, n=w.sum(1), yx=np.einsum('nd,kn,nl->kdl', y, w, x), xx=np.einsum('nd,kn,nl->kdl', x, w, x),
               yy=np.einsum('nd,kn,nl->kdl', y, w, y))
    off0 = m.offset_posterior.mus.copy()
    m._meanfield(mom, 3)
    slope, prec, offs = orc.tam_meanfield_update((sp.M, sp.K), (pp.psi, pp.nu), (op.mus, op.kappas), off0, x, y, w, 3)
    np.testing.assert_allclose(m.slope_posterior.M, slope[0], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(m.slope_posterior.K, slope[1], rtol=1e-10)
    np.testing.assert_allclose(m.precision_posterior.psi, prec[0], rtol=1e-9)
    np.testing.assert_allclose(m.precision_posterior.nu, prec[1], rtol=1e-12)
    np.testing.assert_allclose(m.offset_posterior.mus, offs[0], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(m.offset_posterior.kappas, offs[1], rtol=1e-12)
    np.testing.assert_allclose(m.offset_posterior.lmbdas, np.stack(K * [prec[1] * prec[0]]), rtol=1e-9)
