"""Hierarchical mixtures, host side: the scaled-precision prior classes (no GPU needed) and the model
builder shared with tests/test_gpu_hier.py."""
import os

import numpy as np
import numpy.random as npr
import pytest

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
