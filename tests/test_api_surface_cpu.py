"""API surface of the mirror against the reference (CPU, host-side pieces only; skipped where /root/reference is
absent, i.e. on the GPU box): every class of the sweep path exists under the same name, the driver entry points keep
the reference's positional signatures, and the small closed forms added for drop-in completeness (base measures,
Matrix-Normal-Wishart std <-> natural maps, Dirichlet statistics) give the reference's numbers."""
import inspect
import os
import sys

import numpy as np
import pytest

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'mimo')), reason='reference not present')


@pytest.fixture(scope='module')
def mods():
    sys.path.insert(0, REF)
    import importlib
    ref_d, ref_m = importlib.import_module('mimo.distributions'), importlib.import_module('mimo.mixtures')
    import mimo_b200.distributions as my_d
    import mimo_b200.mixtures as my_m
    return ref_d, ref_m, my_d, my_m


PATH_CLASSES = ['Dirichlet', 'TruncatedStickBreaking', 'Categorical', 'CategoricalWithDirichlet', 'CategoricalWithStickBreaking',
                'NormalWishart', 'StackedNormalWisharts', 'TiedNormalWisharts', 'NormalGamma', 'StackedNormalGammas',
                'TiedNormalGammas', 'MatrixNormalWishart', 'StackedMatrixNormalWisharts', 'TiedMatrixNormalWisharts',
                'Wishart', 'Gamma', 'GaussianWithPrecision', 'StackedGaussiansWithPrecision', 'TiedGaussiansWithPrecision',
                'GaussianWithDiagonalPrecision', 'StackedGaussiansWithDiagonalPrecision', 'TiedGaussiansWithDiagonalPrecision',
                'LinearGaussianWithPrecision', 'StackedLinearGaussiansWithPrecision', 'TiedLinearGaussiansWithPrecision',
                'GaussianWithNormalWishart', 'StackedGaussiansWithNormalWisharts', 'TiedGaussiansWithNormalWisharts',
                'StackedGaussiansWithNormalGammas', 'TiedGaussiansWithNormalGammas',
                'StackedLinearGaussiansWithMatrixNormalWisharts', 'TiedLinearGaussiansWithMatrixNormalWisharts',
                # hierarchical mixtures
                'GaussianWithScaledPrecision', 'TiedGaussiansWithScaledPrecision', 'GaussianWithHierarchicalNormalWishart',
                'TiedGaussiansWithHierarchicalNormalWisharts', 'AffineLinearGaussianWithPrecision',
                'StackedAffineLinearGaussiansWithPrecision', 'AffineLinearGaussianWithMatrixNormalWishart',
                'TiedAffineLinearGaussiansWithMatrixNormalWisharts', 'MatrixNormalWithPrecision']
MIXTURES = ['MixtureOfGaussians', 'BayesianMixtureOfGaussians', 'MixtureOfLinearGaussians', 'BayesianMixtureOfLinearGaussians',
            'MixtureOfMixtureOfGaussians', 'BayesianMixtureOfGaussiansWithHierarchicalPrior', 'BayesianMixtureOfMixtureOfGaussians',
            'MixtureOfMixtureOfLinearGaussians', 'BayesianMixtureOfLinearGaussiansWithTiedActivation',
            'BayesianMixtureOfMixtureOfLinearGaussians']
DRIVER_METHODS = ['resample', 'resample_labels', 'resample_components', 'resample_basis', 'resample_models',
                  'meanfield_update_components', 'meanfield_update_basis', 'meanfield_update_models', 'meanfield_sgd_parameters',
                  'meanfield_coordinate_descent', 'meanfield_update_parameters',
                  'expected_responsibilities', 'expected_log_complete_likelihood', 'max_aposteriori', 'max_likelihood',
                  'meanfield_stochastic_descent', 'variational_lowerbound', 'log_likelihood', 'responsibilities',
                  'meanfield_prediction']


def test_path_classes_exist(mods):
    ref_d, ref_m, my_d, my_m = mods
    for n in PATH_CLASSES:
        assert hasattr(ref_d, n), 'not a reference class: ' + n
        assert hasattr(my_d, n), 'missing from mimo_b200.distributions: ' + n
    for n in MIXTURES:
        assert hasattr(my_m, n), 'missing from mimo_b200.mixtures: ' + n


def test_driver_signatures_match(mods):
    """same parameter names in the same order (the mirror may append keyword-only extras such as comm=)."""
    _, ref_m, _, my_m = mods
    for n in MIXTURES:
        r, m = getattr(ref_m, n), getattr(my_m, n)
        for meth in DRIVER_METHODS:
            if not hasattr(r, meth):
                continue
            assert hasattr(m, meth), '%s.%s missing' % (n, meth)
            pr = list(inspect.signature(getattr(r, meth)).parameters)
            pm = list(inspect.signature(getattr(m, meth)).parameters)
            assert pm[:len(pr)] == pr, '%s.%s: %s vs %s' % (n, meth, pr, pm)


def test_statistics_methods_of_likelihoods(mods):
    ref_d, _, my_d, _ = mods
    for n in ('StackedGaussiansWithPrecision', 'StackedGaussiansWithDiagonalPrecision', 'StackedLinearGaussiansWithPrecision'):
        for meth in ('log_likelihood', 'statistics', 'weighted_statistics', 'max_likelihood'):
            pr = list(inspect.signature(getattr(getattr(ref_d, n), meth)).parameters)
            pm = list(inspect.signature(getattr(getattr(my_d, n), meth)).parameters)
            assert pm[:len(pr)] == pr, '%s.%s: %s vs %s' % (n, meth, pr, pm)


def test_base_measures_and_mnw_maps(mods):
    ref_d, _, my_d, _ = mods
    rng = np.random.default_rng(0)
    d, c, o = 3, 4, 2
    psi = np.eye(d) + 0.1 * np.ones((d, d))
    r, m = ref_d.NormalWishart(d, rng.standard_normal(d), 0.7, psi, d + 2.5), None
    m = my_d.NormalWishart(d, r.gaussian.mu.copy(), 0.7, psi, d + 2.5)
    assert np.isclose(r.base, m.base) and np.isclose(r.log_base(), m.log_base())
    rg = ref_d.NormalGamma(d, rng.standard_normal(d), np.ones(d), 2. * np.ones(d), 3. * np.ones(d))
    mg = my_d.NormalGamma(d, rg.gaussian.mu.copy(), np.ones(d), 2. * np.ones(d), 3. * np.ones(d))
    assert np.isclose(rg.base, mg.base) and np.isclose(rg.log_base(), mg.log_base())
    M, K = rng.standard_normal((o, c)), np.eye(c) * 0.5 + 0.05
    psi_o, nu = np.eye(o) + 0.2, o + 3.0
    rm = ref_d.MatrixNormalWishart(c, o, M, K, psi_o, nu)
    mm = my_d.MatrixNormalWishart(c, o, M, K, psi_o, nu)
    assert np.isclose(rm.base, mm.base) and np.isclose(rm.log_base(), mm.log_base())
    nat_r, nat_m = rm.std_to_nat((M, K, psi_o, nu)), mm.std_to_nat((M, K, psi_o, nu))
    for a, b in zip(nat_r, nat_m):
        assert np.allclose(a, b, rtol=1e-12, atol=1e-12)
    back = mm.nat_to_std(nat_m)
    for a, b in zip(rm.nat_to_std(nat_r), back):
        assert np.allclose(a, b, rtol=1e-10, atol=1e-12)
    for a, b in zip(back, (M, K, psi_o, nu)):
        assert np.allclose(a, b, rtol=1e-10, atol=1e-12)
    assert ref_d.Wishart(d, psi, d + 1.).base == my_d.Wishart(d, psi, d + 1.).base == 1.
    assert ref_d.Gamma(d, np.ones(d), np.ones(d)).base == my_d.Gamma(d, np.ones(d), np.ones(d)).base == 1.


def test_dirichlet_statistics(mods):
    ref_d, _, my_d, _ = mods
    rng = np.random.default_rng(1)
    x = rng.dirichlet(np.ones(4), size=30)
    w = rng.random(30)
    r, m = ref_d.Dirichlet(4, np.ones(4)), my_d.Dirichlet(4, np.ones(4))
    assert np.allclose(r.statistics(x), m.statistics(x))
    assert np.allclose(r.weighted_statistics(x, w), m.weighted_statistics(x, w))
    lst = m.statistics([x[:10], x[10:]])
    assert isinstance(lst, list) and np.allclose(lst[1], np.log(x[10:]))


def test_matrix_normal_with_precision(mods):
    """mimo/distributions/matrix.py:10-176 -- densities, entropies, natural parameters and a seeded draw."""
    import numpy.random as npr
    ref_d, _, my_d, _ = mods
    rng = np.random.default_rng(2)
    o, c = 2, 3
    M = rng.standard_normal((o, c))
    a = rng.standard_normal((o, o + 2))
    V = a @ a.T / o + 0.3 * np.eye(o)
    b = rng.standard_normal((c, c + 2))
    K = b @ b.T / c + 0.3 * np.eye(c)
    r, m = ref_d.MatrixNormalWithPrecision(c, o, M, V, K), my_d.MatrixNormalWithPrecision(c, o, M, V, K)
    x = rng.standard_normal((7, o, c))
    x[3, 0, 1] = np.nan
    assert np.allclose(r.log_likelihood(x.copy()), m.log_likelihood(x.copy()), rtol=1e-12, atol=1e-12)
    assert np.isclose(r.log_partition(), m.log_partition())
    M2 = M + 0.5
    r2, m2 = ref_d.MatrixNormalWithPrecision(c, o, M2, V * 1.3, K * 0.7), my_d.MatrixNormalWithPrecision(c, o, M2, V * 1.3, K * 0.7)
    assert np.isclose(r.relative_entropy(r2), m.relative_entropy(m2))
    # entropy / cross-entropy: the reference raises (shape mismatch, matrix.py:160-168); check the closed forms instead
    n = o * c
    sig = m.sigma
    assert np.isclose(m.entropy(), 0.5 * n * (1. + np.log(2. * np.pi)) + 0.5 * np.linalg.slogdet(sig)[1])
    diff = m._vec(M)[0] - m2._vec(M2)[0]
    L2 = m2.lmbda
    ce = 0.5 * n * np.log(2. * np.pi) - 0.5 * np.linalg.slogdet(L2)[1] + 0.5 * np.trace(L2 @ sig) + 0.5 * diff @ L2 @ diff
    assert np.isclose(m.cross_entropy(m2), ce)
    assert r.nb_params == m.nb_params and np.isclose(r.base, m.base)
    for u, v in zip(r.nat_param, m.nat_param):
        assert np.allclose(u, v)
    for u, v in zip(r.nat_to_std(r.nat_param), m.nat_to_std(m.nat_param)):
        assert np.allclose(u, v)
    for u, v in zip(r.expected_statistics(), m.expected_statistics()):
        assert np.allclose(u, v)
    assert np.allclose(r.lmbda_chol, m.lmbda_chol) and np.allclose(r.sigma, m.sigma)
    npr.seed(11)
    dr = r.rvs()
    npr.seed(11)
    dm = m.rvs()
    assert np.allclose(dr, dm, rtol=1e-12, atol=1e-12)


def test_single_gaussian_entropies(mods):
    ref_d, _, my_d, _ = mods
    rng = np.random.default_rng(4)
    d = 4
    a, b = rng.standard_normal((d, d + 2)), rng.standard_normal((d, d + 2))
    l1, l2 = a @ a.T / d + 0.2 * np.eye(d), b @ b.T / d + 0.2 * np.eye(d)
    m1, m2 = rng.standard_normal(d), rng.standard_normal(d)
    r1, r2 = ref_d.GaussianWithPrecision(d, m1, l1), ref_d.GaussianWithPrecision(d, m2, l2)
    g1, g2 = my_d.GaussianWithPrecision(d, m1, l1), my_d.GaussianWithPrecision(d, m2, l2)
    assert np.isclose(r1.entropy(), g1.entropy(), rtol=1e-12)
    assert np.isclose(r1.cross_entropy(r2), g1.cross_entropy(g2), rtol=1e-12)
