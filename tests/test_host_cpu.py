"""Host-side pieces that need no GPU: Statistics algebra, one_hot, batches, feature
tables, gating / prior closed forms against the oracle."""
import numpy as np
import pytest

from oracle import mimo_oracle as orc


def test_statistics_algebra():
    from mimo_b200.utils.abstraction import Statistics
    a = Statistics([np.ones(3), 2.0, np.eye(2)])
    b = Statistics([np.arange(3.0), 1.0, 2 * np.eye(2)])
    c = a + b
    assert isinstance(c, Statistics) and np.allclose(c[0], [1, 2, 3]) and c[1] == 3.0
    d = 0.5 * (c - a)
    assert np.allclose(d[2], np.eye(2)) and np.allclose((d * 2.0)[0], np.arange(3.0))
    # list-valued entries add element-wise (per-shard statistics)
    e = Statistics([[np.ones(2), np.ones(2)]]) + Statistics([[np.ones(2), 2 * np.ones(2)]])
    assert np.allclose(e[0][1], 3.0)


def test_one_hot_and_batches():
    from mimo_b200.utils.data import one_hot, batches
    z = np.array([0, 2, 1, 2])
    assert np.array_equal(one_hot(z, 3), orc.one_hot(z, 3))
    with pytest.raises(AssertionError):
        one_hot(np.array([0, 3]), 3)
    out = list(batches(5, 20))
    assert len(out) == 1 and len(set(out[0])) == 5


def test_feature_tables():
    from mimo_b200 import _engine as E
    f = E.quad_features(3)
    assert f.F == 10 and f.fi_host[-1] == 3 and f.fj_host[-1] == 3
    assert all(E.tri(i, j) == k for k, (i, j) in enumerate(zip(f.fi_host, f.fj_host)))
    g = E.diag_features(4)
    assert g.F == 9 and list(g.fj_host[:4]) == [4] * 4 and list(g.fi_host[4:8]) == list(g.fj_host[4:8])
    assert E.pad_rows(2) == 8 and E.pad_rows(18) == 32 and E.pad_cols(17) == 20
    with pytest.raises(NotImplementedError):
        E.pad_rows(200)


def test_prior_closed_forms_match_oracle():
    from mimo_b200.distributions import (StackedNormalWisharts, TiedNormalWisharts, StackedNormalGammas,
                                         StackedMatrixNormalWisharts, Dirichlet, TruncatedStickBreaking)
    rng = np.random.default_rng(0)

    def spd(d):
        a = rng.standard_normal((d, d + 2))
        return a @ a.T / d + 0.1 * np.eye(d)
    K, d = 4, 3
    p = (rng.standard_normal((K, d)), rng.random(K) + 0.2, np.stack([spd(d) for _ in range(K)]), d + 1 + rng.random(K))
    nw = StackedNormalWisharts(K, d, *p)
    for a, b in zip(nw.nat_param, orc.nw_std_to_nat(*p)):
        assert np.allclose(a, b)
    for a, b in zip(nw.expected_statistics(), orc.nw_expected_statistics(*p)):
        assert np.allclose(a, b)
    assert np.allclose(nw.log_partition(), orc.nw_log_partition(*p))
    q = StackedNormalWisharts(K, d, *p)
    x, w = rng.standard_normal((50, d)), rng.random((K, 50))
    q.nat_param = nw.nat_param + orc.gauss_full_wstats(x, w)
    assert np.allclose(q.entropy() - q.cross_entropy(nw), orc.nw_vlb(p, q.params))
    t = TiedNormalWisharts(K, d, *p)
    t.nat_param = nw.nat_param
    for a, b in zip(t.params, orc.nw_nat_to_std(orc.nw_std_to_nat(*p), tied=True)):
        assert np.allclose(a, b)
    g = (rng.standard_normal((K, d)), rng.random((K, d)) + 0.2, rng.random((K, d)) + 1, rng.random((K, d)) + 0.5)
    ng = StackedNormalGammas(K, d, *g)
    for a, b in zip(ng.expected_statistics(), orc.ng_expected_statistics(*g)):
        assert np.allclose(a, b)
    assert np.allclose(ng.log_partition(), orc.ng_log_partition(*g))
    o, c = 2, 4
    m = (rng.standard_normal((K, o, c)), np.stack([spd(c) for _ in range(K)]), np.stack([spd(o) for _ in range(K)]),
         o + 1 + rng.random(K))
    mnw = StackedMatrixNormalWisharts(K, c, o, *m)
    for a, b in zip(mnw.nat_param, orc.mnw_std_to_nat(*m)):
        assert np.allclose(a, b)
    for a, b in zip(mnw.expected_statistics(), orc.mnw_expected_statistics(*m)):
        assert np.allclose(a, b)
    back = StackedMatrixNormalWisharts(K, c, o, *m)
    back.nat_param = mnw.nat_param
    for a, b in zip(back.params, m):
        assert np.allclose(a, b)
    al = rng.random(K) + 1.5
    assert np.allclose(Dirichlet(K, al).expected_statistics(), orc.dirichlet_expected_log(al))
    assert np.allclose(Dirichlet(K, al).mode(), orc.dirichlet_mode(al))
    sb = TruncatedStickBreaking(K, al, 2 * al)
    assert np.allclose(sb.mean(), orc.stick_mean(al, 2 * al))
    assert np.allclose(sb.entropy() - sb.cross_entropy(TruncatedStickBreaking(K, np.ones(K), np.ones(K))),
                       orc.stick_vlb((np.ones(K), np.ones(K)), (al, 2 * al)))


def test_bench_algorithmic_work_matches_survey():
    """bench.py's flop / byte model is SURVEY.md 8(d): E-step 2d(d+1)+2d, soft statistics 2(d+1)^2 per pair,
    4*D bytes per point; diagonal 4d+1; linear-Gaussian experts 362 + 362 at d_in = 8, d_out = 1."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location('bench', os.path.join(os.path.dirname(os.path.dirname(__file__)), 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    w5 = bench.algorithmic_work(bench.WORKLOADS['cfg5'])
    assert w5['e_flops_pair'] == 2 * 128 * 129 + 2 * 128 and w5['s_flops_pair'] == 2 * 129 ** 2 and w5['bytes_pt'] == 512
    total = (w5['e_flops_pair'] + w5['s_flops_pair']) * 50_000_000 * 1024
    assert abs(total - 3.4e15) / 3.4e15 < 0.02                      # "flops/sweep 3.4e15" of the SURVEY table
    w3 = bench.algorithmic_work(bench.WORKLOADS['cfg3'])
    assert w3['e_flops_pair'] == 4 * 64 + 1 and w3['bytes_pt'] == 4 * 64 + 4
    w2 = bench.algorithmic_work(bench.WORKLOADS['cfg2'])
    assert w2['e_flops_pair'] == (2 * 8 * 9 + 2 * 8) + (2 * 1 * 9 + 2 + 2 * 81 + 2 + 18) and w2['s_flops_pair'] == 2 * 81 + 2 * 100
    assert w2['e_flops_pair'] == 362 and w2['s_flops_pair'] == 362


# ---- device-resident SVI (mixtures/_svi.py): its closed forms and pseudo-priors are plain torch, checked here on CPU tensors
def _spd(rng, d):
    a = rng.standard_normal((d, d + 2))
    return a @ a.T / d + 0.2 * np.eye(d)


def test_svi_lower_bound_closed_forms_match_oracle():
    import torch
    from oracle import mimo_oracle as orc
    from mimo_b200.mixtures import _svi
    rng = np.random.default_rng(0)
    K, d, o, c = 4, 3, 2, 4
    T = lambda *a: [torch.from_numpy(np.ascontiguousarray(np.asarray(v, dtype=np.float64))) for v in a]      # noqa: E731
    nw_p = (rng.standard_normal((K, d)), rng.random(K) + .1, np.stack([_spd(rng, d) for _ in range(K)]), d + 1. + rng.random(K))
    nw_q = (rng.standard_normal((K, d)), rng.random(K) + 5., np.stack([_spd(rng, d) / 30. for _ in range(K)]), d + 30. + rng.random(K))
    np.testing.assert_allclose(_svi.nw_lower_bound(T(*nw_p), T(*nw_q)).numpy(), orc.nw_vlb(nw_p, nw_q), rtol=1e-10)
    mnw_p = (rng.standard_normal((K, o, c)), np.stack([_spd(rng, c) for _ in range(K)]), np.stack([_spd(rng, o) for _ in range(K)]), o + 1. + rng.random(K))
    mnw_q = (rng.standard_normal((K, o, c)), np.stack([_spd(rng, c) * 20. for _ in range(K)]), np.stack([_spd(rng, o) / 25. for _ in range(K)]),
             o + 25. + rng.random(K))
    np.testing.assert_allclose(_svi.mnw_lower_bound(T(*mnw_p), T(*mnw_q)).numpy(), orc.mnw_vlb(mnw_p, mnw_q), rtol=1e-10)
    a0, a = np.ones(K), 1. + 10. * rng.random(K)
    np.testing.assert_allclose(float(_svi.dirichlet_lower_bound(*T(a0, a))), orc.dirichlet_vlb(a0, a), rtol=1e-10)
    g0, d0, g, dl = np.ones(K), 5. * np.ones(K), 1. + 9. * rng.random(K), 5. + 9. * rng.random(K)
    np.testing.assert_allclose(float(_svi.stick_lower_bound(T(g0, d0), T(g, dl))), orc.stick_vlb((g0, d0), (g, dl)), rtol=1e-10)


def test_svi_pseudo_prior_is_the_natural_parameter_blend():
    """nat_to_std((1 - rho) nat(q) + rho nat(p)): what folds the SVI blend into the conjugate-update kernels."""
    import torch
    from oracle import mimo_oracle as orc
    from mimo_b200.mixtures import _svi
    rng = np.random.default_rng(1)
    K, d, o, c, rho = 3, 4, 2, 3, 0.3
    T = lambda *a: [torch.from_numpy(np.ascontiguousarray(np.asarray(v, dtype=np.float64))) for v in a]      # noqa: E731

    def part(cls, prior, post):
        p = object.__new__(cls)
        p.prior, p.post = T(*prior), T(*post)
        p.prior_psi_inv = torch.linalg.inv(p.prior[2])
        return p
    nw_p = (rng.standard_normal((K, d)), rng.random(K) + .1, np.stack([_spd(rng, d) for _ in range(K)]), d + 1. + rng.random(K))
    nw_q = (rng.standard_normal((K, d)), rng.random(K) + 5., np.stack([_spd(rng, d) / 30. for _ in range(K)]), d + 30. + rng.random(K))
    blend = [(1. - rho) * a + rho * b for a, b in zip(orc.nw_std_to_nat(*nw_q), orc.nw_std_to_nat(*nw_p))]
    for got, ref in zip(part(_svi._NWPart, nw_p, nw_q).pseudo_prior(rho), orc.nw_nat_to_std(blend)):
        np.testing.assert_allclose(got.numpy(), ref, rtol=1e-9, atol=1e-12)
    mnw_p = (rng.standard_normal((K, o, c)), np.stack([_spd(rng, c) for _ in range(K)]), np.stack([_spd(rng, o) for _ in range(K)]), o + 1. + rng.random(K))
    mnw_q = (rng.standard_normal((K, o, c)), np.stack([_spd(rng, c) * 20. for _ in range(K)]), np.stack([_spd(rng, o) / 25. for _ in range(K)]),
             o + 25. + rng.random(K))
    blend = [(1. - rho) * a + rho * b for a, b in zip(orc.mnw_std_to_nat(*mnw_q), orc.mnw_std_to_nat(*mnw_p))]
    for got, ref in zip(part(_svi._MNWPart, mnw_p, mnw_q).pseudo_prior(rho), orc.mnw_nat_to_std(blend)):
        np.testing.assert_allclose(got.numpy(), ref, rtol=1e-9, atol=1e-12)


def test_device_parameter_variates_layout_on_cpu_tensors():
    """draw_gibbs_variates('device') helpers are plain torch: the layout of draw_wishart_variates ([normal(d(d-1)/2) |
    chisquare(nu - i) | normal(extra)]), the Normal-Gamma layout ([gamma(alpha, 1/beta) | normal]) and the gating draws."""
    import torch
    from mimo_b200.distributions.bayesian import (_ComponentsBase, StackedGaussiansWithNormalGammas, CategoricalWithDirichlet,
                                                  CategoricalWithStickBreaking)
    g = torch.Generator().manual_seed(5)
    K, d = 2000, 3
    nus = torch.full((K,), 7.5, dtype=torch.float64)
    v = _ComponentsBase()._wishart_variates_device(nus, d, d, g)
    nt = d * (d - 1) // 2
    assert v.shape == (K, nt + d + d)
    chi = v[:, nt:nt + d].numpy()
    np.testing.assert_allclose(chi.mean(0), 7.5 - np.arange(d), rtol=0.05)          # E chi2(df) = df
    np.testing.assert_allclose(chi.var(0), 2. * (7.5 - np.arange(d)), rtol=0.15)
    assert abs(v[:, :nt].mean()) < 0.05 and abs(v[:, :nt].var() - 1.) < 0.1
    ng = object.__new__(StackedGaussiansWithNormalGammas)
    ng.size, ng.dim, ng.bug_compat, ng._tied = K, d, False, False
    stat = torch.zeros((K, 2 * d + 1), dtype=torch.float64)
    stat[:, 2 * d] = 10.
    stat[:, :d] = 20.
    stat[:, d:2 * d] = 50.
    prior = [torch.zeros((K, d), dtype=torch.float64), torch.full((K, d), 0.1, dtype=torch.float64),
             torch.full((K, d), 2., dtype=torch.float64), torch.full((K, d), 1., dtype=torch.float64)]
    w = ng._draw_variates_device(None, stat, g, prior)
    kap = 0.1 + 10.
    al, be = 2. + 5., 1. + 0.5 * (50. - kap * (20. / kap) ** 2)
    np.testing.assert_allclose(w[:, :d].mean().item(), al / be, rtol=0.05)           # E gamma(alpha, 1 / beta)
    assert abs(w[:, d:].mean().item()) < 0.05
    gate = object.__new__(CategoricalWithDirichlet)
    counts = torch.tensor([3., 0., 5.], dtype=torch.float64)
    gv = gate._draw_variates_device(counts, g, (torch.ones(3, dtype=torch.float64), None))
    assert gv.shape == (3,) and bool((gv > 0).all())
    stick = object.__new__(CategoricalWithStickBreaking)
    sv = stick._draw_variates_device(counts, g, (torch.ones(3, dtype=torch.float64), 2. * torch.ones(3, dtype=torch.float64)))
    assert sv.shape == (2,) and bool(((sv > 0) & (sv < 1)).all())
    g2 = torch.Generator().manual_seed(5)
    assert torch.equal(_ComponentsBase()._wishart_variates_device(nus, d, d, g2), v)   # same seed, same draws: ranks agree without a broadcast
