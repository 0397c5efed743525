"""Parity of every C-ABI kernel with the CPU oracle on seeded inputs (GPU only).

Tolerances (north_star): FP32 mode rel 1e-4 (FP32 compute, FP64 accumulation),
FP64 mode rel 1e-9; labels bit-exact given the same uniforms, excluding draws
within 1e-6 of a CDF boundary.
"""
import numpy as np
import pytest
import torch

from oracle import mimo_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = {'fp32': 1e-4, 'fp64': 1e-9}


def eng():
    from mimo_b200 import _engine
    return _engine


def spd(rng, d, scale=1.0):
    a = rng.standard_normal((d, d + 2))
    return scale * ((a @ a.T) / d + 0.1 * np.eye(d))


def close(a, b, tol, what=''):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(1.0, float(np.max(np.abs(b))))
    err = float(np.max(np.abs(a - b) / (np.abs(b) + scale * 1e-2)))
    assert np.allclose(a, b, rtol=tol, atol=tol * scale), '%s: max scaled err %.3e (tol %.1e)' % (what, err, tol)


def unpack_quad(stat, D):
    """packed lower-triangular (K, F) -> full symmetric (K, D+1, D+1)."""
    K = stat.shape[0]
    S = np.zeros((K, D + 1, D + 1))
    il = np.tril_indices(D + 1)
    S[:, il[0], il[1]] = stat
    S[:, il[1], il[0]] = stat
    return S


@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
@pytest.mark.parametrize('K,d,N', [(4, 2, 500), (9, 16, 1000), (5, 128, 300), (3, 5, 129)])
def test_loglik_quad(precision, K, d, N):
    E = eng()
    rng = np.random.default_rng(d)
    x = rng.standard_normal((N, d)) * 2 + rng.standard_normal(d)
    mus = rng.standard_normal((K, d)) * 2
    lmbdas = np.stack([spd(rng, d) for _ in range(K)])
    logw = np.log(rng.dirichlet(np.ones(K)))
    ops = E.QuadOperands(K, d, d, precision)
    E.set_log_weights(ops, logw)
    info = E.operands_gauss(ops, E.to_dev(mus), E.to_dev(lmbdas))
    info.check()
    Z = E.to_dev(x, E.tdtype(precision))
    ll = E.loglik(Z, ops)
    xr = Z.double().cpu().numpy()          # the oracle sees the same (rounded) data
    ref = orc.gauss_full_loglik(xr, mus, lmbdas) + logw[:, None]
    close(ll, ref, RTOL[precision], 'log-lik')


@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
def test_loglik_diag_and_softmax_labels(precision):
    E = eng()
    rng = np.random.default_rng(3)
    K, d, N = 37, 64, 3000
    x = rng.standard_normal((N, d)) * 1.5
    mus = rng.standard_normal((K, d))
    lam = rng.random((K, d)) + 0.5
    logw = np.log(rng.dirichlet(np.ones(K)))
    ops = E.DiagOperands(K, d, precision)
    E.set_log_weights(ops, logw)
    E.operands_gauss_diag(ops, E.to_dev(mus), E.to_dev(lam))
    Z = E.to_dev(x, E.tdtype(precision))
    ll = E.loglik(Z, ops)
    xr = Z.double().cpu().numpy()
    ref = orc.gauss_diag_loglik(xr, mus, lam) + logw[:, None]
    close(ll, ref, RTOL[precision], 'diag log-lik')
    # label draw with supplied uniforms + lse + resp
    u = np.random.default_rng(5).random(N)
    llh = ll.double().cpu().numpy()
    out = E.softmax(ll, precision, resp=True, lse=True, labels=True, lse_sum=True, uniforms=u)
    lab_ref = orc.sample_discrete_from_log(llh, u)
    safe = orc.label_boundary_distance(llh, u) > (1e-6 if precision == 'fp64' else 1e-4)
    lab = out['labels'].cpu().numpy()
    assert lab.dtype == np.int32
    assert np.array_equal(lab[safe], lab_ref[safe])
    assert safe.mean() > 0.99
    resp_ref, lse_ref = orc.responsibilities(llh)
    close(ll, resp_ref, RTOL[precision] * 10 if precision == 'fp32' else 1e-9, 'resp')
    close(out['lse'], lse_ref, RTOL[precision], 'lse')
    close(out['lse_sum'], [lse_ref.sum()], RTOL[precision], 'lse sum')


@pytest.mark.parametrize('K,d,N,spread', [(37, 64, 3000, 1.0), (256, 64, 5000, 4.0), (200, 40, 2500, 2.0), (5, 8, 700, 0.5),
                                          (64, 33, 1500, 1.5), (129, 30, 257, 1.0), (1, 16, 300, 1.0)])
def test_loglik_diag_tc_labels(K, d, N, spread):
    """tensor-core diagonal E-step (tc_diag.cu): log-joints, log-normalisers and labels against the oracle
    (gaussian.py:837-850, gmm.py:72-75, stats.py:8-21); the data sit away from the origin (the kernel centres them)."""
    E = eng()
    rng = np.random.default_rng(31)
    mus = spread * rng.standard_normal((K, d)) + 7.0
    lam = rng.random((K, d)) * 2 + 0.5
    lab0 = rng.integers(0, K, N)
    x = mus[lab0] + rng.standard_normal((N, d)) / np.sqrt(lam[lab0])
    logw = np.log(rng.dirichlet(np.ones(K)))
    ops = E.DiagOperands(K, d, 'fp32')
    E.set_log_weights(ops, logw)
    E.operands_gauss_diag(ops, E.to_dev(mus), E.to_dev(lam))
    Z = E.to_dev(x, torch.float32)
    xr = Z.double().cpu().numpy()
    ref = orc.gauss_diag_loglik(xr, mus, lam) + logw[:, None]
    u = np.random.default_rng(5).random(N)
    got = E.loglik_diag_tc(Z, ops, out=True, labels=True, lse=True, lse_sum=True, uniforms=u)
    assert got['guard'] == 0
    close(got['out'], ref, 1e-4, 'diag log-joint (tcgen05 feature GEMM)')
    resp_ref, lse_ref = orc.responsibilities(ref)
    close(got['lse'], lse_ref, 1e-4, 'lse')
    close(got['lse_sum'], [lse_ref.sum()], 1e-4, 'lse sum')
    # what the labels feel: the error of the log-joints of the components that carry mass, in nats
    a = got['out'].double().cpu().numpy()
    heavy = resp_ref > 1e-6
    assert np.max(np.abs(a - ref)[heavy]) < 2e-3, np.max(np.abs(a - ref)[heavy])
    lab_ref = orc.sample_discrete_from_log(ref, u)
    safe = orc.label_boundary_distance(ref, u) > 3e-3
    lab = got['labels'].cpu().numpy()
    assert lab.dtype == np.int32 and lab.min() >= 0 and lab.max() < K
    assert np.array_equal(lab[safe], lab_ref[safe]), int((lab[safe] != lab_ref[safe]).sum())
    assert safe.mean() > 0.97
    # labels without the (K, N) output, Philox uniforms keyed by the global index: independent of the chunking
    l1 = E.loglik_diag_tc(Z, ops, labels=True, seed=99, offset=0)['labels'].cpu().numpy()
    cut = min(500, N // 2)
    l2 = E.loglik_diag_tc(Z[cut:], ops, labels=True, seed=99, offset=cut)['labels'].cpu().numpy()
    assert np.array_equal(l1[cut:], l2)


def test_loglik_diag_tc_guard():
    """components far from the common centre relative to their width: the kernel must decline (device-side guard)"""
    E = eng()
    rng = np.random.default_rng(2)
    K, d, N = 16, 32, 512
    mus = 100.0 * rng.standard_normal((K, d))
    lam = np.full((K, d), 4.0)
    ops = E.DiagOperands(K, d, 'fp32')
    E.set_log_weights(ops, np.log(np.full(K, 1.0 / K)))
    E.operands_gauss_diag(ops, E.to_dev(mus), E.to_dev(lam))
    Z = E.to_dev(mus[rng.integers(0, K, N)] + 0.5 * rng.standard_normal((N, d)), torch.float32)
    assert E.loglik_diag_tc(Z, ops, labels=True, seed=1)['guard'] == 1
    # ... and a Gibbs sweep on such operands still matches the oracle (CUDA-core kernels behind the gate)
    feats = E.diag_features(d)
    buf = E.SweepBuffers(N, K, feats.F, 'fp32', hard=True)
    u = rng.random(N)
    E.sweep(Z, ops, feats, buf, uniforms=E.to_dev(u))
    xr = Z.double().cpu().numpy()
    ref = orc.gauss_diag_loglik(xr, mus, lam) + np.log(1.0 / K)
    safe = orc.label_boundary_distance(ref, u) > 1e-3
    assert np.array_equal(buf.labels.cpu().numpy()[safe], orc.sample_discrete_from_log(ref, u)[safe])


@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
@pytest.mark.parametrize('K,din,o', [(6, 1, 1), (17, 8, 1), (5, 3, 2)])
def test_predict_lingauss_kernel(precision, K, din, o):
    """mimo_predict_lingauss / mimo_studentt_from_quad against the oracle's restatement of ilr.py:325-430."""
    E = eng()
    rng = np.random.default_rng(K + din)
    N, c = 300, din + 1
    x = rng.standard_normal((N, din)) * 2
    y = rng.standard_normal((N, o))
    Ms = rng.standard_normal((K, o, c))
    Ks = np.stack([spd(rng, c) for _ in range(K)])
    psis = np.stack([spd(rng, o) for _ in range(K)])
    nus = o + 2.5 + 3 * rng.random(K)
    w = rng.dirichlet(np.ones(K), size=N).T
    tol = 1e-9 if precision == 'fp64' else 1e-4
    dt = E.tdtype(precision)
    X, W, Y = E.to_dev(x, dt), E.to_dev(w, dt), E.to_dev(y, dt)
    xr, wr, yr = X.double().cpu().numpy(), W.double().cpu().numpy(), Y.double().cpu().numpy()
    for dist in ('gaussian', 'studentt'):
        for mode, pred in ((0, 'average'), (1, 'mode')):
            mu, cov, nlpd = E.predict_lingauss(X, W, Ms, np.linalg.inv(Ks), np.linalg.inv(psis), psis, np.linalg.slogdet(psis)[1],
                                               nus - o + 1, True, mode, dist == 'studentt', precision, Y=Y, eps=1e-300)
            mu_r, cov_r, nlpd_r = orc.ilr_prediction(xr, wr, (Ms, Ks, psis, nus), pred, dist, y=yr, eps=1e-300)
            close(mu, mu_r, tol, 'mean %s %s' % (dist, pred))
            close(cov, cov_r, tol, 'covariance %s %s' % (dist, pred))
            close(nlpd, nlpd_r, tol, 'nlpd %s %s' % (dist, pred))
    # Student-t form of a Gaussian-form log-joint
    d = din
    mus = rng.standard_normal((K, d))
    lmbdas = np.stack([spd(rng, d) for _ in range(K)])
    dfs = 3.0 + rng.random(K)
    ops = E.QuadOperands(K, d, d, precision)
    E.set_log_weights(ops, np.zeros(K))
    E.operands_gauss(ops, E.to_dev(mus), E.to_dev(lmbdas)).check()
    a = E.loglik(X, ops)
    half_logdet = 0.5 * np.linalg.slogdet(lmbdas)[1]
    from scipy.special import gammaln
    aux = gammaln((dfs + d) / 2.) - gammaln(dfs / 2.) + half_logdet - (d / 2.) * np.log(dfs * np.pi) - 0.5 * (dfs + d)
    E.studentt_from_quad(a, precision, half_logdet - 0.5 * d * np.log(2 * np.pi), aux, dfs)
    close(a, orc.studentt_loglik_reference_form(xr, mus, lmbdas, dfs), tol * 10, 'Student-t form')


def test_philox_labels_independent_of_chunking():
    E = eng()
    rng = np.random.default_rng(8)
    K, N = 6, 4096
    a = rng.standard_normal((K, N))
    t1 = E.to_dev(a)
    l1 = E.softmax(t1, 'fp64', labels=True, seed=99, offset=0)['labels'].cpu().numpy()
    t2 = E.to_dev(a[:, 1000:])
    l2 = E.softmax(t2, 'fp64', labels=True, seed=99, offset=1000)['labels'].cpu().numpy()
    assert np.array_equal(l1[1000:], l2)
    counts = np.bincount(l1, minlength=K) / N
    p = orc.responsibilities(a)[0].mean(1)
    assert np.max(np.abs(counts - p)) < 0.03


@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
@pytest.mark.parametrize('K,d,N', [(4, 2, 700), (70, 16, 2100), (3, 128, 400)])
def test_stats_quad(precision, K, d, N):
    E = eng()
    rng = np.random.default_rng(K)
    x = rng.standard_normal((N, d)) + 2.0
    w = rng.random((K, N))
    w /= w.sum(0)
    Z = E.to_dev(x, E.tdtype(precision))
    R = E.to_dev(w, E.tdtype(precision))
    feats = E.quad_features(d)
    st = E.stats_soft(Z, R, feats, precision).cpu().numpy()
    xr, wr = Z.double().cpu().numpy(), R.double().cpu().numpy()
    ref = orc.gauss_full_wstats(xr, wr)
    S = unpack_quad(st, d)
    close(S[:, :d, :d], ref[2], RTOL[precision], 'sum r xx')
    close(S[:, d, :d], ref[0], RTOL[precision], 'sum r x')
    close(S[:, d, d], ref[1], RTOL[precision], 'sum r')
    labels = rng.integers(0, K, size=N).astype(np.int32)
    sh = E.stats_hard(Z, E.to_dev(labels, torch.int32), K, feats, precision).cpu().numpy()
    ref = orc.gauss_full_wstats(xr, orc.one_hot(labels, K))
    S = unpack_quad(sh, d)
    # FP32 data with d >= 8 goes through the register-tiled pair-list kernel (FP32 slab sums, FP64 across slabs)
    htol = 1e-10 if precision == 'fp64' else 1e-5
    close(S[:, :d, :d], ref[2], htol, 'hard sum xx')
    close(S[:, d, :d], ref[0], htol, 'hard sum x')
    assert np.array_equal(S[:, d, d], np.bincount(labels, minlength=K))


@pytest.mark.parametrize('K,d,N', [(3, 8, 5000), (5, 17, 3000), (40, 24, 9000), (7, 33, 2500), (9, 64, 4000),
                                   (6, 100, 3000), (4, 128, 6000), (300, 128, 2000)])
def test_stats_hard_pair_list(K, d, N):
    """pair-list statistics kernel (pair_stats.cu) behind mimo_stats_hard: uneven and empty components, lists longer
    than one slab, every column-group count.  Reference: gaussian.py:491-505 on the one-hot matrix of data.py:160-169."""
    E = eng()
    rng = np.random.default_rng(K * 1000 + d)
    x = rng.standard_normal((N, d)) * (1.0 + rng.random(d)) + rng.standard_normal(d)
    p = rng.dirichlet(0.3 * np.ones(K))
    p[K // 2] = 0.0                                        # an empty component
    labels = rng.choice(K, size=N, p=p / p.sum()).astype(np.int32)
    Z = E.to_dev(x, torch.float32)
    feats = E.quad_features(d)
    sh = E.stats_hard(Z, E.to_dev(labels, torch.int32), K, feats, 'fp32').cpu().numpy()
    ref = orc.gauss_full_wstats(Z.double().cpu().numpy(), orc.one_hot(labels, K))
    S = unpack_quad(sh, d)
    close(S[:, :d, :d], ref[2], 1e-5, 'pair-list sum xx')
    close(S[:, d, :d], ref[0], 1e-5, 'pair-list sum x')
    assert np.array_equal(S[:, d, d], np.bincount(labels, minlength=K))
    assert np.all(S[K // 2] == 0.0)


def test_stats_hard_rejects_bad_labels():
    E = eng()
    Z = E.to_dev(np.zeros((10, 3)), torch.float32)
    lab = E.to_dev(np.array([0, 1, 2, 3, 0, 1, 2, 7, 0, 1], dtype=np.int32), torch.int32)
    with pytest.raises(AssertionError):
        E.stats_hard(Z, lab, 4, E.quad_features(3), 'fp32')


@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
def test_stats_diag(precision):
    E = eng()
    rng = np.random.default_rng(2)
    K, d, N = 11, 64, 1500
    x = rng.standard_normal((N, d)) + 1.0
    w = rng.random((K, N))
    Z, R = E.to_dev(x, E.tdtype(precision)), E.to_dev(w, E.tdtype(precision))
    feats = E.diag_features(d)
    st = E.stats_soft(Z, R, feats, precision).cpu().numpy()
    ref = orc.gauss_diag_wstats(Z.double().cpu().numpy(), R.double().cpu().numpy())
    close(st[:, :d], ref[0], RTOL[precision], 'sum r x')
    close(st[:, d:2 * d], ref[3], RTOL[precision], 'sum r x^2')
    close(st[:, 2 * d], ref[1][:, 0], RTOL[precision], 'sum r')


def nw_prior(rng, K, d):
    return (rng.standard_normal((K, d)), rng.random(K) + 0.1,
            np.stack([spd(rng, d) for _ in range(K)]), d + 1.0 + 3 * rng.random(K))


@pytest.mark.parametrize('tied', [False, True])
@pytest.mark.parametrize('K,d', [(5, 3), (3, 16), (2, 128)])
def test_nw_posterior_meanfield(tied, K, d):
    E = eng()
    rng = np.random.default_rng(d + K)
    N = 4 * d + 50
    prior = nw_prior(rng, K, d)
    x = rng.standard_normal((N, d)) * 1.5 + 1.0
    w = rng.random((K, N))
    w /= w.sum(0)
    feats = E.quad_features(d)
    Z, R = E.to_dev(x), E.to_dev(w)
    stat = E.stats_soft(Z, R, feats, 'fp64')
    ops = E.QuadOperands(K, d, d, 'fp64')
    out = E.nw_posterior([E.to_dev(p) for p in prior], stat, feats.F, E.identity_map(d, d), d + 1,
                         mode=0, tied=tied, ops=ops, col_map=E.identity_map(d, d))
    out['info'].check()
    post = orc.nw_nat_to_std(orc.add_stats(orc.nw_std_to_nat(*prior), orc.gauss_full_wstats(x, w)), tied=tied)
    for key, ref in zip(('m', 'kappa', 'psi', 'nu'), post):
        close(out[key], ref, 1e-9, 'posterior ' + key)
    close(out['vlb'], orc.nw_vlb(prior, post), 1e-8, 'vlb term')
    ell = E.loglik(Z, ops)
    close(ell, orc.nw_expected_loglik(x, *post), 1e-9, 'expected log-lik')
    # MAP operands: Gaussian log-lik at the posterior mode
    ops2 = E.QuadOperands(K, d, d, 'fp64')
    out2 = E.nw_posterior([E.to_dev(p) for p in prior], stat, feats.F, E.identity_map(d, d), d + 1,
                          mode=2, tied=tied, ops=ops2, col_map=E.identity_map(d, d), want_lik=True)
    mu_m, l_m = orc.nw_mode(*post)
    close(out2['lik_mu'], mu_m, 1e-9, 'mode mu')
    close(out2['lik_lmbda'], l_m, 1e-9, 'mode lmbda')
    close(E.loglik(Z, ops2), orc.gauss_full_loglik(x, mu_m, l_m), 1e-9, 'MAP log-lik')


@pytest.mark.parametrize('K,d', [(4, 2), (3, 16), (2, 128)])
def test_nw_posterior_gibbs_replay(K, d):
    """Sampled (mu, lmbda) from the reference's variates, and the log-lik operands built from them."""
    E = eng()
    rng = np.random.default_rng(7 * d)
    N = 4 * d + 40
    prior = nw_prior(rng, K, d)
    x = rng.standard_normal((N, d)) + 0.5
    labels = rng.integers(0, K, size=N).astype(np.int32)
    feats = E.quad_features(d)
    Z = E.to_dev(x)
    stat = E.stats_hard(Z, E.to_dev(labels, torch.int32), K, feats, 'fp64')
    post = orc.nw_nat_to_std(orc.add_stats(orc.nw_std_to_nat(*prior), orc.gauss_full_wstats(x, orc.one_hot(labels, K))))
    nt = d * (d - 1) // 2
    var = np.zeros((K, nt + 2 * d))
    mus_ref, lm_ref = np.zeros((K, d)), np.zeros((K, d, d))
    for k in range(K):
        var[k, :nt] = rng.standard_normal(nt)
        var[k, nt:nt + d] = rng.chisquare(post[3][k] - np.arange(d))
        var[k, nt + d:] = rng.standard_normal(d)
        mus_ref[k], lm_ref[k] = orc.nw_rvs_from_variates(post[0][k], post[1][k], post[2][k], post[3][k],
                                                        var[k, :nt], var[k, nt:nt + d], var[k, nt + d:])
    ops = E.QuadOperands(K, d, d, 'fp64')
    out = E.nw_posterior([E.to_dev(p) for p in prior], stat, feats.F, E.identity_map(d, d), d + 1,
                         mode=1, variates=var, ops=ops, col_map=E.identity_map(d, d), want_lik=True)
    out['info'].check()
    close(out['lik_mu'], mus_ref, 1e-8, 'sampled mu')
    close(out['lik_lmbda'], lm_ref, 1e-8, 'sampled lmbda')
    close(E.loglik(Z, ops), orc.gauss_full_loglik(x, mus_ref, lm_ref), 1e-8, 'Gibbs log-lik')


def test_nw_not_positive_definite_raises():
    E = eng()
    K, d = 2, 3
    prior = [np.zeros((K, d)), np.ones(K), np.stack([np.eye(d), -np.eye(d)]), 5.0 * np.ones(K)]
    feats = E.quad_features(d)
    out = E.nw_posterior([E.to_dev(p) for p in prior], E.zeros((K, feats.F)), feats.F, E.identity_map(d, d), d + 1, mode=3)
    with pytest.raises(np.linalg.LinAlgError):
        out['info'].check()


@pytest.mark.parametrize('tied,bug', [(False, False), (True, False), (False, True)])
def test_ng_posterior(tied, bug):
    E = eng()
    rng = np.random.default_rng(12)
    K, d, N = 6, 9, 300
    prior = (rng.standard_normal((K, d)), rng.random((K, d)) + 0.1, rng.random((K, d)) + 1.0, rng.random((K, d)) + 0.5)
    x = rng.standard_normal((N, d)) + 0.3
    w = rng.random((K, N))
    w /= w.sum(0)
    feats = E.diag_features(d)
    Z, R = E.to_dev(x), E.to_dev(w)
    stat = E.stats_soft(Z, R, feats, 'fp64')
    ops = E.DiagOperands(K, d, 'fp64')
    out = E.ng_posterior([E.to_dev(p) for p in prior], stat, feats.F, mode=0, tied=tied, bug_compat=bug, ops=ops)
    post = orc.ng_nat_to_std(orc.add_stats(orc.ng_std_to_nat(*prior), orc.gauss_diag_wstats(x, w)), tied=tied)
    if bug:
        post = (post[0], post[1], prior[2], prior[3])
    for key, ref in zip(('m', 'kappa', 'alpha', 'beta'), post):
        close(out[key], ref, 1e-10, 'NG posterior ' + key)
    close(E.loglik(Z, ops), orc.ng_expected_loglik(x, *post), 1e-9, 'NG expected log-lik')
    if not tied:
        close(out['vlb'], orc.ng_vlb(prior, post), 1e-8, 'NG vlb')
    # Gibbs replay
    g = rng.gamma(post[2], 1.0 / post[3])
    z = rng.standard_normal((K, d))
    ops2 = E.DiagOperands(K, d, 'fp64')
    out2 = E.ng_posterior([E.to_dev(p) for p in prior], stat, feats.F, mode=1, tied=tied, bug_compat=bug,
                          variates=np.hstack((g, z)), ops=ops2, want_lik=True)
    mu_s, l_s = orc.ng_rvs_from_variates(post[0], post[1], post[2], post[3], g, z)
    close(out2['lik_mu'], mu_s, 1e-10, 'NG sampled mu')
    close(E.loglik(Z, ops2), orc.gauss_diag_loglik(x, mu_s, l_s), 1e-9, 'NG Gibbs log-lik')


def ilr_layout(E, din, o):
    """zt = [x (din) | y (o) | 1]; xt = [x ; 1] -> columns 0..din-1 and D."""
    D = din + o
    c = din + 1
    x_cols = list(range(din)) + [D]
    y_cols = list(range(din, din + o))
    mnw_idx = E._i32(x_cols + y_cols + [D])      # stat positions: xt, y, constant
    mnw_map = E._i32(x_cols + y_cols)            # operand columns: xt then y
    nw_idx = E._i32(list(range(din)) + [D])
    return D, c, mnw_idx, mnw_map, nw_idx


@pytest.mark.parametrize('tied', [False, True])
@pytest.mark.parametrize('din,o', [(2, 1), (8, 1), (3, 2)])
def test_mnw_posterior_and_ilr_operands(tied, din, o):
    E = eng()
    rng = np.random.default_rng(din * 10 + o)
    K, N = 5, 400
    D, c, mnw_idx, mnw_map, nw_idx = ilr_layout(E, din, o)
    bprior = nw_prior(rng, K, din)
    mprior = (rng.standard_normal((K, o, c)) * 0.3, np.stack([spd(rng, c) for _ in range(K)]),
              np.stack([spd(rng, o) for _ in range(K)]), o + 1.0 + 3 * rng.random(K))
    x = rng.standard_normal((N, din))
    y = rng.standard_normal((N, o)) + x[:, :1]
    w = rng.random((K, N))
    w /= w.sum(0)
    z = np.hstack((x, y))
    feats = E.quad_features(D)
    Z, R = E.to_dev(z), E.to_dev(w)
    stat = E.stats_soft(Z, R, feats, 'fp64')
    ops = E.QuadOperands(K, D, din + o + c, 'fp64')
    ob = E.nw_posterior([E.to_dev(p) for p in bprior], stat, feats.F, nw_idx, D + 1, mode=0, ops=ops,
                        row_off=0, col_map=nw_idx)
    om = E.mnw_posterior([E.to_dev(p) for p in mprior], stat, feats.F, mnw_idx, D + 1, mode=0, tied=tied, ops=ops,
                         row_off=din, col_map=mnw_map)
    ob['info'].check()
    om['info'].check()
    bpost = orc.nw_nat_to_std(orc.add_stats(orc.nw_std_to_nat(*bprior), orc.gauss_full_wstats(x, w)))
    mpost = orc.mnw_nat_to_std(orc.add_stats(orc.mnw_std_to_nat(*mprior), orc.lingauss_wstats(x, y, w)), tied=tied)
    for key, ref in zip(('M', 'K', 'psi', 'nu'), mpost):
        close(om[key], ref, 1e-9, 'MNW posterior ' + key)
    close(om['vlb'], orc.mnw_vlb(mprior, mpost), 1e-8, 'MNW vlb')
    ref = orc.nw_expected_loglik(x, *bpost) + orc.mnw_expected_loglik(x, y, *mpost)
    close(E.loglik(Z, ops), ref, 1e-9, 'ILR expected log-lik')
    # Gibbs replay of the experts
    nt = o * (o - 1) // 2
    var = np.zeros((K, nt + o + o * c))
    As, lms = np.zeros((K, o, c)), np.zeros((K, o, o))
    for k in range(K):
        var[k, :nt] = rng.standard_normal(nt)
        var[k, nt:nt + o] = rng.chisquare(mpost[3][k] - np.arange(o))
        var[k, nt + o:] = rng.standard_normal(o * c)
        As[k], lms[k] = orc.mnw_rvs_from_variates(mpost[0][k], mpost[1][k], mpost[2][k], mpost[3][k],
                                                  var[k, :nt], var[k, nt:nt + o], var[k, nt + o:])
    ops2 = E.QuadOperands(K, D, o, 'fp64')
    og = E.mnw_posterior([E.to_dev(p) for p in mprior], stat, feats.F, mnw_idx, D + 1, mode=1, tied=tied,
                         variates=var, ops=ops2, row_off=0, col_map=mnw_map, want_lik=True)
    og['info'].check()
    close(og['lik_A'], As, 1e-8, 'sampled A')
    close(og['lik_lmbda'], lms, 1e-8, 'sampled lmbda')
    close(E.loglik(Z, ops2), orc.lingauss_loglik(x, y, As, lms), 1e-8, 'lin-Gauss log-lik')
    # explicit-parameter operands + M-step
    ops3 = E.QuadOperands(K, D, o, 'fp64')
    E.operands_lingauss(ops3, E.to_dev(As), E.to_dev(lms), 0, mnw_map).check()
    close(E.loglik(Z, ops3), orc.lingauss_loglik(x, y, As, lms), 1e-9, 'lin-Gauss operands')
    A_m, l_m, info = E.mstep_lingauss(stat, feats.F, mnw_idx, D + 1, K, c, o)
    info.check()
    A_r, l_r = orc.lingauss_mstep(orc.lingauss_wstats(x, y, w))
    close(A_m, A_r, 1e-8, 'M-step A')
    close(l_m, l_r, 1e-7, 'M-step lmbda')


@pytest.mark.parametrize('kind', [0, 1])
def test_gating(kind):
    E = eng()
    rng = np.random.default_rng(4)
    K = 13
    counts = rng.random(K) * 20
    stat = E.to_dev(counts[:, None].copy())
    a0 = np.ones(K) * 1.5
    b0 = np.ones(K) * 5.0
    pa, pb = E.to_dev(a0), (E.to_dev(b0) if kind == 1 else None)
    ops = E.DiagOperands(K, 1, 'fp64')
    out = E.gating_posterior(kind, pa, pb, stat, 1, 0, mode=0, ops=ops)
    if kind == 0:
        post = orc.dirichlet_posterior(a0, counts)
        close(out['a'], post, 1e-12)
        close(ops.cst, orc.dirichlet_expected_log(post), 1e-10, 'E log pi')
        close(out['vlb'], [orc.dirichlet_vlb(a0, post)], 1e-9, 'dirichlet vlb')
        g = rng.gamma(post)
        E.gating_posterior(kind, pa, pb, stat, 1, 0, mode=1, variates=g, ops=ops)
        close(ops.cst, np.log(orc.dirichlet_probs_from_gammas(g)), 1e-12, 'Gibbs log pi')
        E.gating_posterior(kind, pa, pb, stat, 1, 0, mode=2, ops=ops)
        close(ops.cst, np.log(orc.dirichlet_mode(post)), 1e-12, 'mode')
    else:
        gp, dp = orc.stick_posterior(a0, b0, counts)
        close(out['a'], gp, 1e-12)
        close(out['b'], dp, 1e-12)
        close(ops.cst, orc.stick_expected_log(gp, dp)[0], 1e-10, 'E log pi (stick)')
        close(out['vlb'], [orc.stick_vlb((a0, b0), (gp, dp))], 1e-9, 'stick vlb')
        v = rng.beta(gp[:-1], dp[:-1])
        o2 = E.gating_posterior(kind, pa, pb, stat, 1, 0, mode=1, variates=v, ops=ops)
        close(o2['probs'], orc.stick_probs_from_betas(v), 1e-12, 'stick probs')
        o3 = E.gating_posterior(kind, pa, pb, stat, 1, 0, mode=4, ops=ops)
        close(o3['probs'], orc.stick_mean(gp, dp), 1e-12, 'stick mean')


@pytest.mark.parametrize('tied', [False, True])
def test_msteps(tied):
    E = eng()
    rng = np.random.default_rng(6)
    K, d, N = 4, 5, 500
    x = rng.standard_normal((N, d)) + rng.standard_normal(d)
    w = rng.random((K, N))
    w /= w.sum(0)
    Z, R = E.to_dev(x), E.to_dev(w)
    feats = E.quad_features(d)
    stat = E.stats_soft(Z, R, feats, 'fp64')
    mu, lm, info = E.mstep_gauss(stat, feats.F, E.identity_map(d, d), d + 1, K, d, tied=tied)
    info.check()
    st = orc.gauss_full_wstats(x, w)
    if not tied:
        mu_r, lm_r = orc.gauss_full_mstep(st)
    else:
        mu_r = st[0] / st[1][:, None]
        sig = (st[2].sum(0) - np.einsum('k,kd,kl->dl', st[1], mu_r, mu_r)) / st[1].sum()
        lm_r = np.stack(K * [np.linalg.inv(sig)])
    close(mu, mu_r, 1e-10, 'M-step mu')
    close(lm, lm_r, 1e-8, 'M-step lmbda')
    fd = E.diag_features(d)
    sd = E.stats_soft(Z, R, fd, 'fp64')
    mu_d, lam_d = E.mstep_gauss_diag(sd, fd.F, K, d, tied=tied)
    sr = orc.gauss_diag_wstats(x, w)
    if not tied:
        mu_r, lam_r = orc.gauss_diag_mstep(sr)
    else:
        mu_r = sr[0] / sr[1]
        lam_r = np.stack(K * [1.0 / ((sr[3].sum(0) - (sr[1] * mu_r ** 2).sum(0)) / sr[1][:, 0].sum() + 1e-16)])
    close(mu_d, mu_r, 1e-10, 'diag M-step mu')
    close(lam_d, lam_r, 1e-9, 'diag M-step lmbda')


@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
@pytest.mark.parametrize('hard', [False, True])
def test_sweep_matches_separate_kernels(precision, hard):
    """mimo_sweep (chunked, fused) == loglik -> softmax -> stats on the whole block."""
    E = eng()
    rng = np.random.default_rng(20)
    K, d, N = 12, 6, 70000
    x = rng.standard_normal((N, d)) + rng.integers(0, 3, size=(N, 1))
    mus = rng.standard_normal((K, d)) * 2
    lmbdas = np.stack([spd(rng, d) for _ in range(K)])
    ops = E.QuadOperands(K, d, d, precision)
    E.set_log_weights(ops, np.log(rng.dirichlet(np.ones(K))))
    E.operands_gauss(ops, E.to_dev(mus), E.to_dev(lmbdas)).check()
    Z = E.to_dev(x, E.tdtype(precision))
    feats = E.quad_features(d)
    u = E.to_dev(rng.random(N))
    buf = E.SweepBuffers(N, K, feats.F, precision, hard)
    E.sweep(Z, ops, feats, buf, uniforms=u if hard else None)
    ll = E.loglik(Z, ops)
    if hard:
        out = E.softmax(ll, precision, labels=True, lse_sum=True, uniforms=u)
        assert torch.equal(out['labels'], buf.labels)
        ref = E.stats_hard(Z, out['labels'], K, feats, precision)
        close(buf.stat, ref.cpu().numpy(), 1e-12, 'hard stats')
    else:
        out = E.softmax(ll, precision, resp=True, lse_sum=True)
        ref = E.stats_soft(Z, ll, feats, precision)
        close(buf.stat, ref.cpu().numpy(), 1e-6 if precision == 'fp32' else 1e-12, 'soft stats')
    close(buf.lse_sum, out['lse_sum'].cpu().numpy(), 1e-12, 'lse sum')


def test_digamma_device_matches_scipy():
    """digamma/lgamma device functions through the Dirichlet expected log-weights."""
    from scipy.special import digamma
    E = eng()
    for scale in (1e-3, 1.0, 50.0, 1e5):
        a = np.random.default_rng(1).random(64) * scale + 1e-6
        ops = E.DiagOperands(64, 1, 'fp64')
        E.gating_posterior(0, E.to_dev(a), None, E.zeros((64, 1)), 1, 0, mode=0, ops=ops)
        ref = digamma(a) - digamma(a.sum())
        got = ops.cst.cpu().numpy()
        assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) < 1e-12


@pytest.mark.parametrize('hard', [False, True])
@pytest.mark.parametrize('K,d,N,segment', [(12, 6, 70000, 0), (12, 6, 300000, 100000), (40, 32, 50000, 20000)])
def test_sweep_host_matches_resident_sweep(hard, K, d, N, segment):
    """mimo_sweep_host (host buffers, segmented upload overlapped with the sweep) == mimo_sweep on resident data:
    statistics and the lower-bound term are sums over segments, labels are per point (global Philox / uniform index)."""
    from mimo_b200 import _lib
    E = eng()
    rng = np.random.default_rng(K + d)
    x = (rng.standard_normal((N, d)) + rng.integers(0, 3, size=(N, 1))).astype(np.float32)
    mus = rng.standard_normal((K, d)) * 2
    lmbdas = np.stack([spd(rng, d) for _ in range(K)])
    ops = E.QuadOperands(K, d, d, 'fp32')
    E.set_log_weights(ops, np.log(rng.dirichlet(np.ones(K))))
    E.operands_gauss(ops, E.to_dev(mus), E.to_dev(lmbdas)).check()
    Z = E.to_dev(x, torch.float32)
    feats = E.quad_features(d)
    u = rng.random(N)
    buf = E.SweepBuffers(N, K, feats.F, 'fp32', hard)
    E.sweep(Z, ops, feats, buf, uniforms=E.to_dev(u) if hard else None)
    torch.cuda.synchronize()
    W, cst = ops.W.cpu().contiguous(), ops.cst.cpu().contiguous()
    stat_h = np.zeros((K, feats.F))
    lse_h = np.zeros(1)
    lab_h = np.zeros(N, dtype=np.int32)
    _lib.call('mimo_sweep_host_set_segment', segment)
    try:
        _lib.call('mimo_sweep_host', 0, 0, 1 if hard else 0, x.ctypes.data, N, d, W.data_ptr(), None, cst.data_ptr(),
                  ops.K, ops.Rp, ops.Dpp, feats.fi_host.ctypes.data, feats.fj_host.ctypes.data, feats.F,
                  u.ctypes.data if hard else None, 5, stat_h.ctypes.data, lse_h.ctypes.data, lab_h.ctypes.data if hard else None)
    finally:
        _lib.call('mimo_sweep_host_set_segment', 0)
        _lib.call('mimo_sweep_host_release')
    close(stat_h, buf.stat.cpu().numpy(), 1e-5, 'host-buffer sweep statistics')
    close(lse_h, buf.lse_sum.cpu().numpy().reshape(1), 1e-6, 'host-buffer sweep lse sum')
    if hard:
        ref = buf.labels.cpu().numpy()
        if d < 24:
            assert np.array_equal(lab_h, ref)
        else:   # tensor-core path: the FP16-split scale is per segment, draws on a CDF boundary may move
            assert (lab_h == ref).mean() > 0.9995


@pytest.mark.parametrize('K,d,N,sep', [(20, 16, 60000, 6.0), (20, 16, 30000, 0.2), (130, 9, 40000, 3.0), (8, 8, 5000, 5.0)])
def test_sweep_resp_list_statistics(K, d, N, sep):
    """CUDA-core FP32 sweep, 8 <= d < 24: statistics summed over the pairs with r >= e^-40 (pair-list kernel) when
    they are few, dense CUDA-core statistics otherwise (device-side choice).  Both must match the plain dense sweep
    (mimo_set_tensor_cores(0)) and the oracle (gaussian.py:491-505)."""
    E = eng()
    rng = np.random.default_rng(K * d)
    centres = sep * rng.standard_normal((K, d))
    z = rng.integers(0, K, size=N)
    x = centres[z] + rng.standard_normal((N, d))
    mus = centres + 0.1 * rng.standard_normal((K, d))
    lmbdas = np.stack([spd(rng, d) for _ in range(K)])
    logw = np.log(rng.dirichlet(np.ones(K)))
    ops = E.QuadOperands(K, d, d, 'fp32')
    E.set_log_weights(ops, logw)
    E.operands_gauss(ops, E.to_dev(mus), E.to_dev(lmbdas)).check()
    Z = E.to_dev(x, torch.float32)
    feats = E.quad_features(d)
    buf = E.SweepBuffers(N, K, feats.F, 'fp32', False)
    old_min = E.set_tc_min_dim(24)                 # keep these dimensions on the CUDA cores (since round 2 the default sends d >= 8 to the tensor pipe)
    try:
        E.sweep(Z, ops, feats, buf)
    finally:
        E.set_tc_min_dim(old_min)
    old = E.set_tensor_cores(0)
    try:
        ref = E.SweepBuffers(N, K, feats.F, 'fp32', False)
        E.sweep(Z, ops, feats, ref)
    finally:
        E.set_tensor_cores(old)
    close(buf.stat, ref.stat.cpu().numpy(), 1e-5, 'list vs dense statistics')
    close(buf.lse_sum, ref.lse_sum.cpu().numpy(), 1e-9, 'lse sum')
    xr = Z.double().cpu().numpy()
    resp, lse = orc.responsibilities(orc.gauss_full_loglik(xr, mus, lmbdas) + logw[:, None])
    st = orc.gauss_full_wstats(xr, resp)
    S = unpack_quad(buf.stat.cpu().numpy(), d)
    close(S[:, :d, :d], st[2], 1e-4, 'sum r xx')
    close(S[:, d, :d], st[0], 1e-4, 'sum r x')
    close(S[:, d, d], st[1], 1e-4, 'sum r')


@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
@pytest.mark.parametrize('K,d,N', [(7, 64, 9000), (300, 5, 4000), (3, 300, 2500), (256, 64, 40000)])
def test_stats_hard_diag(precision, K, d, N):
    """hard statistics of the diagonal family (gaussian.py:819-832 on the one-hot matrix of data.py:160-169): FP32 goes
    through the streaming kernel over the label-sorted lists, FP64 through the generic per-feature kernel."""
    E = eng()
    rng = np.random.default_rng(K + d)
    x = rng.standard_normal((N, d)) * (0.5 + rng.random(d)) + rng.standard_normal(d)
    p = rng.dirichlet(0.3 * np.ones(K))
    p[K // 2] = 0.0
    labels = rng.choice(K, size=N, p=p / p.sum()).astype(np.int32)
    Z = E.to_dev(x, E.tdtype(precision))
    feats = E.diag_features(d)
    st = E.stats_hard(Z, E.to_dev(labels, torch.int32), K, feats, precision).cpu().numpy()
    ref = orc.gauss_diag_wstats(Z.double().cpu().numpy(), orc.one_hot(labels, K))
    tol = 1e-10 if precision == 'fp64' else 1e-5
    close(st[:, :d], ref[0], tol, 'hard sum x')
    close(st[:, d:2 * d], ref[3], tol, 'hard sum x^2')
    assert np.array_equal(st[:, 2 * d], np.bincount(labels, minlength=K))
    assert np.all(st[K // 2] == 0.0)


@pytest.mark.parametrize('D', [9, 16])
def test_stats_hard_non_canonical_feature_table(D):
    """A feature table with the SIZE of a canonical layout but a different order must not take a fast kernel: the layout
    is verified on the device and the generic per-feature kernel runs."""
    E = eng()
    rng = np.random.default_rng(D)
    N, K = 3000, 4
    x = rng.standard_normal((N, D)) + 1.0
    labels = rng.integers(0, K, size=N).astype(np.int32)
    canon = E.quad_features(D)
    perm = rng.permutation(canon.F)
    feats = E.Features(canon.fi_host[perm], canon.fj_host[perm], D)
    Z = E.to_dev(x, torch.float32)
    st = E.stats_hard(Z, E.to_dev(labels, torch.int32), K, feats, 'fp32').cpu().numpy()
    zt = np.concatenate([Z.double().cpu().numpy(), np.ones((N, 1))], axis=1)
    phi = zt[:, feats.fi_host] * zt[:, feats.fj_host]
    ref = orc.one_hot(labels, K) @ phi
    close(st, ref, 1e-10, 'permuted feature table')
    # diagonal-sized table (F = 2 D + 1) in a different order
    dcanon = E.diag_features(D)
    perm = rng.permutation(dcanon.F)
    feats = E.Features(dcanon.fi_host[perm], dcanon.fj_host[perm], D)
    st = E.stats_hard(Z, E.to_dev(labels, torch.int32), K, feats, 'fp32').cpu().numpy()
    phi = zt[:, feats.fi_host] * zt[:, feats.fj_host]
    close(st, orc.one_hot(labels, K) @ phi, 1e-10, 'permuted diagonal feature table')


@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
def test_sweep_empty_and_single_point(precision):
    """edge cases of mimo_sweep / mimo_stats_hard: no points (statistics stay zero), one point, one component."""
    E = eng()
    rng = np.random.default_rng(1)
    for K, d in ((3, 2), (1, 5), (40, 96)):
        mus = rng.standard_normal((K, d))
        lmbdas = np.stack([spd(rng, d) for _ in range(K)])
        logw = np.log(rng.dirichlet(np.ones(K))) if K > 1 else np.zeros(1)
        ops = E.QuadOperands(K, d, d, precision)
        E.set_log_weights(ops, logw)
        E.operands_gauss(ops, E.to_dev(mus), E.to_dev(lmbdas)).check()
        feats = E.quad_features(d)
        for N in (0, 1, 2):
            x = rng.standard_normal((N, d)) + mus[0]
            Z = E.to_dev(x.reshape(N, d), E.tdtype(precision))
            for hard in (False, True):
                buf = E.SweepBuffers(max(N, 1), K, feats.F, precision, hard)
                buf.N = N
                E.sweep(Z, ops, feats, buf, uniforms=E.to_dev(rng.random(max(N, 1))) if hard else None)
                st = buf.stat.cpu().numpy()
                if N == 0:
                    assert np.all(st == 0.0) and buf.lse_sum.item() == 0.0
                    continue
                xr = Z.double().cpu().numpy()
                ll = orc.gauss_full_loglik(xr, mus, lmbdas) + logw[:, None]
                resp, lse = orc.responsibilities(ll)
                assert abs(buf.lse_sum.item() - lse.sum()) <= RTOL[precision] * max(1.0, abs(lse.sum()))
                S = unpack_quad(st, d)
                assert abs(S[:, d, d].sum() - N) <= 1e-6 * N          # responsibilities / labels sum to the points
                if not hard:
                    close(S[:, d, :d], orc.gauss_full_wstats(xr, resp)[0], RTOL[precision], 'tiny sum r x')
    # hard statistics of nothing
    z0 = E.to_dev(np.zeros((0, 3)), E.tdtype(precision))
    s0 = E.stats_hard(z0, E.to_dev(np.zeros(0, dtype=np.int32), torch.int32), 4, E.quad_features(3), precision)
    assert np.all(s0.cpu().numpy() == 0.0)
