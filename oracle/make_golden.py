"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (imported from /root/reference) on small seeded problems.

TEST INFRASTRUCTURE.  Run in the build container only (the reference tree does
not travel to the GPU box):

    python oracle/make_golden.py

Each fixture is an .npz of float64/int32 arrays: inputs, every intermediate the
sweep produces (statistics, posterior parameters, sampled parameters,
log-probabilities, uniforms, labels, responsibilities, lower bounds) and the
raw random variates in the order the reference consumed them from the global
legacy numpy.random stream, so the phases can be replayed one at a time by the
oracle (tests/test_oracle_golden.py) and by the CUDA path (tests/test_gpu_*.py).
While generating, every recorded sampled parameter is re-derived from the
recorded variates through the oracle and asserted equal -- i.e. the generator
also pins the oracle's sampling restatement.
"""
import os
import sys

import numpy as np
import numpy.random as npr

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

import mimo.distributions as D          # noqa: E402
import mimo.mixtures as M               # noqa: E402
from oracle import mimo_oracle as orc   # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def spd(rng, d):
    a = rng.standard_normal((d, d + 2))
    return (a @ a.T) / d + 0.1 * np.eye(d)


def blobs(rng, N, d, K, spread=4.0, full=True):
    centres = spread * rng.standard_normal((K, d))
    z = rng.integers(0, K, size=N)
    x = np.empty((N, d))
    for k in range(K):
        idx = np.where(z == k)[0]
        if full:
            L = np.linalg.cholesky(spd(rng, d))
            x[idx] = centres[k] + rng.standard_normal((idx.size, d)) @ L.T
        else:
            x[idx] = centres[k] + rng.standard_normal((idx.size, d)) * (0.5 + rng.random(d))
    return x


def close(a, b, tol=1e-9):
    np.testing.assert_allclose(np.asarray(a, float), np.asarray(b, float), rtol=tol,
                               atol=tol * max(1.0, float(np.max(np.abs(b)))))


def draw_nw_variates(nus, d):
    """Reference order per component: normal(d(d-1)/2), d x chisquare(nu - i),
    normal(d)   (wishart.py:72-80, gaussian.py:311-313)."""
    K = len(nus)
    nt = d * (d - 1) // 2
    normals, chisq, z = np.zeros((K, nt)), np.zeros((K, d)), np.zeros((K, d))
    for k in range(K):
        normals[k] = npr.normal(size=nt)
        chisq[k] = [npr.chisquare(nus[k] - i, size=1)[0] for i in range(d)]
        z[k] = npr.normal(size=d)
    return normals, chisq, z


def draw_mnw_variates(nus, o, c):
    K = len(nus)
    nt = o * (o - 1) // 2
    normals, chisq, z = np.zeros((K, nt)), np.zeros((K, o)), np.zeros((K, o * c))
    for k in range(K):
        normals[k] = npr.normal(size=nt)
        chisq[k] = [npr.chisquare(nus[k] - i, size=1)[0] for i in range(o)]
        z[k] = npr.normal(size=o * c)
    return normals, chisq, z


def gating_kind(g):
    return 'dirichlet' if isinstance(g, D.CategoricalWithDirichlet) else 'stick'


def gating_prior_arrays(g):
    if gating_kind(g) == 'dirichlet':
        return dict(gate_alphas0=g.prior.alphas)
    return dict(gate_gammas0=g.prior.gammas, gate_deltas0=g.prior.deltas)


def resample_gating_recorded(g, labels, rec, t):
    state = npr.get_state()
    g.resample(labels)
    npr.set_state(state)
    if gating_kind(g) == 'dirichlet':
        gam = npr.standard_gamma(g.posterior.alphas)   # consumes the same stream as npr.dirichlet
        close(orc.dirichlet_probs_from_gammas(gam), g.likelihood.probs)
        rec[f'gate_gamma_{t}'] = gam
        rec[f'gate_alphas_{t}'] = g.posterior.alphas
    else:
        v = npr.beta(g.posterior.gammas[:-1], g.posterior.deltas[:-1])
        close(orc.stick_probs_from_betas(v), g.likelihood.probs)
        rec[f'gate_beta_{t}'] = v
        rec[f'gate_gammas_{t}'] = g.posterior.gammas
        rec[f'gate_deltas_{t}'] = g.posterior.deltas
    rec[f'probs_{t}'] = g.likelihood.probs


def gmm_gibbs_case(name, x, components, gating, sweeps, seed, diag=False):
    """mixtures/gmm.py:207-225 replayed phase by phase."""
    model = M.BayesianMixtureOfGaussians(gating=gating, components=components)
    K, d = model.size, model.dim
    rec = dict(obs=x, K=K, d=d, seed=seed, sweeps=sweeps)
    rec.update(gating_prior_arrays(gating))
    names = ('mus0', 'kappas0', 'alphas0', 'betas0') if diag else ('mus0', 'kappas0', 'psis0', 'nus0')
    for n, p in zip(names, components.prior.params):
        rec[n] = p
    npr.seed(seed)
    labels = npr.choice(K, size=len(x))
    rec['labels_init'] = labels.astype(np.int32)
    for t in range(sweeps):
        # components
        state = npr.get_state()
        model.resample_components(x, labels)
        npr.set_state(state)
        post = components.posterior.params
        if diag:
            # per-dist reference update (composite.py:332-337); the stacked
            # alphas/betas setters are broken so posterior.params
            # keeps the PRIOR alphas/betas: record both.
            w = orc.one_hot(labels, K)
            nat = orc.add_stats(orc.ng_std_to_nat(*components.prior.params), orc.gauss_diag_wstats(x, w))
            fixed = orc.ng_nat_to_std(nat)
            for n, p in zip(('mus', 'kappas', 'alphas', 'betas'), fixed):
                rec[f'post_{n}_{t}'] = p
            rec[f'bug_alphas_{t}'], rec[f'bug_betas_{t}'] = post[2], post[3]
            close(post[0], fixed[0]); close(post[1], fixed[1])
            g = npr.gamma(post[2], 1.0 / post[3])   # one call per component in the reference,
            # but element order is identical only if drawn per component: redo faithfully
            npr.set_state(state)
            gam, z = np.zeros((K, d)), np.zeros((K, d))
            for k in range(K):
                gam[k] = npr.gamma(post[2][k], 1.0 / post[3][k])
                z[k] = npr.normal(size=d)
            mu_s, l_s = orc.ng_rvs_from_variates(post[0], post[1], post[2], post[3], gam, z)
            close(mu_s, components.likelihood.mus); close(l_s, components.likelihood.lmbdas_diags)
            rec[f'var_gamma_{t}'], rec[f'var_z_{t}'] = gam, z
            rec[f'lik_mus_{t}'] = components.likelihood.mus
            rec[f'lik_lmbdas_diags_{t}'] = components.likelihood.lmbdas_diags
        else:
            for n, p in zip(('mus', 'kappas', 'psis', 'nus'), post):
                rec[f'post_{n}_{t}'] = p
            normals, chisq, z = draw_nw_variates(post[3], d)
            for k in range(K):
                mu_s, l_s = orc.nw_rvs_from_variates(post[0][k], post[1][k], post[2][k], post[3][k],
                                                     normals[k], chisq[k], z[k])
                close(mu_s, components.likelihood.mus[k]); close(l_s, components.likelihood.lmbdas[k])
            rec[f'var_normals_{t}'], rec[f'var_chisq_{t}'], rec[f'var_z_{t}'] = normals, chisq, z
            rec[f'lik_mus_{t}'] = components.likelihood.mus
            rec[f'lik_lmbdas_{t}'] = components.likelihood.lmbdas
        # gating
        resample_gating_recorded(gating, labels, rec, t)
        # labels
        state = npr.get_state()
        log_prob, labels = model.resample_labels(x)
        npr.set_state(state)
        u = npr.random(size=(1, len(x)))
        assert np.array_equal(orc.sample_discrete_from_log(log_prob, u), labels)
        rec[f'log_prob_{t}'] = log_prob
        rec[f'u_{t}'] = u[0]
        rec[f'labels_{t}'] = labels.astype(np.int32)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'ok')


def gmm_vi_case(name, x, components, gating, iters, seed, diag=False):
    """mixtures/gmm.py:261-287: VI trajectory with the randomised start."""
    model = M.BayesianMixtureOfGaussians(gating=gating, components=components)
    K, d = model.size, model.dim
    rec = dict(obs=x, K=K, d=d, seed=seed, iters=iters)
    rec.update(gating_prior_arrays(gating))
    names = ('mus0', 'kappas0', 'alphas0', 'betas0') if diag else ('mus0', 'kappas0', 'psis0', 'nus0')
    for n, p in zip(names, components.prior.params):
        rec[n] = p
    npr.seed(seed)
    resp = npr.rand(K, len(x))
    resp /= np.sum(resp, axis=0)
    rec['resp_init'] = resp
    vlbs = []
    for t in range(iters):
        model.meanfield_update_parameters(x, resp)
        resp = model.expected_responsibilities(x)
        vlbs.append(model.variational_lowerbound(x, resp))
        if not diag:
            for n, p in zip(('mus', 'kappas', 'psis', 'nus'), components.posterior.params):
                rec[f'post_{n}_{t}'] = p
        else:
            for n, p in zip(('mus', 'kappas'), components.posterior.params[:2]):
                rec[f'post_{n}_{t}'] = p
        if gating_kind(gating) == 'dirichlet':
            rec[f'gate_alphas_{t}'] = gating.posterior.alphas
        else:
            rec[f'gate_gammas_{t}'] = gating.posterior.gammas
            rec[f'gate_deltas_{t}'] = gating.posterior.deltas
        rec[f'ell_{t}'] = model.expected_log_complete_likelihood(x)
        rec[f'resp_{t}'] = resp
    rec['vlb'] = np.array(vlbs)
    rec['vlb_components'] = np.sum(components.variational_lowerbound())
    rec['vlb_gating'] = gating.variational_lowerbound()
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'ok', 'vlb monotone:', bool(np.all(np.diff(vlbs) >= -1e-8)))


def gmm_svi_case(name, x, components, gating, iters, batch_size, step_size, seed):
    """mixtures/gmm.py:300-336 (stochastic variational inference): one minibatch per iteration drawn with Python's
    `random.sample` (utils/data.py:9-12), natural-parameter blending in distributions/bayesian.py:85-91, 232-238, and the
    full-data lower bound after every iteration.  Seeds: random.seed (minibatches), numpy.random.seed (the randomised
    first responsibilities and the reference's trailing posterior.rvs() draws)."""
    import random
    model = M.BayesianMixtureOfGaussians(gating=gating, components=components)
    K, d = model.size, model.dim
    rec = dict(obs=x, K=K, d=d, seed=seed, iters=iters, batch_size=batch_size, step_size=step_size)
    rec.update(gating_prior_arrays(gating))
    for n, p in zip(('mus0', 'kappas0', 'psis0', 'nus0'), components.prior.params):
        rec[n] = p
    for n, p in zip(('pmus0', 'pkappas0', 'ppsis0', 'pnus0'), components.posterior.params):
        rec[n] = p
    # replay of the draws the run below makes, for the oracle test: batch indices and the first responsibilities
    random.seed(seed)
    npr.seed(seed)
    rec['batches'] = np.array([random.sample(range(len(x)), batch_size) for _ in range(iters)])
    r0 = npr.rand(K, batch_size)
    rec['resp0'] = r0 / np.sum(r0, axis=0)
    random.seed(seed)
    npr.seed(seed)
    vlb = model.meanfield_stochastic_descent(x, randomize=True, maxiter=iters, step_size=step_size, batch_size=batch_size,
                                             progress_bar=False)
    rec['vlb'] = np.array(vlb)
    for n, p in zip(('mus', 'kappas', 'psis', 'nus'), components.posterior.params):
        rec[f'post_{n}'] = p
    if gating_kind(gating) == 'dirichlet':
        rec['gate_alphas'] = gating.posterior.alphas
    else:
        rec['gate_gammas'], rec['gate_deltas'] = gating.posterior.gammas, gating.posterior.deltas
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'ok', 'vlb', vlb[0], '->', vlb[-1])


def gmm_em_case(name, x, K, seed):
    """mixtures/gmm.py:77-103 (EM) on full-covariance components."""
    d = x.shape[1]
    rng = np.random.default_rng(seed)
    comp = D.StackedGaussiansWithPrecision(K, d, mus=rng.standard_normal((K, d)),
                                           lmbdas=np.stack(K * [np.eye(d)]))
    model = M.MixtureOfGaussians(gating=D.Categorical(K), components=comp)
    npr.seed(seed)
    resp = npr.rand(K, len(x))
    resp /= np.sum(resp, axis=0)
    npr.seed(seed)
    ll = model.max_likelihood(x, maxiter=6, progress_bar=False)
    rec = dict(obs=x, K=K, d=d, resp_init=resp, ll=np.array(ll), mus=comp.mus, lmbdas=comp.lmbdas,
               probs=model.gating.probs, resp_final=model.responsibilities(x))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'ok', 'll monotone:', bool(np.all(np.diff(ll) >= -1e-8)))


def ilr_models(K, din, o, tied, rng):
    c = din + 1
    basis_prior = D.StackedNormalWisharts(K, din, mus=np.zeros((K, din)), kappas=1e-2 * np.ones(K),
                                          psis=np.stack(K * [1e2 * np.eye(din)]),
                                          nus=(din + 1) * np.ones(K) + 1e-16)
    basis = D.StackedGaussiansWithNormalWisharts(K, din, prior=basis_prior)
    pcls = D.TiedMatrixNormalWisharts if tied else D.StackedMatrixNormalWisharts
    mcls = D.TiedLinearGaussiansWithMatrixNormalWisharts if tied else D.StackedLinearGaussiansWithMatrixNormalWisharts
    models_prior = pcls(K, c, o, Ms=np.zeros((K, o, c)), Ks=np.stack(K * [1e-2 * np.eye(c)]),
                        psis=np.stack(K * [1e1 * np.eye(o)]), nus=(o + 1) * np.ones(K) + 1e-16)
    models = mcls(K, c, o, models_prior, affine=True)
    gating = D.CategoricalWithStickBreaking(K, D.TruncatedStickBreaking(K, np.ones(K), 5.0 * np.ones(K)))
    return basis, models, gating


def ilr_case(name, x, y, K, tied, seed, sweeps=2, iters=3):
    """mixtures/ilr.py:134-159 (Gibbs) then :196-228 (VI, randomize=False) --
    the order examples/ilr/evaluate_sine.py:124-147 uses."""
    din, o = x.shape[1], y.shape[1]
    c = din + 1
    rng = np.random.default_rng(seed)
    npr.seed(seed)
    basis, models, gating = ilr_models(K, din, o, tied, rng)
    ilr = M.BayesianMixtureOfLinearGaussians(K, din, o, gating=gating, basis=basis, models=models)
    rec = dict(x=x, y=y, K=K, din=din, o=o, tied=int(tied), seed=seed, sweeps=sweeps, iters=iters)
    for n, p in zip(('b_mus0', 'b_kappas0', 'b_psis0', 'b_nus0'), basis.prior.params):
        rec[n] = p
    for n, p in zip(('m_Ms0', 'm_Ks0', 'm_psis0', 'm_nus0'), models.prior.params):
        rec[n] = p
    rec.update(gating_prior_arrays(gating))
    labels = npr.choice(K, size=len(x))
    rec['labels_init'] = labels.astype(np.int32)
    for t in range(sweeps):
        state = npr.get_state()
        ilr.resample_basis(x, labels)
        npr.set_state(state)
        bp = basis.posterior.params
        normals, chisq, z = draw_nw_variates(bp[3], din)
        for n, p in zip(('mus', 'kappas', 'psis', 'nus'), bp):
            rec[f'b_post_{n}_{t}'] = p
        rec[f'b_var_normals_{t}'], rec[f'b_var_chisq_{t}'], rec[f'b_var_z_{t}'] = normals, chisq, z
        rec[f'b_lik_mus_{t}'], rec[f'b_lik_lmbdas_{t}'] = basis.likelihood.mus, basis.likelihood.lmbdas
        state = npr.get_state()
        ilr.resample_models(x, y, labels)
        npr.set_state(state)
        mp = models.posterior.params
        normals, chisq, z = draw_mnw_variates(mp[3], o, c)
        for k in range(K):
            A_s, l_s = orc.mnw_rvs_from_variates(mp[0][k], mp[1][k], mp[2][k], mp[3][k], normals[k], chisq[k], z[k])
            close(A_s, models.likelihood.As[k], 1e-8); close(l_s, models.likelihood.lmbdas[k], 1e-8)
        for n, p in zip(('Ms', 'Ks', 'psis', 'nus'), mp):
            rec[f'm_post_{n}_{t}'] = p
        rec[f'm_var_normals_{t}'], rec[f'm_var_chisq_{t}'], rec[f'm_var_z_{t}'] = normals, chisq, z
        rec[f'm_lik_As_{t}'], rec[f'm_lik_lmbdas_{t}'] = models.likelihood.As, models.likelihood.lmbdas
        resample_gating_recorded(gating, labels, rec, t)
        state = npr.get_state()
        log_prob, labels = ilr.resample_labels(x, y)
        npr.set_state(state)
        u = npr.random(size=(1, len(x)))
        assert np.array_equal(orc.sample_discrete_from_log(log_prob, u), labels)
        rec[f'log_prob_{t}'], rec[f'u_{t}'], rec[f'labels_{t}'] = log_prob, u[0], labels.astype(np.int32)
    # mean-field continuation from the Gibbs state (randomize=False)
    resp = ilr.expected_responsibilities(x, y)
    rec['vi_resp_init'] = resp
    vlbs = []
    for t in range(iters):
        ilr.meanfield_update_parameters(x, y, resp)
        resp = ilr.expected_responsibilities(x, y)
        vlbs.append(ilr.variational_lowerbound(x, y, resp))
        for n, p in zip(('mus', 'kappas', 'psis', 'nus'), basis.posterior.params):
            rec[f'vi_b_post_{n}_{t}'] = p
        for n, p in zip(('Ms', 'Ks', 'psis', 'nus'), models.posterior.params):
            rec[f'vi_m_post_{n}_{t}'] = p
        rec[f'vi_gate_gammas_{t}'], rec[f'vi_gate_deltas_{t}'] = gating.posterior.gammas, gating.posterior.deltas
        rec[f'vi_ell_{t}'] = ilr.expected_log_complete_likelihood(x, y)
        rec[f'vi_resp_{t}'] = resp
    rec['vlb'] = np.array(vlbs)
    mu, var, std = ilr.meanfield_prediction(x[:32], prediction='average')
    rec['pred_mu'], rec['pred_var'] = mu, var
    # the branches of the prediction path (ilr.py:325-430) the reference can run: Gaussian weights and moments, mixture /
    # mode.  Its Student-t branches (stats.py:79 divides (K, N) by (K,); ilr.py:358 contracts a 4-D array as 'ndl') and its
    # NLPD branch (bayesian.py:964-966 feeds per-point precisions to stacked_mvn_logpdf) raise broadcasting errors for
    # every K != N, so no fixture can exist for them: tests check those against oracle/mimo_oracle.py's restatement with
    # the evident broadcasting.
    rec['pred_weights_gaussian'] = ilr.meanfield_predictive_weights(x[:48], 'gaussian')
    for pred in ('average', 'mode'):
        mu, cov, std = ilr.meanfield_prediction(x[:48].copy(), prediction=pred, dist='gaussian', variance='full')
        rec[f'pred_gaussian_{pred}_mu'], rec[f'pred_gaussian_{pred}_cov'] = mu, cov
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'ok', 'vlb monotone:', bool(np.all(np.diff(vlbs) >= -1e-8)))


def pointwise_case(name, K, d, N, seed):
    """Per-point kernels at a cfg5-like shape (d=128): Gibbs log-lik and VI
    expected log-lik of the full-covariance family + weighted statistics."""
    rng = np.random.default_rng(seed)
    x = blobs(rng, N, d, K, spread=2.0)
    mus = 2.0 * rng.standard_normal((K, d))
    lmbdas = np.stack([spd(rng, d) for _ in range(K)])
    kappas = rng.random(K) + 0.5
    psis = np.stack([spd(rng, d) / (d + 4.0) for _ in range(K)])
    nus = d + 2.0 + 5 * rng.random(K)
    lik = D.StackedGaussiansWithPrecision(K, d, mus=mus, lmbdas=lmbdas)
    prior = D.StackedNormalWisharts(K, d, mus, kappas, psis, nus)
    wrap = D.StackedGaussiansWithNormalWisharts(K, d, prior=prior, likelihood=lik)
    w = rng.random((K, N))
    w /= w.sum(0)
    st = lik.weighted_statistics(x, w)
    rec = dict(obs=x, mus=mus, lmbdas=lmbdas, kappas=kappas, psis=psis, nus=nus, weights=w,
               log_lik=lik.log_likelihood(x.copy()), exp_log_lik=wrap.expected_log_likelihood(x),
               st_x=st[0], st_n=st[1], st_xx=st[2])
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'ok')


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(1337)

    # cfg1: toy Bayesian GMM as shipped (examples/gmm/toy/gibbs_toy.py:41-59, vi_toy.py:41-59)
    K, d = 4, 2
    toy = blobs(rng, 500, d, K, spread=4.0)

    def toy_model(gate='dirichlet', alpha=1.0):
        prior = D.StackedNormalWisharts(K, d, mus=np.zeros((K, d)), kappas=1e-2 * np.ones(K),
                                        psis=np.stack(K * [np.eye(d)]), nus=3.0 * np.ones(K) + 1e-8)
        npr.seed(1)
        comp = D.StackedGaussiansWithNormalWisharts(K, d, prior=prior)
        if gate == 'dirichlet':
            g = D.CategoricalWithDirichlet(K, D.Dirichlet(K, alpha * np.ones(K)))
        else:
            g = D.CategoricalWithStickBreaking(K, D.TruncatedStickBreaking(K, np.ones(K), 5.0 * np.ones(K)))
        return comp, g

    gmm_gibbs_case('gmm_toy_gibbs', toy, *toy_model(), sweeps=3, seed=1337)
    gmm_vi_case('gmm_toy_vi', toy, *toy_model(), iters=6, seed=1337)
    gmm_vi_case('gmm_toy_vi_stick', toy, *toy_model('stick'), iters=4, seed=7)
    gmm_svi_case('gmm_toy_svi', toy, *toy_model(), iters=8, batch_size=96, step_size=0.2, seed=5)
    gmm_svi_case('gmm_toy_svi_stick', toy, *toy_model('stick'), iters=6, batch_size=64, step_size=0.1, seed=9)
    gmm_em_case('gmm_toy_em', toy, K, seed=3)

    # cfg1 at its shipped size: examples/gmm/sine/{vi,gibbs}_gmm.py (N = 2500 points on a noisy sine, K = 25, alpha = 2)
    rs = np.random.default_rng(4242)
    xs = np.arange(2500) * (14. * np.pi / 2500) - 6.
    sine = np.stack([xs + rs.normal(0, 0.1, 2500), 3. * (np.sin(xs) + rs.normal(0, .1, 2500))], axis=1)

    def sine_model():
        Ks = 25
        prior = D.StackedNormalWisharts(Ks, 2, mus=np.zeros((Ks, 2)), kappas=0.01 * np.ones(Ks),
                                        psis=np.stack(Ks * [np.eye(2)]), nus=3. * np.ones(Ks) + 1e-8)
        npr.seed(1)
        return (D.StackedGaussiansWithNormalWisharts(Ks, 2, prior=prior),
                D.CategoricalWithDirichlet(Ks, D.Dirichlet(Ks, 2. * np.ones(Ks))))

    gmm_vi_case('gmm_sine_vi', sine, *sine_model(), iters=2, seed=2500)
    gmm_gibbs_case('gmm_sine_gibbs', sine, *sine_model(), sweeps=3, seed=2500)

    # cfg4-shaped: full covariance d=16, stick-breaking DP-GMM
    K4, d4 = 8, 16
    x4 = blobs(rng, 600, d4, K4, spread=3.0)
    prior = D.StackedNormalWisharts(K4, d4, mus=np.zeros((K4, d4)), kappas=1e-2 * np.ones(K4),
                                    psis=np.stack(K4 * [np.eye(d4)]), nus=(d4 + 1) * np.ones(K4) + 1e-8)
    npr.seed(2)
    comp = D.StackedGaussiansWithNormalWisharts(K4, d4, prior=prior)
    g = D.CategoricalWithStickBreaking(K4, D.TruncatedStickBreaking(K4, np.ones(K4), 5.0 * np.ones(K4)))
    gmm_vi_case('gmm_d16_vi_stick', x4, comp, g, iters=4, seed=11)
    npr.seed(2)
    comp = D.StackedGaussiansWithNormalWisharts(K4, d4, prior=prior)
    g = D.CategoricalWithStickBreaking(K4, D.TruncatedStickBreaking(K4, np.ones(K4), 5.0 * np.ones(K4)))
    gmm_gibbs_case('gmm_d16_gibbs_stick', x4, comp, g, sweeps=2, seed=12)

    # cfg3-shaped: diagonal mixture (examples/dgmm/gibbs_dgmm.py:42-58)
    K3, d3 = 6, 8
    x3 = blobs(rng, 500, d3, K3, spread=4.0, full=False)

    def diag_model():
        prior = D.StackedNormalGammas(K3, d3, mus=np.zeros((K3, d3)), kappas=1e-2 * np.ones((K3, d3)),
                                      alphas=(3.0 + 1e-8) / 2 * np.ones((K3, d3)), betas=0.5 * np.ones((K3, d3)))
        npr.seed(3)
        comp = D.StackedGaussiansWithNormalGammas(K3, d3, prior=prior)
        return comp, D.CategoricalWithDirichlet(K3, D.Dirichlet(K3, np.ones(K3)))

    gmm_gibbs_case('dgmm_gibbs', x3, *diag_model(), sweeps=2, seed=21, diag=True)
    gmm_vi_case('dgmm_vi_bugcompat', x3, *diag_model(), iters=3, seed=22, diag=True)

    # cfg2-shaped: ILR, stick-breaking, tied and stacked MNW (examples/ilr/evaluate_sine.py:88-127)
    N2, din, o = 400, 2, 1
    xi = rng.standard_normal((N2, din)) * 2.0
    yi = np.sin(xi @ np.array([[1.0], [0.5]])) + 0.3 * rng.standard_normal((N2, o))
    xi = (xi - xi.mean(0)) / xi.std(0)
    yi = (yi - yi.mean(0)) / yi.std(0)
    ilr_case('ilr_tied', xi, yi, K=6, tied=True, seed=31)
    ilr_case('ilr_stacked', xi, yi, K=5, tied=False, seed=32)
    y2 = np.hstack((yi, np.cos(xi[:, :1]) + 0.2 * rng.standard_normal((N2, 1))))
    ilr_case('ilr_stacked_o2', xi, y2, K=4, tied=False, seed=33, sweeps=1, iters=2)

    # cfg5-shaped per-point kernels at d=128 (N small: the reference's VI E-step is 8*K*N*d^2 bytes)
    pointwise_case('pointwise_d128', K=5, d=128, N=48, seed=41)
    pointwise_case('pointwise_d16', K=9, d=16, N=200, seed=42)


# ---- hierarchical mixtures (mixtures/hgmm.py, distributions/bayesian.py:595-793) -------------------------
def hgmm_model(K, d, stick, ctor_seed):
    """examples/hgmm/vi_component.py:42-60 with a parameterised size; returns the model and what its constructor drew."""
    npr.seed(ctor_seed)
    if stick:
        gating = D.CategoricalWithStickBreaking(dim=K, prior=D.TruncatedStickBreaking(dim=K, gammas=np.ones(K), deltas=2. * np.ones(K)))
    else:
        gating = D.CategoricalWithDirichlet(dim=K, prior=D.Dirichlet(dim=K, alphas=np.ones(K)))
    hp = D.NormalWishart(dim=d, mu=np.zeros(d), kappa=1e-2, psi=np.eye(d), nu=d + 1 + 1e-8)
    pr = D.TiedGaussiansWithScaledPrecision(size=K, dim=d, kappas=1e-2 * (1. + np.arange(K)))
    comp = D.TiedGaussiansWithHierarchicalNormalWisharts(size=K, dim=d, hyper_prior=hp, prior=pr)
    model = M.BayesianMixtureOfGaussiansWithHierarchicalPrior(size=K, dim=d, gating=gating, components=comp)
    rec = dict(K=K, d=d, stick=int(stick), ctor_seed=ctor_seed, kappas0=pr.kappas.copy(),
               hyper_mu0=hp.gaussian.mu, hyper_kappa0=hp.kappa, hyper_psi0=hp.wishart.psi, hyper_nu0=hp.wishart.nu,
               init_prior_mus=comp.prior.mus.copy(), init_prior_lmbdas=comp.prior.lmbdas.copy(),
               init_lik_mus=comp.likelihood.mus.copy(), init_probs=gating.likelihood.probs.copy())
    rec.update(gating_prior_arrays(gating))
    return model, rec


def hier_state(model, rec, tag):
    comp = model.components
    rec[f'post_mus_{tag}'], rec[f'post_kappas_{tag}'] = comp.posterior.mus.copy(), comp.posterior.kappas.copy()
    for n, p in zip(('rho', 'kappa', 'psi', 'nu'), comp.hyper_posterior.params):
        rec[f'hyper_{n}_{tag}'] = np.array(p)
    rec[f'lik_mus_{tag}'], rec[f'lik_lmbdas_{tag}'] = comp.likelihood.mus.copy(), comp.likelihood.lmbdas.copy()
    g = model.gating
    if gating_kind(g) == 'dirichlet':
        rec[f'gate_alphas_{tag}'] = g.posterior.alphas.copy()
    else:
        rec[f'gate_gammas_{tag}'], rec[f'gate_deltas_{tag}'] = g.posterior.gammas.copy(), g.posterior.deltas.copy()


def hgmm_vi_case(name, x, K, stick, iters, subiters, ctor_seed, seed):
    """mixtures/hgmm.py:186-215; the oracle's restatement is asserted against the run while recording."""
    model, rec = hgmm_model(K, x.shape[1], stick, ctor_seed)
    comp = model.components
    lm0 = comp.posterior.lmbdas.copy()
    rec.update(obs=x, iters=iters, subiters=subiters, seed=seed)
    npr.seed(seed)
    vlb = model.meanfield_coordinate_descent(x, randomize=True, maxiter=iters, maxsubiter=subiters, tol=0., progress_bar=False)
    rec['vlb'] = np.array(vlb)
    hier_state(model, rec, 'end')
    rec['ell_end'] = model.expected_log_complete_likelihood(x)
    rec['resp_end'] = model.expected_responsibilities(x)
    rec['comp_vlb_end'] = comp.variational_lowerbound()
    npr.seed(seed)
    r0 = npr.rand(K, len(x))
    r0 /= r0.sum(0)
    gp = ('stick', model.gating.prior.gammas, model.gating.prior.deltas) if stick else ('dirichlet', model.gating.prior.alphas)
    out = orc.hgmm_meanfield(x, r0, gp, tuple(comp.hyper_prior.params), comp.prior.kappas, lm0, iters, subiters)
    close(out['vlb'], vlb)
    close(out['mus'], comp.posterior.mus)
    close(out['ell'], rec['ell_end'])
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'vlb', vlb[0], '->', vlb[-1])


def hgmm_gibbs_case(name, x, K, stick, sweeps, subiters, ctor_seed, seed):
    """mixtures/hgmm.py:137-161."""
    model, rec = hgmm_model(K, x.shape[1], stick, ctor_seed)
    rec.update(obs=x, sweeps=sweeps, subiters=subiters, seed=seed)
    npr.seed(seed)
    model.resample(x, maxiter=sweeps, maxsubiter=subiters, progress_bar=False)
    hier_state(model, rec, 'end')
    rec['probs_end'] = model.gating.likelihood.probs.copy()
    npr.seed(seed + 1)
    lp, lab = model.resample_labels(x)
    rec['log_prob_end'], rec['labels_next'] = lp, lab.astype(np.int32)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'label counts', np.bincount(lab, minlength=K))


def hgmm_svi_case(name, x, K, stick, iters, subiters, step_size, ctor_seed, seed):
    """mixtures/hgmm.py:228-262 (full-batch natural-gradient steps)."""
    model, rec = hgmm_model(K, x.shape[1], stick, ctor_seed)
    rec.update(obs=x, iters=iters, subiters=subiters, seed=seed, step_size=step_size)
    npr.seed(seed)
    model.meanfield_stochastic_descent(x, randomize=True, maxiter=iters, maxsubiter=subiters, step_size=step_size, progress_bar=False)
    hier_state(model, rec, 'end')
    rec['resp_end'] = model.expected_responsibilities(x)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'posterior kappas', model.components.posterior.kappas)


def hmom_case(name, x, M_, K, iters, subiters, subsubiters, ctor_seed, seed):
    """mixtures/hgmm.py:298-431: a mixture of M_ hierarchical mixtures of K Gaussians, mean field."""
    d = x.shape[1]
    npr.seed(ctor_seed)
    gating = D.CategoricalWithDirichlet(dim=M_, prior=D.Dirichlet(dim=M_, alphas=np.ones(M_)))
    subs = []
    for m in range(M_):
        sub, _ = hgmm_model(K, d, False, ctor_seed + 1 + m)
        subs.append(sub)
    model = M.BayesianMixtureOfMixtureOfGaussians(M_, K, d, gating=gating, components=subs)
    rec = dict(obs=x, M=M_, K=K, d=d, iters=iters, subiters=subiters, subsubiters=subsubiters, ctor_seed=ctor_seed, seed=seed)
    npr.seed(seed)
    model.meanfield_coordinate_descent(x, randomize=True, maxiter=iters, maxsubiter=subiters, maxsubsubiter=subsubiters, progress_bar=False)
    rec['resp_end'] = model.expected_responsibilities(x)
    rec['gate_alphas_end'] = gating.posterior.alphas.copy()
    for m, sub in enumerate(subs):
        rec[f'sub{m}_post_mus'] = sub.components.posterior.mus.copy()
        rec[f'sub{m}_hyper_psi'] = np.array(sub.components.hyper_posterior.params[2])
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'cluster masses', rec['resp_end'].sum(1))


def hilr_model(K, din, o, stick, ctor_seed):
    """examples/hilr/vi_component.py:45-79 with parameterised sizes; returns the model and what its constructor drew."""
    npr.seed(ctor_seed)
    if stick:
        gating = D.CategoricalWithStickBreaking(dim=K, prior=D.TruncatedStickBreaking(dim=K, gammas=np.ones(K), deltas=2. * np.ones(K)))
    else:
        gating = D.CategoricalWithDirichlet(dim=K, prior=D.Dirichlet(dim=K, alphas=np.ones(K)))
    bh = D.NormalWishart(dim=din, mu=np.zeros(din), kappa=1e-2, psi=np.eye(din), nu=din + 1 + 1e-8)
    bp = D.TiedGaussiansWithScaledPrecision(size=K, dim=din, kappas=1e-2 * np.ones(K))
    basis = D.TiedGaussiansWithHierarchicalNormalWisharts(size=K, dim=din, hyper_prior=bh, prior=bp)
    sp = D.MatrixNormalWithPrecision(column_dim=din, row_dim=o, M=np.zeros((o, din)), K=1e-2 * np.eye(din))
    op = D.TiedGaussiansWithScaledPrecision(size=K, dim=o, mus=np.zeros((K, o)), kappas=1e-2 * (1. + np.arange(K)))
    pp = D.Wishart(dim=o, psi=np.eye(o), nu=o + 1 + 1e-8)
    models = D.TiedAffineLinearGaussiansWithMatrixNormalWisharts(size=K, column_dim=din, row_dim=o, slope_prior=sp,
                                                                 offset_prior=op, precision_prior=pp)
    model = M.BayesianMixtureOfLinearGaussiansWithTiedActivation(size=K, input_dim=din, output_dim=o, gating=gating,
                                                                 basis=basis, models=models)
    rec = dict(K=K, din=din, o=o, stick=int(stick), ctor_seed=ctor_seed,
               init_basis_lmbdas=basis.prior.lmbdas.copy(), init_basis_lik_mus=basis.likelihood.mus.copy(),
               init_As=models.likelihood.As.copy(), init_cs=models.likelihood.cs.copy(), init_lmbdas=models.likelihood.lmbdas.copy(),
               init_probs=gating.likelihood.probs.copy(), off_kappas0=op.kappas.copy())
    rec.update(gating_prior_arrays(gating))
    return model, rec


def hilr_state(model, rec):
    b, m = model.basis, model.models
    rec['basis_post_mus'], rec['basis_post_kappas'] = b.posterior.mus.copy(), b.posterior.kappas.copy()
    for n, p in zip(('rho', 'kappa', 'psi', 'nu'), b.hyper_posterior.params):
        rec[f'basis_hyper_{n}'] = np.array(p)
    rec['slope_M'], rec['slope_K'] = m.slope_posterior.M.copy(), m.slope_posterior.K.copy()
    rec['prec_psi'], rec['prec_nu'] = m.precision_posterior.psi.copy(), m.precision_posterior.nu
    rec['off_mus'], rec['off_kappas'] = m.offset_posterior.mus.copy(), m.offset_posterior.kappas.copy()
    rec['lik_As'], rec['lik_cs'], rec['lik_lmbdas'] = m.likelihood.As.copy(), m.likelihood.cs.copy(), m.likelihood.lmbdas.copy()
    rec['basis_lik_mus'], rec['basis_lik_lmbdas'] = b.likelihood.mus.copy(), b.likelihood.lmbdas.copy()
    rec['probs'] = model.gating.likelihood.probs.copy()


def hilr_vi_case(name, x, y, K, stick, iters, subiters, ctor_seed, seed):
    model, rec = hilr_model(K, x.shape[1], y.shape[1], stick, ctor_seed)
    rec.update(x=x, y=y, iters=iters, subiters=subiters, seed=seed)
    npr.seed(seed)
    model.meanfield_coordinate_descent(x, y, randomize=True, maxiter=iters, maxsubiter=subiters, progress_bar=False)
    hilr_state(model, rec)
    rec['ell_end'] = model.expected_log_complete_likelihood(x, y)
    rec['resp_end'] = model.expected_responsibilities(x, y)
    rec['models_vlb'] = model.models.variational_lowerbound()
    rec['vlb_end'] = model.variational_lowerbound(x, y, rec['resp_end'])
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'slope', model.models.slope_posterior.M.ravel(), 'vlb', rec['vlb_end'])


def hilr_gibbs_case(name, x, y, K, stick, sweeps, subiters, ctor_seed, seed):
    model, rec = hilr_model(K, x.shape[1], y.shape[1], stick, ctor_seed)
    rec.update(x=x, y=y, sweeps=sweeps, subiters=subiters, seed=seed)
    npr.seed(seed)
    model.resample(x, y, maxiter=sweeps, maxsubiter=subiters, progress_bar=False)
    hilr_state(model, rec)
    npr.seed(seed + 1)
    lp, lab = model.resample_labels(x, y)
    rec['log_prob_end'], rec['labels_next'] = lp, lab.astype(np.int32)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'label counts', np.bincount(lab, minlength=K))


def hmoilr_case(name, x, y, M_, K, iters, subiters, subsubiters, ctor_seed, seed):
    """mixtures/hilr.py:293-609: a mixture of M_ tied-activation mixtures of K experts: mean field, then the predictive path."""
    din, o = x.shape[1], y.shape[1]
    npr.seed(ctor_seed)
    gating = D.CategoricalWithDirichlet(dim=M_, prior=D.Dirichlet(dim=M_, alphas=np.ones(M_)))
    subs = [hilr_model(K, din, o, False, ctor_seed + 1 + m)[0] for m in range(M_)]
    model = M.BayesianMixtureOfMixtureOfLinearGaussians(M_, K, din, o, gating=gating, components=subs)
    rec = dict(x=x, y=y, M=M_, K=K, din=din, o=o, iters=iters, subiters=subiters, subsubiters=subsubiters, ctor_seed=ctor_seed, seed=seed,
               off_kappas0=1e-2 * (1. + np.arange(K)), gate_alphas0=np.ones(K), stick=0)
    npr.seed(seed)
    model.meanfield_coordinate_descent(x, y, randomize=True, maxiter=iters, maxsubiter=subiters, maxsubsubiter=subsubiters, progress_bar=False)
    rec['resp_end'] = model.expected_responsibilities(x, y)
    rec['gate_alphas_end'] = gating.posterior.alphas.copy()
    for m, sub in enumerate(subs):
        rec[f'sub{m}_slope_M'] = sub.models.slope_posterior.M.copy()
        rec[f'sub{m}_off_mus'] = sub.models.offset_posterior.mus.copy()
        rec[f'sub{m}_basis_mus'] = sub.basis.posterior.mus.copy()
    rec['weights'] = model.meanfield_predictive_weights(x)
    for pred in ('average', 'mode'):
        mean, var, std = model.meanfield_prediction(x, prediction=pred)
        rec[f'pred_{pred}_mean'], rec[f'pred_{pred}_var'] = mean, var
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'cluster masses', rec['resp_end'].sum(1), 'rmse', np.sqrt(np.mean((rec['pred_average_mean'] - y) ** 2)))


def affine_expert_gibbs_case(name, x, y, iters, ctor_seed, seed):
    """distributions/bayesian.py:1137-1219 (examples/lingauss/gibbs_affine.py): one affine expert, seeded Gibbs."""
    din, o = x.shape[1], y.shape[1]
    npr.seed(ctor_seed)
    sp = D.MatrixNormalWithPrecision(column_dim=din, row_dim=o, M=np.zeros((o, din)), K=1e-2 * np.eye(din))
    op = D.GaussianWithScaledPrecision(dim=o, kappa=1e-2, mu=np.zeros(o))
    pp = D.Wishart(dim=o, psi=np.eye(o), nu=o + 1 + 1e-8)
    w = D.AffineLinearGaussianWithMatrixNormalWishart(din, o, slope_prior=sp, offset_prior=op, precision_prior=pp)
    rec = dict(x=x, y=y, iters=iters, ctor_seed=ctor_seed, seed=seed, init_A=w.likelihood.A.copy(), init_c=w.likelihood.c.copy())
    npr.seed(seed)
    w.resample(x, y, nb_iter=iters)
    rec.update(A=w.likelihood.A, c=w.likelihood.c, lmbda=w.likelihood.lmbda, slope_M=w.slope_posterior.M, slope_K=w.slope_posterior.K,
               psi=w.precision_posterior.psi, nu=w.precision_posterior.nu, off_mu=w.offset_posterior.mu, off_kappa=w.offset_posterior.kappa)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'A', w.likelihood.A.ravel(), 'c', w.likelihood.c)


def hierarchical_ilr():
    rng = np.random.default_rng(78)
    n = 300
    x1 = np.linspace(-1.5, 1.5, n)[:, None]
    y1 = np.vstack([np.linspace(-.5, .5, n // 3)[:, None] for _ in range(3)]) + 0.05 * rng.standard_normal((n, 1))
    y1[:n // 3] += .5
    y1[-n // 3:] -= .5
    hilr_vi_case('hilr_vi', x1, y1, 3, False, iters=5, subiters=4, ctor_seed=1337, seed=3)
    hilr_gibbs_case('hilr_gibbs', x1, y1, 3, False, sweeps=3, subiters=3, ctor_seed=1337, seed=4)
    x2 = rng.uniform(-1.5, 1.5, (240, 2))
    y2 = x2 @ rng.standard_normal((2, 2)) + 0.5 * rng.integers(-1, 2, (240, 1)) + 0.05 * rng.standard_normal((240, 2))
    hilr_vi_case('hilr_d2_vi_stick', x2, y2, 4, True, iters=4, subiters=3, ctor_seed=9, seed=10)
    affine_expert_gibbs_case('affine_expert_gibbs', x2, y2 + np.array([1., -2.]), iters=5, ctor_seed=70, seed=8)
    hmoilr_case('hmoilr_vi', x1[::2], y1[::2], 2, 2, iters=2, subiters=2, subsubiters=2, ctor_seed=60, seed=7)


def ilr_svi_case(name, x, y, K, tied, iters, batch_size, step_size, seed):
    """mixtures/ilr.py:245-291 (stochastic variational inference of the linear-expert mixture): seeds as in gmm_svi_case."""
    import random
    din, o = x.shape[1], y.shape[1]
    rng = np.random.default_rng(seed)
    npr.seed(seed)
    basis, models, gating = ilr_models(K, din, o, tied, rng)
    ilr = M.BayesianMixtureOfLinearGaussians(K, din, o, gating=gating, basis=basis, models=models)
    rec = dict(x=x, y=y, K=K, din=din, o=o, tied=int(tied), seed=seed, iters=iters, batch_size=batch_size, step_size=step_size)
    for n, p in zip(('b_mus0', 'b_kappas0', 'b_psis0', 'b_nus0'), basis.prior.params):
        rec[n] = p
    for n, p in zip(('m_Ms0', 'm_Ks0', 'm_psis0', 'm_nus0'), models.prior.params):
        rec[n] = p
    rec.update(gating_prior_arrays(gating))
    random.seed(seed)
    npr.seed(seed)
    vlb = ilr.meanfield_stochastic_descent(x, y, randomize=True, maxiter=iters, step_size=step_size, batch_size=batch_size, progress_bar=False)
    rec['vlb'] = np.array(vlb)
    for n, p in zip(('mus', 'kappas', 'psis', 'nus'), basis.posterior.params):
        rec[f'b_post_{n}'] = p
    for n, p in zip(('Ms', 'Ks', 'psis', 'nus'), models.posterior.params):
        rec[f'm_post_{n}'] = p
    rec['gate_gammas'], rec['gate_deltas'] = gating.posterior.gammas, gating.posterior.deltas
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'vlb', vlb[0], '->', vlb[-1])


def ilr_svi():
    rng = np.random.default_rng(20170)
    N2, din, o = 400, 2, 1
    xi = rng.standard_normal((N2, din)) * 2.0
    yi = np.sin(xi @ np.array([[1.0], [0.5]])) + 0.3 * rng.standard_normal((N2, o))
    xi = (xi - xi.mean(0)) / xi.std(0)
    yi = (yi - yi.mean(0)) / yi.std(0)
    ilr_svi_case('ilr_svi_stacked', xi, yi, K=5, tied=False, iters=8, batch_size=96, step_size=0.1, seed=51)
    ilr_svi_case('ilr_svi_tied', xi, yi, K=4, tied=True, iters=6, batch_size=64, step_size=0.2, seed=52)


def hmom_em_case(name, x, M_, K, iters, subiters, seed):
    """mixtures/hgmm.py:16-89 (examples/hgmm/em_hgmm.py:66-93): EM of a mixture of mixtures of tied Gaussians."""
    d = x.shape[1]
    rng = np.random.default_rng(seed)
    rec = dict(obs=x, M=M_, K=K, d=d, iters=iters, subiters=subiters, seed=seed)
    comps = []
    for m in range(M_):
        prob = rng.random(K)
        prob /= prob.sum()
        mus = 3. * rng.standard_normal((K, d))
        rec[f'probs{m}'], rec[f'mus{m}'] = prob, mus
        comps.append(M.MixtureOfGaussians(gating=D.Categorical(dim=K, probs=prob),
                                          components=D.TiedGaussiansWithPrecision(size=K, dim=d, mus=mus, lmbdas=np.stack(K * [0.5 * np.eye(d)]))))
    model = M.MixtureOfMixtureOfGaussians(cluster_size=M_, mixture_size=K, dim=d, gating=D.Categorical(dim=M_), components=comps)
    npr.seed(seed)
    ll = model.max_likelihood(x, maxiter=iters, maxsubiter=subiters, progress_bar=False)
    rec['ll'] = np.array(ll)
    rec['gate_probs'] = model.gating.probs.copy()
    for m, c in enumerate(comps):
        rec[f'end_mus{m}'], rec[f'end_lmbdas{m}'], rec[f'end_probs{m}'] = c.components.mus.copy(), c.components.lmbdas.copy(), c.gating.probs.copy()
    rec['resp_end'] = model.responsibilities(x)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print(name, 'll', ll[0], '->', ll[-1], 'monotone', bool(np.all(np.diff(ll) >= -1e-8)))


def api_misc():
    """small public methods of the drivers that the trajectory fixtures do not reach: used_labels, log_marginal_likelihood,
    posterior_predictive_studentt, meanfield_update_{gating,components}, ILR resample_{basis,models},
    meanfield_predictive_activation -- on the models of gmm_toy_vi / ilr_stacked after a short seeded mean-field run."""
    g = dict(np.load(os.path.join(OUT, 'gmm_toy_vi.npz')))
    K, d = int(g['K']), int(g['d'])
    prior = D.StackedNormalWisharts(K, d, g['mus0'], g['kappas0'], g['psis0'], g['nus0'])
    npr.seed(1)
    comp = D.StackedGaussiansWithNormalWisharts(K, d, prior=prior)
    gating = D.CategoricalWithDirichlet(K, D.Dirichlet(K, g['gate_alphas0']))
    model = M.BayesianMixtureOfGaussians(gating=gating, components=comp)
    x = g['obs']
    npr.seed(17)
    model.meanfield_coordinate_descent(x, maxiter=3, tol=0., progress_bar=False)
    rec = dict(used_labels=model.used_labels(x), lml=comp.log_marginal_likelihood())
    for n, p in zip(('mus', 'lmbdas', 'dfs'), comp.posterior_predictive_studentt()):
        rec['pst_' + n] = p
    resp = model.expected_responsibilities(x)
    rec['resp'] = resp
    npr.seed(18)
    model.meanfield_update_gating(resp)
    rec['gate_alphas'] = gating.posterior.alphas.copy()
    npr.seed(19)
    model.meanfield_update_components(x, resp)
    for n, p in zip(('mus', 'kappas', 'psis', 'nus'), comp.posterior.params):
        rec['post_' + n] = p
    np.savez_compressed(os.path.join(OUT, 'api_misc_gmm.npz'), **rec)
    # ILR
    g = dict(np.load(os.path.join(OUT, 'ilr_stacked.npz')))
    K, din, o = int(g['K']), int(g['din']), int(g['o'])
    c = din + 1
    npr.seed(2)
    basis = D.StackedGaussiansWithNormalWisharts(K, din, prior=D.StackedNormalWisharts(K, din, g['b_mus0'], g['b_kappas0'], g['b_psis0'], g['b_nus0']))
    models = D.StackedLinearGaussiansWithMatrixNormalWisharts(K, c, o, D.StackedMatrixNormalWisharts(K, c, o, g['m_Ms0'], g['m_Ks0'], g['m_psis0'], g['m_nus0']), affine=True)
    gating = D.CategoricalWithStickBreaking(K, D.TruncatedStickBreaking(K, g['gate_gammas0'], g['gate_deltas0']))
    ilr = M.BayesianMixtureOfLinearGaussians(K, din, o, gating=gating, basis=basis, models=models)
    x, y = g['x'], g['y']
    npr.seed(27)
    ilr.meanfield_coordinate_descent(x, y, maxiter=3, tol=0., progress_bar=False)
    rec = dict(used_labels=ilr.used_labels(x, y), activation=ilr.meanfield_predictive_activation(x[:40]))
    z = np.argmax(ilr.expected_responsibilities(x, y), axis=0)
    rec['z'] = z.astype(np.int32)
    npr.seed(28)
    ilr.resample_basis(x, z)
    rec['b_lik_mus'], rec['b_lik_lmbdas'] = basis.likelihood.mus.copy(), basis.likelihood.lmbdas.copy()
    npr.seed(29)
    ilr.resample_models(x, y, z)
    rec['m_lik_As'], rec['m_lik_lmbdas'] = models.likelihood.As.copy(), models.likelihood.lmbdas.copy()
    np.savez_compressed(os.path.join(OUT, 'api_misc_ilr.npz'), **rec)
    print('api_misc ok: used labels', rec['used_labels'])


def hierarchical():
    rng = np.random.default_rng(77)
    K, d = 4, 2
    centres = np.array([[3., -3.], [-3., 3.], [-5., -5.], [5., 5.]])
    x = centres[rng.integers(0, K, 500)] + rng.standard_normal((500, d))
    hgmm_vi_case('hgmm_vi', x, K, False, iters=8, subiters=5, ctor_seed=1337, seed=3)
    hgmm_gibbs_case('hgmm_gibbs', x, K, False, sweeps=4, subiters=3, ctor_seed=1337, seed=4)
    hgmm_svi_case('hgmm_svi', x, K, False, iters=6, subiters=3, step_size=0.2, ctor_seed=1337, seed=5)
    x8 = blobs(rng, 400, 8, 5, spread=3.0)
    x8 = x8 @ np.linalg.inv(np.linalg.cholesky(np.cov(x8.T))).T          # one shared covariance scale: the model ties Lambda
    hgmm_vi_case('hgmm_d8_vi_stick', x8, 5, True, iters=5, subiters=4, ctor_seed=8, seed=9)
    xm = np.vstack([centres[:2][rng.integers(0, 2, 150)] + rng.standard_normal((150, d)),
                    (centres[2:][rng.integers(0, 2, 150)] + rng.standard_normal((150, d))) * np.array([1., 0.4])])
    hmom_case('hmom_vi', xm, 2, 2, iters=3, subiters=3, subsubiters=3, ctor_seed=50, seed=6)
    hmom_em_case('hmom_em', xm, 2, 2, iters=4, subiters=3, seed=12)



if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'hier':       # only the hierarchical fixtures (the others stay untouched)
        hierarchical()
    elif len(sys.argv) > 1 and sys.argv[1] == 'hilr':
        hierarchical_ilr()
    elif len(sys.argv) > 1 and sys.argv[1] == 'ilr_svi':
        ilr_svi()
    elif len(sys.argv) > 1 and sys.argv[1] == 'api_misc':
        api_misc()
    else:
        main()
        hierarchical()
        hierarchical_ilr()
        ilr_svi()
        api_misc()
