"""CPU oracle for the mimo mixture-inference sweep  --  TEST INFRASTRUCTURE ONLY.

This module is a plain NumPy/SciPy (float64) restatement of the reference's
algorithm for the sweep named in BASELINE.json.  It is the *checker* for the
CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``mimo_b200/`` imports or calls it; the product path has no CPU fallback.

Parity status: PINNED.  Every function below is checked against the unmodified
reference (imported from /root/reference in the build container) by
``tests/test_oracle_vs_reference.py`` and against the committed fixtures under
``tests/golden/`` (made by ``oracle/make_golden.py``) by
``tests/test_oracle_golden.py``.  The reference has no tests or golden vectors
of its own, so "pinned" means "pinned by executing the
reference on identical inputs".

Each function cites the reference file:line it follows (paths relative to
/root/reference/mimo).  The arithmetic follows the reference's contractions;
where the reference materialises K*N*d*d tensors (distributions/gaussian.py:
481-485) the same numbers are produced through an equivalent contraction that
never builds that tensor, and point-axis chunking is available because every
per-point quantity is independent across points and every statistic is a sum
over points (the reference's own list-of-arrays semantics,
distributions/gaussian.py:503-505).
"""

import numpy as np
from scipy.special import digamma, gammaln, betaln, multigammaln, logsumexp
from scipy.linalg import cholesky as sp_cholesky

LOG_2PI = np.log(2.0 * np.pi)


# ---------------------------------------------------------------------------
# L0 utilities
# ---------------------------------------------------------------------------

def one_hot(z, K):
    """utils/data.py:160-169 -- labels (N,) -> dense (K, N) float64 indicator."""
    z = np.atleast_1d(z).astype(int)
    if not (np.all(z >= 0) and np.all(z < K)):
        raise AssertionError("labels out of range")
    out = np.zeros((K, z.size))
    out[z, np.arange(z.size)] = 1.0
    return out


def sample_discrete_from_log(p_log, u):
    """utils/stats.py:8-21 with the uniform draw made explicit.

    ``u`` is what ``npr.random(size=(1, N))`` returns at stats.py:14.
    cdf is a sequential float64 cumsum over the component axis; the label is
    the number of cdf entries strictly below u * cdf[-1].
    """
    lognorm = logsumexp(p_log, axis=0)
    cdf = np.exp(p_log - lognorm[None, :]).cumsum(axis=0)
    thresh = np.reshape(u, (1, -1)) * cdf[-1:, :]
    return np.sum(thresh > cdf, axis=0, dtype=np.int32)


def label_boundary_distance(p_log, u):
    """min_k |u*cdf_K - cdf_k| per point: draws closer than 1e-6 to a CDF
    boundary are excluded from bit-exact label comparisons (north_star)."""
    lognorm = logsumexp(p_log, axis=0)
    cdf = np.exp(p_log - lognorm[None, :]).cumsum(axis=0)
    thresh = np.reshape(u, (1, -1)) * cdf[-1:, :]
    return np.min(np.abs(thresh - cdf), axis=0)


def responsibilities(log_joint):
    """mixtures/gmm.py:72-75, 256-259 -- softmax over the component axis."""
    lse = logsumexp(log_joint, axis=0, keepdims=True)
    return np.exp(log_joint - lse), lse[0]


# ---------------------------------------------------------------------------
# L1a likelihoods: per-point log-likelihoods
# ---------------------------------------------------------------------------

def gauss_full_loglik(x, mus, lmbdas):
    """distributions/gaussian.py:510-523 (+ log_partition :352-354, log_base
    :69-74).  x (N,d), mus (K,d), lmbdas (K,d,d) -> (K,N)."""
    d = x.shape[1]
    lin = np.einsum('kd,kdl,nl->kn', mus, lmbdas, x, optimize=True)
    quad = np.einsum('nd,kdl,nl->kn', x, lmbdas, x, optimize=True)
    out = lin - 0.5 * quad
    logpart = np.empty(mus.shape[0])
    for k in range(mus.shape[0]):
        U = sp_cholesky(lmbdas[k], lower=False)            # gaussian.py:298
        logpart[k] = 0.5 * mus[k] @ lmbdas[k] @ mus[k] - np.sum(np.log(np.diag(U)))
    return out - logpart[:, None] - 0.5 * d * LOG_2PI


def gauss_diag_loglik(x, mus, lmbdas_diags):
    """distributions/gaussian.py:837-850 (the reference builds dense diagonal
    matrices; the numbers are the same)."""
    d = x.shape[1]
    lin = (mus * lmbdas_diags) @ x.T
    quad = lmbdas_diags @ (x * x).T
    logpart = 0.5 * np.sum(mus * lmbdas_diags * mus, axis=1) \
        - np.sum(np.log(np.sqrt(lmbdas_diags)), axis=1)      # gaussian.py:679-681
    return lin - 0.5 * quad - logpart[:, None] - 0.5 * d * LOG_2PI


def _augment(x, affine):
    """lingauss.py:312-313 -- append the constant-1 column when affine."""
    return np.hstack((x, np.ones((x.shape[0], 1)))) if affine else x


def lingauss_predict(x, As, affine=True):
    """distributions/lingauss.py:251-257 -> (K,N,o)."""
    return np.einsum('kdl,nl->knd', As, _augment(x, affine), optimize=True)


def lingauss_loglik(x, y, As, lmbdas, affine=True):
    """distributions/lingauss.py:330-347 (+ log_partition :166-169, 327-328)."""
    o = y.shape[1]
    mu = lingauss_predict(x, As, affine)                       # (K,N,o)
    lin = np.einsum('knd,kdl,nl->kn', mu, lmbdas, y, optimize=True)
    quad = np.einsum('nd,kdl,nl->kn', y, lmbdas, y, optimize=True)
    out = lin - 0.5 * quad
    for k in range(As.shape[0]):
        U = sp_cholesky(lmbdas[k], lower=False)
        logpart = 0.5 * np.einsum('nd,dl,nl->n', mu[k], lmbdas[k], mu[k]) \
            - np.sum(np.log(np.diag(U)))
        out[k] -= logpart
    return out - 0.5 * o * LOG_2PI


# ---------------------------------------------------------------------------
# L1a likelihoods: weighted sufficient statistics (the all-reduced tensors)
# ---------------------------------------------------------------------------

def gauss_full_wstats(x, w):
    """distributions/gaussian.py:491-505 -> [sum r x (K,d), sum r (K,),
    sum r x x^T (K,d,d), sum r (K,)]."""
    xk = np.einsum('kn,nd->kd', w, x, optimize=True)
    xxk = np.einsum('nd,kn,nl->kdl', x, w, x, optimize=True)
    nk = np.sum(w, axis=1)
    return [xk, nk, xxk, nk]


def gauss_full_wstats_blas(x, w):
    """gauss_full_wstats with the K d x d contractions as K GEMMs (x^T diag(w_k) x): the same sums in another order,
    for the K = 1024, d = 128 parity cases where einsum's generic path takes minutes.  Pinned to gauss_full_wstats
    in tests/test_oracle_golden.py."""
    K, d = w.shape[0], x.shape[1]
    xxk = np.empty((K, d, d))
    for k in range(K):
        xxk[k] = (x * w[k][:, None]).T @ x
    nk = np.sum(w, axis=1)
    return [w @ x, nk, xxk, nk]


def gauss_diag_wstats(x, w):
    """distributions/gaussian.py:819-832 -> [sum r x, n bcast, n bcast,
    sum r x^2], all (K,d)."""
    xk = np.einsum('kn,nd->kd', w, x)
    xxk = np.einsum('nd,kn,nd->kd', x, w, x)
    ndk = np.broadcast_to(np.sum(w, axis=1, keepdims=True), xk.shape).copy()
    return [xk, ndk, ndk.copy(), xxk]


def lingauss_wstats(x, y, w, affine=True):
    """distributions/lingauss.py:306-325 -> [sum r y xt^T (K,o,c),
    sum r xt xt^T (K,c,c), sum r y y^T (K,o,o), sum r (K,)]."""
    xt = _augment(x, affine)
    yx = np.einsum('nd,kn,nl->kdl', y, w, xt, optimize=True)
    xx = np.einsum('nd,kn,nl->kdl', xt, w, xt, optimize=True)
    yy = np.einsum('nd,kn,nl->kdl', y, w, y, optimize=True)
    return [yx, xx, yy, np.sum(w, axis=1)]


def categorical_wstats(w):
    """distributions/categorical.py:41-46."""
    return np.sum(np.atleast_2d(w), axis=1)


def categorical_stats(labels, K):
    """distributions/categorical.py:35-39."""
    return np.bincount(labels, minlength=K)


def add_stats(a, b):
    """utils/abstraction.py:12-14."""
    return [ai + bi for ai, bi in zip(a, b)]


# ---------------------------------------------------------------------------
# L1b Normal-Wishart (stacked: leading axis K)
# ---------------------------------------------------------------------------

def nw_std_to_nat(mus, kappas, psis, nus):
    """distributions/composite.py:50-65 per component (stack :174-178)."""
    d = mus.shape[1]
    a = kappas[:, None] * mus
    c = np.linalg.inv(psis) + kappas[:, None, None] * np.einsum('kd,kl->kdl', mus, mus)
    return [a, kappas.copy(), c, nus - d]


def nw_nat_to_std(nat, tied=False):
    """composite.py:67-72 (stacked :180-184); tied :275-283."""
    a, b, c, e = nat
    d = a.shape[1]
    mus = a / b[:, None]
    inner = c - b[:, None, None] * np.einsum('kd,kl->kdl', mus, mus)
    if tied:
        psi = np.linalg.inv(np.mean(inner, axis=0))
        nu = np.mean(e + d)
        K = a.shape[0]
        return mus, b.copy(), np.array(K * [psi]), np.array(K * [nu])
    return mus, b.copy(), np.linalg.inv(inner), e + d


def wishart_expected_logdet(psis, nus):
    """distributions/wishart.py:139-143 / composite.py:115-116."""
    d = psis.shape[-1]
    out = np.empty(psis.shape[0])
    for k in range(psis.shape[0]):
        C = np.linalg.cholesky(psis[k])                      # wishart.py:59-62
        out[k] = np.sum(digamma((nus[k] - np.arange(d)) / 2.0)) \
            + d * np.log(2.0) + 2.0 * np.sum(np.log(np.diag(C)))
    return out


def nw_expected_statistics(mus, kappas, psis, nus):
    """composite.py:106-118 (stacked :248-250)."""
    d = mus.shape[1]
    E_lm = nus[:, None] * np.einsum('kdl,kl->kd', psis, mus)
    E_mlm = -0.5 * (d / kappas + np.einsum('kd,kd->k', mus, E_lm))
    E_l = -0.5 * nus[:, None, None] * psis
    E_logdet = 0.5 * wishart_expected_logdet(psis, nus)
    return E_lm, E_mlm, E_l, E_logdet


def nw_expected_loglik(x, mus, kappas, psis, nus):
    """distributions/bayesian.py:287-301.  Same four terms; the per-point
    statistics of gaussian.py:466-485 (x, 1, x x^T, 1 replicated K times) are
    contracted on the fly instead of being materialised."""
    d = x.shape[1]
    E_lm, E_mlm, E_l, E_logdet = nw_expected_statistics(mus, kappas, psis, nus)
    out = E_lm @ x.T
    out += E_mlm[:, None]
    out += np.einsum('kdl,nd,nl->kn', E_l, x, x, optimize=True)
    out += E_logdet[:, None]
    return out - 0.5 * d * LOG_2PI


def nw_expected_loglik_blas(x, mus, kappas, psis, nus):
    """nw_expected_loglik with the quadratic term x^T E_l[k] x as one GEMM per component (same contraction, BLAS
    order); pinned to nw_expected_loglik in tests/test_oracle_golden.py."""
    d = x.shape[1]
    E_lm, E_mlm, E_l, E_logdet = nw_expected_statistics(mus, kappas, psis, nus)
    out = E_lm @ x.T
    out += E_mlm[:, None]
    for k in range(mus.shape[0]):
        out[k] += np.einsum('nd,nd->n', x @ E_l[k], x)
    out += E_logdet[:, None]
    return out - 0.5 * d * LOG_2PI


def wishart_log_partition(psis, nus):
    """distributions/wishart.py:129-132."""
    d = psis.shape[-1]
    out = np.empty(psis.shape[0])
    for k in range(psis.shape[0]):
        C = np.linalg.cholesky(psis[k])
        out[k] = 0.5 * nus[k] * d * np.log(2.0) + multigammaln(nus[k] / 2.0, d) \
            + nus[k] * np.sum(np.log(np.diag(C)))
    return out


def nw_log_partition(mus, kappas, psis, nus):
    """composite.py:95-98."""
    d = mus.shape[1]
    return -0.5 * d * np.log(kappas) + wishart_log_partition(psis, nus)


def _nw_dot(nat, stats):
    return np.einsum('kd,kd->k', nat[0], stats[0]) + nat[1] * stats[1] \
        + np.einsum('kdl,kdl->k', nat[2], stats[2]) + nat[3] * stats[3]


def nw_vlb(prior, post):
    """bayesian.py:240-243 with composite.py:120-134: entropy(q) - cross_entropy(q, p),
    per component.  prior/post are (mus, kappas, psis, nus) tuples.  The
    log-base terms (composite.py:88-93) cancel between the two."""
    stats = nw_expected_statistics(*post)
    ent = nw_log_partition(*post) - _nw_dot(nw_std_to_nat(*post), stats)
    xent = nw_log_partition(*prior) - _nw_dot(nw_std_to_nat(*prior), stats)
    return ent - xent


def nw_mode(mus, kappas, psis, nus):
    """composite.py:77-80."""
    d = mus.shape[1]
    return mus.copy(), (nus - d)[:, None, None] * psis


def wishart_rvs_from_variates(psi, normals, chisq):
    """distributions/wishart.py:72-92 with the variates made explicit:
    normals = npr.normal(size=d(d-1)/2) in np.tril_indices(d, -1) order,
    chisq[i] = npr.chisquare(nu - i) (wishart.py:78-79)."""
    d = psi.shape[0]
    A = np.zeros((d, d))
    A[np.tril_indices(d, k=-1)] = normals
    A[np.diag_indices(d)] = np.sqrt(chisq)
    T = np.linalg.cholesky(psi) @ A
    return T @ T.T


def nw_rvs_from_variates(mu, kappa, psi, nu, normals, chisq, z):
    """composite.py:82-86 + gaussian.py:295-313: lmbda ~ W(psi, nu);
    mu ~ N(m, (kappa lmbda)^-1) via the upper Cholesky factor of kappa*lmbda."""
    lmbda = wishart_rvs_from_variates(psi, normals, chisq)
    U = sp_cholesky(kappa * lmbda, lower=False)
    return mu + np.linalg.inv(U) @ z, lmbda


# ---------------------------------------------------------------------------
# L1b Normal-Gamma (diagonal path)
# ---------------------------------------------------------------------------

def ng_std_to_nat(mus, kappas, alphas, betas):
    """composite.py:313-329 (all (K,d))."""
    return [kappas * mus, kappas.copy(), 2.0 * alphas - 1.0, 2.0 * betas + kappas * mus ** 2]


def ng_nat_to_std(nat, tied=False):
    """composite.py:331-337; tied :539-547."""
    a, b, c, e = nat
    mus = a / b
    alphas = 0.5 * (c + 1.0)
    betas = 0.5 * (e - b * mus ** 2)
    if tied:
        K = a.shape[0]
        alphas = np.array(K * [np.mean(alphas, axis=0)])
        betas = np.array(K * [np.mean(betas, axis=0)])
    return mus, b.copy(), alphas, betas


def ng_expected_statistics(mus, kappas, alphas, betas):
    """composite.py:371-382."""
    E_lm = alphas / betas * mus
    E_lmm = -0.5 * (1.0 / kappas + mus * E_lm)
    E_logl = 0.5 * (digamma(alphas) - np.log(betas))
    E_l = -0.5 * (alphas / betas)
    return E_lm, E_lmm, E_logl, E_l


def ng_expected_loglik(x, mus, kappas, alphas, betas):
    """bayesian.py:446-460 with the per-point stats of gaussian.py:794-813
    (x, 1, 1, x^2)."""
    d = x.shape[1]
    E_lm, E_lmm, E_logl, E_l = ng_expected_statistics(mus, kappas, alphas, betas)
    out = E_lm @ x.T + E_l @ (x * x).T
    out += np.sum(E_lmm, axis=1)[:, None] + np.sum(E_logl, axis=1)[:, None]
    return out - 0.5 * d * LOG_2PI


def ng_log_partition(mus, kappas, alphas, betas):
    """composite.py:360-363 + gamma.py:92-93."""
    return -0.5 * np.sum(np.log(kappas), axis=1) \
        + np.sum(gammaln(alphas) - alphas * np.log(betas), axis=1)


def ng_vlb(prior, post):
    """bayesian.py:401-404 with composite.py:384-398."""
    stats = ng_expected_statistics(*post)

    def dot(nat):
        return sum(np.sum(n * s, axis=1) for n, s in zip(nat, stats))
    ent = ng_log_partition(*post) - dot(ng_std_to_nat(*post))
    xent = ng_log_partition(*prior) - dot(ng_std_to_nat(*prior))
    return ent - xent


def ng_mode(mus, kappas, alphas, betas):
    """composite.py:342-345."""
    return mus.copy(), (alphas - 0.5) / betas


def ng_rvs_from_variates(mu, kappas, alphas, betas, g, z):
    """composite.py:347-351: lmbda_diag = gamma draw (g = npr.gamma(alphas,
    1/betas), gamma.py:54-56); mu ~ N(m, 1/(kappa lmbda)) (gaussian.py:644-646)."""
    return mu + z / np.sqrt(kappas * g), g


# ---------------------------------------------------------------------------
# L1b Matrix-Normal-Wishart (linear-Gaussian experts)
# ---------------------------------------------------------------------------

def mnw_std_to_nat(Ms, Ks, psis, nus):
    """composite.py:577-592."""
    o, c = Ms.shape[1], Ms.shape[2]
    a = np.einsum('kdl,klm->kdm', Ms, Ks)
    cc = np.linalg.inv(psis) + np.einsum('kdl,klm,khm->kdh', Ms, Ks, Ms)
    return [a, Ks.copy(), cc, nus - o - 1.0 + c]


def mnw_nat_to_std(nat, tied=False):
    """composite.py:594-599; tied :800-808."""
    a, b, cc, e = nat
    o, c = a.shape[1], a.shape[2]
    Ms = np.einsum('kdl,klh->kdh', a, np.linalg.inv(b))
    inner = cc - np.einsum('kdl,klm,khm->kdh', Ms, b, Ms)
    if tied:
        K = a.shape[0]
        psi = np.linalg.inv(np.mean(inner, axis=0))
        nu = np.mean(e + o + 1 - c)
        return Ms, b.copy(), np.array(K * [psi]), np.array(K * [nu])
    return Ms, b.copy(), np.linalg.inv(inner), e + o + 1.0 - c


def mnw_expected_statistics(Ms, Ks, psis, nus):
    """composite.py:635-647."""
    o = Ms.shape[1]
    E_LA = nus[:, None, None] * np.einsum('kdl,klm->kdm', psis, Ms)
    E_ALA = -0.5 * (o * np.linalg.inv(Ks) + np.einsum('kdl,kdm->klm', Ms, E_LA))
    E_L = -0.5 * nus[:, None, None] * psis
    E_logdet = 0.5 * wishart_expected_logdet(psis, nus)
    return E_LA, E_ALA, E_L, E_logdet


def mnw_expected_loglik(x, y, Ms, Ks, psis, nus, affine=True):
    """bayesian.py:933-947 with the per-point stats of lingauss.py:275-300
    (y xt^T, xt xt^T, y y^T, 1) contracted on the fly."""
    o = y.shape[1]
    xt = _augment(x, affine)
    E_LA, E_ALA, E_L, E_logdet = mnw_expected_statistics(Ms, Ks, psis, nus)
    out = np.einsum('kdl,nd,nl->kn', E_LA, y, xt, optimize=True)
    out += np.einsum('kdl,nd,nl->kn', E_ALA, xt, xt, optimize=True)
    out += np.einsum('kdl,nd,nl->kn', E_L, y, y, optimize=True)
    out += E_logdet[:, None]
    return out - 0.5 * o * LOG_2PI


def mnw_log_partition(Ms, Ks, psis, nus):
    """composite.py:622-625."""
    o = Ms.shape[1]
    return -0.5 * o * np.linalg.slogdet(Ks)[1] + wishart_log_partition(psis, nus)


def mnw_vlb(prior, post):
    """bayesian.py:854-857 with composite.py:649-663."""
    stats = mnw_expected_statistics(*post)

    def dot(nat):
        return np.einsum('kdl,kdl->k', nat[0], stats[0]) + np.einsum('kdl,kdl->k', nat[1], stats[1]) \
            + np.einsum('kdl,kdl->k', nat[2], stats[2]) + nat[3] * stats[3]
    ent = mnw_log_partition(*post) - dot(mnw_std_to_nat(*post))
    xent = mnw_log_partition(*prior) - dot(mnw_std_to_nat(*prior))
    return ent - xent


def mnw_mode(Ms, Ks, psis, nus):
    """composite.py:604-607."""
    o = Ms.shape[1]
    return Ms.copy(), (nus - o)[:, None, None] * psis


def mnw_rvs_from_variates(M, K, psi, nu, normals, chisq, z):
    """composite.py:609-613 + matrix.py:98-125: lmbda ~ W(psi, nu); A = M +
    unvec_F(U^-1 z) with U the upper Cholesky factor of kron(K, lmbda)."""
    o, c = M.shape
    lmbda = wishart_rvs_from_variates(psi, normals, chisq)
    U = sp_cholesky(np.kron(K, lmbda), lower=False)
    aux = z @ np.linalg.inv(U).T
    return M + np.reshape(aux, (o, c), order='F'), lmbda


# ---------------------------------------------------------------------------
# L1b gating: Dirichlet and truncated stick-breaking
# ---------------------------------------------------------------------------

def dirichlet_posterior(alphas0, counts):
    """bayesian.py:70-72, 78-81 with dirichlet.py:30-38 (nat = alpha - 1)."""
    return alphas0 + counts


def dirichlet_expected_log(alphas):
    """dirichlet.py:85-87."""
    return digamma(alphas) - digamma(np.sum(alphas))


def dirichlet_log_partition(alphas):
    """dirichlet.py:78-79."""
    return np.sum(gammaln(alphas)) - gammaln(np.sum(alphas))


def dirichlet_vlb(alphas0, alphas):
    """bayesian.py:93-96 with dirichlet.py:89-97."""
    s = dirichlet_expected_log(alphas)
    ent = dirichlet_log_partition(alphas) - (alphas - 1.0) @ s
    xent = dirichlet_log_partition(alphas0) - (alphas0 - 1.0) @ s
    return ent - xent


def dirichlet_mode(alphas):
    """dirichlet.py:43-45."""
    if not np.all(alphas > 1.0):
        raise AssertionError("Make sure alphas > 1.")
    return (alphas - 1.0) / (np.sum(alphas) - alphas.size)


def dirichlet_probs_from_gammas(g):
    """npr.dirichlet (dirichlet.py:47-48) = normalised Gamma(alpha_k, 1)
    draws; the Gibbs step clips to >= spacing(1) (bayesian.py:75)."""
    return np.clip(g / np.sum(g), np.spacing(1.0), np.inf)


def stick_posterior(gammas0, deltas0, counts):
    """bayesian.py:140-146, 151-157: Blei & Jordan tail counts."""
    acc = np.hstack((np.cumsum(counts[::-1])[-2::-1], 0))
    return gammas0 + counts, deltas0 + acc


def stick_expected_log(gammas, deltas):
    """dirichlet.py:201-204 and the prefix sum of gmm.py:250-252."""
    E_stick = digamma(gammas) - digamma(gammas + deltas)
    E_rest = digamma(deltas) - digamma(gammas + deltas)
    return E_stick + np.hstack((0, np.cumsum(E_rest)[:-1])), E_stick, E_rest


def stick_vlb(prior, post):
    """bayesian.py:173-176 with dirichlet.py:195-214."""
    g, dl = post
    g0, d0 = prior
    _, E_stick, E_rest = stick_expected_log(g, dl)
    ent = np.sum(betaln(g, dl)) - ((g - 1.0) @ E_stick + (dl - 1.0) @ E_rest)
    xent = np.sum(betaln(g0, d0)) - ((g0 - 1.0) @ E_stick + (d0 - 1.0) @ E_rest)
    return ent - xent


def stick_probs_from_betas(v):
    """dirichlet.py:177-186: v = npr.beta(gammas[:-1], deltas[:-1])."""
    v = np.hstack((v, 1.0))
    probs = np.empty(v.size)
    probs[0] = v[0]
    probs[1:] = v[1:] * np.cumprod(1.0 - v[:-1])
    return probs


def stick_mean(gammas, deltas):
    """dirichlet.py:141-150."""
    return stick_probs_from_betas(gammas[:-1] / (gammas[:-1] + deltas[:-1]))


# ---------------------------------------------------------------------------
# EM M-steps
# ---------------------------------------------------------------------------

def gauss_full_mstep(stats):
    """gaussian.py:525-542."""
    xk, nk, xxk, _ = stats
    K, d = xk.shape
    mus = xk / nk[:, None]
    lmbdas = np.empty((K, d, d))
    for k in range(K):
        sigma = xxk[k] / nk[k] - np.outer(mus[k], mus[k])
        sigma = 0.5 * (sigma + sigma.T) + 1e-16 * np.eye(d)
        if not np.all(np.linalg.eigvalsh(sigma) > 0.0):
            raise AssertionError("covariance not positive definite")
        lmbdas[k] = np.linalg.inv(sigma)
    return mus, lmbdas


def gauss_diag_mstep(stats):
    """gaussian.py:852-862."""
    xk, ndk, _, xxk = stats
    mus = xk / ndk
    return mus, 1.0 / (xxk / ndk - mus ** 2 + 1e-16)


def lingauss_mstep(stats):
    """lingauss.py:350-367."""
    yx, xx, yy, nk = stats
    K, o, c = yx.shape
    As = np.empty((K, o, c))
    lmbdas = np.empty((K, o, o))
    for k in range(K):
        As[k] = np.linalg.solve(xx[k], yx[k].T).T
        sigma = (yy[k] - As[k] @ yx[k].T) / nk[k]
        sigma = 0.5 * (sigma + sigma.T) + 1e-16 * np.eye(o)
        if not np.all(np.linalg.eigvalsh(sigma) > 0.0):
            raise AssertionError("covariance not positive definite")
        lmbdas[k] = np.linalg.inv(sigma)
    return As, lmbdas


def categorical_mstep(counts):
    """categorical.py:65-68."""
    return counts / counts.sum()


# ---------------------------------------------------------------------------
# L3 sweep pieces (mixtures/gmm.py, mixtures/ilr.py)
# ---------------------------------------------------------------------------

def vlb_labels_dirichlet(resp, E_log_pi):
    """gmm.py:341-344, 352-356."""
    with np.errstate(invalid='ignore', divide='ignore'):
        return np.sum(resp * E_log_pi[:, None]) - np.nansum(resp * np.log(resp))


def vlb_labels_stick(resp, E_stick, E_rest):
    """gmm.py:345-356."""
    acc = np.vstack((np.cumsum(resp[::-1, :], axis=0)[-2::-1, :], np.zeros((1, resp.shape[1]))))
    with np.errstate(invalid='ignore', divide='ignore'):
        return np.sum(resp * E_stick[:, None] + acc * E_rest[:, None]) - np.nansum(resp * np.log(resp))


# ---- prediction path (mixtures/ilr.py:325-430) ------------------------------------------------------------------
def nw_predictive(mus, kappas, psis, nus):
    """distributions/bayesian.py:303-308, 314-320: (mus, lmbdas, dfs) of the Normal-Wishart posterior predictive."""
    d = mus.shape[-1]
    dfs = nus - d + 1
    return mus, (dfs / (1. + 1. / kappas))[:, None, None] * psis, dfs


def studentt_loglik_reference_form(x, mus, lmbdas, dfs):
    """utils/stats.py:67-79 with the broadcasting the shapes call for (the reference divides a (K, N) array by a (K,)
    vector and raises for K != N).  The form is the reference's own: the -(df + d)/2 factor sits in the constant."""
    d = mus.shape[-1]
    xc = x[:, None, :] - mus[None, :, :]
    deltas = np.einsum('nkd,kdl,nkl->kn', xc, lmbdas, xc)
    aux = gammaln((dfs + d) / 2.) - gammaln(dfs / 2.) + 0.5 * np.linalg.slogdet(lmbdas)[1] \
        - (d / 2.) * np.log(dfs * np.pi) - 0.5 * (dfs + d)
    return aux[:, None] + np.log1p(deltas / dfs[:, None])


def ilr_predictive_weights(x, gating_mean, basis_post, dist='gaussian'):
    """ilr.py:337-346."""
    mus, lmbdas, dfs = nw_predictive(*basis_post)
    lp = gauss_full_loglik(x, mus, lmbdas) if dist == 'gaussian' else studentt_loglik_reference_form(x, mus, lmbdas, dfs)
    return responsibilities(np.log(gating_mean)[:, None] + lp)[0]


def mnw_predictive(x, Ms, Ks, psis, nus, affine=True):
    """bayesian.py:949-985: mus (K, N, o), lmbdas (K, N, o, o), dfs (K,)."""
    xt = _augment(x, affine)
    o = Ms.shape[1]
    dfs = nus - o + 1
    mus = np.einsum('kdl,nl->knd', Ms, xt)
    cs = 1. + np.einsum('nd,kdl,nl->kn', xt, np.linalg.inv(Ks), xt)
    return mus, np.einsum('kdl,k,kn->kndl', psis, dfs, 1. / cs), dfs


def ilr_prediction(x, weights, models_post, prediction='average', dist='gaussian', y=None, eps=np.finfo(np.float64).tiny):
    """ilr.py:348-411 with the evident shapes: (mu (N, o), covar (N, o, o), nlpd (N) or None)."""
    mus, lmbdas, dfs = mnw_predictive(x, *models_post)
    covars = np.linalg.inv(lmbdas)
    if dist == 'studentt':
        covars = covars * (dfs / (dfs - 2))[:, None, None, None]
    if prediction == 'mode':
        k = np.argmax(weights, axis=0)
        n = np.arange(len(k))
        mu, covar = mus[k, n], covars[k, n]
    else:
        mu = np.einsum('knd,kn->nd', mus, weights)
        covar = np.einsum('kndl,kn->ndl', covars + np.einsum('knd,knl->kndl', mus, mus), weights) - np.einsum('nd,nl->ndl', mu, mu)
    nlpd = None
    if y is not None:
        diff = y[None, :, :] - mus
        o = mus.shape[-1]
        log_pl = -0.5 * np.einsum('knd,kndl,knl->kn', diff, lmbdas, diff) + 0.5 * np.linalg.slogdet(lmbdas)[1] - 0.5 * o * np.log(2. * np.pi)
        nlpd = -logsumexp(log_pl + np.log(weights + eps), axis=0)
    return mu, covar, nlpd


# ---------------------------------------------------------------------------
# hierarchical Normal-Wishart: K tied Gaussians whose means have scaled-precision priors N(tau, (kappa_k Lambda)^-1)
# and whose shared (tau, Lambda) has a Normal-Wishart hyper-prior -- SURVEY 8 f4 (mixtures/hgmm.py:118-295)
# ---------------------------------------------------------------------------

def hnw_hyper_update(hyper_prior, kappas0, mus, xk, nk, xxk):
    """distributions/bayesian.py:671-684 (the same block at :643-656 and :709-722): hyper-posterior (rho, kappa, psi, nu)
    of the shared (tau, Lambda) given the component means `mus` and the weighted statistics.  hyper_prior = (mu0, kappa0,
    psi0, nu0) of one Normal-Wishart; kappas0 (K,) are the scaled-precision prior's kappas."""
    mu0, k0, psi0, nu0 = hyper_prior
    K = mus.shape[0]
    rho = np.sum(kappas0[:, None] * mus + k0 * mu0, axis=0) / np.sum(kappas0 + k0)
    kappa = np.sum(kappas0 + k0) / K
    dm = mu0[None, :] - mus
    spread = np.sum((k0 * kappas0 / (k0 + kappas0))[:, None, None] * np.einsum('kd,kl->kdl', dm, dm), axis=0) / K
    inner = np.linalg.inv(psi0) + spread + np.sum(xxk, axis=0) / K - np.einsum('kd,kl->dl', mus, xk) / K \
        - np.einsum('kd,kl->dl', xk, mus) / K + np.einsum('k,kd,kl->dl', nk, mus, mus) / K
    return rho, kappa, np.linalg.inv(inner), np.sum(nu0 + nk + 1) / K


def hnw_meanfield_update(hyper_prior, hyper_post, kappas0, xk, nk, xxk, nb_iter):
    """bayesian.py:662-684: nb_iter alternations of the variational E-step on the component means (their posterior
    precision scale kappa_k + n_k, mean pulled towards the hyper-posterior's location) and the hyper-posterior update.
    Returns (posterior mus, posterior kappas, hyper-posterior params)."""
    post_k = kappas0 + nk
    mus = None
    for _ in range(nb_iter):
        mus = (kappas0[:, None] * hyper_post[0][None, :] + xk) / post_k[:, None]
        hyper_post = hnw_hyper_update(hyper_prior, kappas0, mus, xk, nk, xxk)
    return mus, post_k, hyper_post


def hnw_expected_loglik(x, hyper_post, post_mus, post_omegas):
    """bayesian.py:731-749: E_q log N(x | mu_k, Lambda) with Lambda ~ W(psi, nu) of the hyper-posterior and
    mu_k ~ N(post_mus[k], post_omegas[k]^-1).  post_omegas = posterior.kappas * posterior.lmbdas (gaussian.py:954-955)."""
    _, _, psi, nu = hyper_post
    d = x.shape[1]
    EL = nu * psi
    E_logdet = 0.5 * wishart_expected_logdet(psi[None], np.array([nu]))[0]
    dx = x[None, :, :] - post_mus[:, None, :]
    quad = np.einsum('knd,dl,knl->kn', dx, EL, dx, optimize=True)
    tr = np.einsum('dl,kld->k', EL, np.linalg.inv(post_omegas))
    return -0.5 * d * LOG_2PI + E_logdet - 0.5 * quad - 0.5 * tr[:, None]


def scaled_gaussian_entropy(omega):
    """distributions/gaussian.py:1029-1031."""
    d = omega.shape[0]
    return 0.5 * d * np.log(2.0 * np.pi * np.e) - np.sum(np.log(np.diag(sp_cholesky(omega, lower=False))))


def hnw_vlb(hyper_prior, hyper_post, kappas0, post_mus, post_omegas, entropy_omegas=None):
    """bayesian.py:751-781: the component-parameter terms of the lower bound; the hyper-posterior's entropy and
    cross-entropy are counted once per component, as the reference's loop does.  entropy_omegas: the precisions the
    entropy term sees -- the reference caches the Cholesky factor of kappa * lmbda (gaussian.py:957-961) and only a
    new lmbda resets the cache (:947-951), so within a mean-field run the entropy keeps the factor of the FIRST
    evaluation while kappa moves on (quirk q11); None = the current precisions."""
    K, d = post_mus.shape
    stack = lambda p: (np.asarray(p[0])[None], np.atleast_1d(p[1]), np.asarray(p[2])[None], np.atleast_1d(p[3]))  # noqa: E731
    hyper = nw_vlb(stack(hyper_prior), stack(hyper_post))[0]
    rho, kap, psi, nu = hyper_post
    E_logdet = wishart_expected_logdet(psi[None], np.array([nu]))[0]
    vlb = 0.0
    for k in range(K):
        vlb += hyper + scaled_gaussian_entropy((post_omegas if entropy_omegas is None else entropy_omegas)[k]) - 0.5 * d * LOG_2PI
        vlb += 0.5 * d * np.log(kappas0[k]) + 0.5 * E_logdet - 0.5 * kappas0[k] * d / kap
        dm = post_mus[k] - rho
        EL = kappas0[k] * nu * psi
        vlb += -0.5 * dm @ EL @ dm - 0.5 * np.trace(EL @ np.linalg.inv(post_omegas[k]))
    return vlb


def hnw_resample(hyper_prior, hyper_post, kappas0, xk, nk, xxk, nb_iter, normal, chisquare):
    """bayesian.py:623-659 with the random draws made explicit: normal(size) / chisquare(df) are called in the
    reference's order -- per sub-iteration K hyper-posterior draws (Wishart: normal(d(d-1)/2), d chisquare; then
    normal(d) for tau: composite.py:82-86), then K draws normal(d) of the component means (gaussian.py:973-975).
    Returns (mus, lmbdas, posterior (mus, kappas), hyper-posterior params)."""
    K, d = xk.shape
    nt = d * (d - 1) // 2
    mus = lmbdas = post_mus = None
    post_k = kappas0 + nk
    for _ in range(nb_iter):
        rho, kap, psi, nu = hyper_post
        taus, lmbdas = np.zeros((K, d)), np.zeros((K, d, d))
        for k in range(K):
            nrm = normal(nt)
            chi = np.array([chisquare(nu - i) for i in range(d)])
            taus[k], lmbdas[k] = nw_rvs_from_variates(rho, kap, psi, nu, nrm, chi, normal(d))
        post_mus = (kappas0[:, None] * taus + xk) / post_k[:, None]
        mus = np.zeros((K, d))
        for k in range(K):
            U = sp_cholesky(post_k[k] * lmbdas[k], lower=False)
            mus[k] = post_mus[k] + normal(d) @ np.linalg.inv(U).T
        hyper_post = hnw_hyper_update(hyper_prior, kappas0, mus, xk, nk, xxk)
    return mus, lmbdas, (post_mus, post_k), hyper_post


def hgmm_log_weights(gating):
    """mixtures/hgmm.py:171-176: expected log gating weights; gating = ('dirichlet', alphas) | ('stick', gammas, deltas)."""
    if gating[0] == 'dirichlet':
        return dirichlet_expected_log(gating[1])
    return stick_expected_log(gating[1], gating[2])[0]


def hgmm_meanfield(obs, resp, gating_prior, hyper_prior, kappas0, post_lmbdas, iters, subiters, hyper_post=None,
                   stale_entropy=True):
    """mixtures/hgmm.py:186-225, 264-289: mean-field coordinate descent of the mixture with a hierarchical prior.
    gating_prior = ('dirichlet', alphas0) | ('stick', gammas0, deltas0); post_lmbdas (K, d, d) are the precisions the
    posterior object of the component means carries (the constructor's draw: the mean-field update never refreshes
    them, bayesian.py:662-689).  stale_entropy: quirk q11 of hnw_vlb, as the reference behaves.
    Returns dict(vlb, mus, kappas, hyper, gating, resp, ell)."""
    hyper_post = hyper_prior if hyper_post is None else hyper_post
    vlb, out, ent_omegas = [], {}, None
    for _ in range(iters):
        xk, nk, xxk, _ = gauss_full_wstats(obs, resp)
        mus, post_k, hyper_post = hnw_meanfield_update(hyper_prior, hyper_post, kappas0, xk, nk, xxk, subiters)
        omegas = post_k[:, None, None] * post_lmbdas
        counts = np.sum(resp, axis=1)
        if gating_prior[0] == 'dirichlet':
            gpost = ('dirichlet', dirichlet_posterior(gating_prior[1], counts))
            gv = dirichlet_vlb(gating_prior[1], gpost[1])
        else:
            gpost = ('stick',) + tuple(stick_posterior(gating_prior[1], gating_prior[2], counts))
            gv = stick_vlb(gating_prior[1:], gpost[1:])
        ell = hnw_expected_loglik(obs, hyper_post, mus, omegas)
        joint = ell + hgmm_log_weights(gpost)[:, None]
        resp = responsibilities(joint)[0]
        if gpost[0] == 'dirichlet':
            lab = vlb_labels_dirichlet(resp, dirichlet_expected_log(gpost[1]))
        else:
            _, Es, Er = stick_expected_log(gpost[1], gpost[2])
            lab = vlb_labels_stick(resp, Es, Er)
        if ent_omegas is None or not stale_entropy:
            ent_omegas = omegas
        vlb.append(gv + hnw_vlb(hyper_prior, hyper_post, kappas0, mus, omegas, ent_omegas) + lab + np.sum(resp * ell))
        out = dict(mus=mus, kappas=post_k, hyper=hyper_post, gating=gpost, resp=resp, ell=joint)
    out['vlb'] = np.array(vlb)
    return out


# ---------------------------------------------------------------------------
# tied-slope affine experts y = A x + c_k + eps (shared A and Lambda, own offsets) -- SURVEY 8 f4 (mixtures/hilr.py:79-291)
# ---------------------------------------------------------------------------

def tam_slope_precision(slope_prior, prec_prior, off_prior, cs, x, y, w):
    """distributions/bayesian.py:1301-1335 (mean field; the Gibbs block :1262-1296 is the same arithmetic): slope
    posterior (M, K) and precision posterior (psi, nu) given the offsets cs (K, o) and weights w (K, N), contracted over
    the points like the reference.  slope_prior = (M0, K0), prec_prior = (psi0, nu0), off_prior = (mus0, kappas0)."""
    M0, K0 = slope_prior
    psi0, nu0 = prec_prior
    mus0, kap0 = off_prior
    Kn = w.shape[0]
    nk = np.sum(w, axis=1)
    yx = np.einsum('nd,kn,nl->kdl', y, w, x, optimize=True)
    xx = np.einsum('nd,kn,nl->kdl', x, w, x, optimize=True)
    cx = np.einsum('kd,kn,nl->kdl', cs, w, x, optimize=True)
    G = (M0 @ K0)[None] + yx - cx
    Kinv = np.linalg.inv(K0[None] + xx)
    M = np.sum(np.einsum('kdl,klh->kdh', G, Kinv), axis=0) / Kn
    K = np.sum(K0[None] + xx, axis=0) / Kn
    dy = y[None, :, :] - cs[:, None, :]
    psi = np.linalg.inv(np.linalg.inv(psi0) + M0 @ K @ M0.T
                        + np.sum(np.einsum('kn,knd,knl->kdl', w, dy, dy), axis=0) / Kn
                        + np.sum(np.einsum('k,kd,kl->kdl', kap0, cs - mus0, cs - mus0), axis=0) / Kn
                        - np.sum(np.einsum('kdl,klm,khm->kdh', G, Kinv, G), axis=0) / Kn)
    return M, K, psi, np.sum(nu0 + nk + 1) / Kn


def tam_meanfield_update(slope_prior, prec_prior, off_prior, off_post_mus, x, y, w, nb_iter):
    """bayesian.py:1298-1383: nb_iter alternations; returns (slope (M, K), precision (psi, nu), offsets (mus, kappas))."""
    mus0, kap0 = off_prior
    xm, ym, nk = w @ x, w @ y, np.sum(w, axis=1)
    M = K = psi = nu = None
    mus = off_post_mus
    for _ in range(nb_iter):
        M, K, psi, nu = tam_slope_precision(slope_prior, prec_prior, off_prior, mus, x, y, w)
        mus = (kap0[:, None] * mus0 + ym - xm @ M.T) / (kap0 + nk)[:, None]
    return (M, K), (psi, nu), (mus, kap0 + nk)


def tam_joint(slope, offsets, precision):
    """bayesian.py:1395-1417: the experts [A | c_k] as stacked Matrix-Normal-Wisharts (Ms, Ks, psis, nus) with the
    block-diagonal column precision diag(K, kappa_k)."""
    M, K = slope
    mus, kappas = offsets
    psi, nu = precision
    Kn, c = mus.shape[0], K.shape[0]
    Ms = np.stack([np.hstack((M, mus[k][:, None])) for k in range(Kn)])
    Ks = np.zeros((Kn, c + 1, c + 1))
    Ks[:, :c, :c] = K
    Ks[:, c, c] = kappas
    return Ms, Ks, np.stack(Kn * [psi]), np.array(Kn * [nu], dtype=float)


def tam_expected_loglik(x, y, slope, offsets, precision):
    """bayesian.py:1388-1417."""
    return mnw_expected_loglik(x, y, *tam_joint(slope, offsets, precision), affine=True)


def tam_vlb(slope_prior, off_prior, prec_prior, slope, offsets, precision):
    """bayesian.py:1448-1479, per component."""
    return mnw_vlb(tam_joint(slope_prior, off_prior, prec_prior), tam_joint(slope, offsets, precision))


def hilr_meanfield(x, y, resp, gating_prior, basis, models, iters, subiters):
    """mixtures/hilr.py:175-218: basis = dict(hyper_prior, kappas0, post_lmbdas) of the input densities (hnw_*),
    models = dict(slope_prior, prec_prior, off_prior, off_post_mus) of the experts.  Returns the end state and the
    public lower bound (hilr.py:283-290) evaluated once, after the last iteration (the loop itself never evaluates it,
    hilr.py:194, so quirk q11 of hnw_vlb does not arise: the entropy sees the current precisions)."""
    hq = basis['hyper_prior']
    off_mus = models['off_post_mus']
    out = {}
    for _ in range(iters):
        xk, nk, xxk, _ = gauss_full_wstats(x, resp)
        bm, bk, hq = hnw_meanfield_update(basis['hyper_prior'], hq, basis['kappas0'], xk, nk, xxk, subiters)
        slope, prec, offs = tam_meanfield_update(models['slope_prior'], models['prec_prior'], models['off_prior'], off_mus,
                                                 x, y, resp, subiters)
        off_mus = offs[0]
        counts = np.sum(resp, axis=1)
        if gating_prior[0] == 'dirichlet':
            gpost = ('dirichlet', dirichlet_posterior(gating_prior[1], counts))
            gv = dirichlet_vlb(gating_prior[1], gpost[1])
        else:
            gpost = ('stick',) + tuple(stick_posterior(gating_prior[1], gating_prior[2], counts))
            gv = stick_vlb(gating_prior[1:], gpost[1:])
        omegas = bk[:, None, None] * basis['post_lmbdas']
        ell_b = hnw_expected_loglik(x, hq, bm, omegas)
        ell_m = tam_expected_loglik(x, y, slope, offs, prec)
        joint = ell_b + ell_m + hgmm_log_weights(gpost)[:, None]
        resp = responsibilities(joint)[0]
        out = dict(basis_mus=bm, basis_kappas=bk, hyper=hq, slope=slope, precision=prec, offsets=offs, gating=gpost,
                   resp=resp, ell=joint)
        if gpost[0] == 'dirichlet':
            lab = vlb_labels_dirichlet(resp, dirichlet_expected_log(gpost[1]))
        else:
            _, Es, Er = stick_expected_log(gpost[1], gpost[2])
            lab = vlb_labels_stick(resp, Es, Er)
        out['vlb'] = gv + hnw_vlb(basis['hyper_prior'], hq, basis['kappas0'], bm, omegas) \
            + np.sum(tam_vlb(models['slope_prior'], models['off_prior'], models['prec_prior'], slope, offs, prec)) \
            + lab + np.sum(resp * (ell_b + ell_m))
    return out


def chunked(fn, n, chunk):
    """Apply fn(slice) over point chunks and concatenate along the point axis
    (axis 1 for (K,N) outputs).  Exact: every per-point quantity is
    independent across points."""
    outs = [fn(slice(s, min(s + chunk, n))) for s in range(0, n, chunk)]
    return np.concatenate(outs, axis=1)
