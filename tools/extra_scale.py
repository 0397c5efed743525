"""Round-2 extras at scale on one B200 (not bench.py legs: BASELINE.json names neither): the hierarchical mixture of
mixtures/hgmm.py and device-resident SVI at the cfg4 shape (N = 10M, d = 16, K = 64).  Prints a small markdown report."""
import os
import random
import sys
import time

import numpy as np
import numpy.random as npr
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import mimo_b200                                                       # noqa: E402
from mimo_b200 import _engine as E                                     # noqa: E402
from mimo_b200.distributions import (Dirichlet, TruncatedStickBreaking, CategoricalWithDirichlet, CategoricalWithStickBreaking,  # noqa: E402
                                     NormalWishart, TiedGaussiansWithScaledPrecision, TiedGaussiansWithHierarchicalNormalWisharts,
                                     StackedNormalWisharts, StackedGaussiansWithNormalWisharts)
from mimo_b200.mixtures import BayesianMixtureOfGaussiansWithHierarchicalPrior, BayesianMixtureOfGaussians  # noqa: E402

N, d, K = 10_000_000, 16, 64
mimo_b200.set_default_precision('fp32')
g = torch.Generator(device='cuda')
g.manual_seed(1)
centres = 6. * torch.randn((K, d), generator=g, device='cuda')
lab = torch.randint(0, K, (N,), generator=g, device='cuda')
Z = (centres[lab] + torch.randn((N, d), generator=g, device='cuda')).float().contiguous()
del lab
torch.cuda.synchronize()
print('# Extras at the cfg4 shape (N = %d, d = %d, K = %d), one B200, FP32 compute / FP64 accumulation\n' % (N, d, K))

# ---- hierarchical mixture: mean field ---------------------------------------------------------------------------
npr.seed(0)
gating = CategoricalWithStickBreaking(K, TruncatedStickBreaking(K, np.ones(K), 5. * np.ones(K)))
hp = NormalWishart(dim=d, mu=np.zeros(d), kappa=1e-2, psi=np.eye(d), nu=d + 1 + 1e-8)
comp = TiedGaussiansWithHierarchicalNormalWisharts(K, d, hyper_prior=hp, prior=TiedGaussiansWithScaledPrecision(K, d, kappas=1e-2 * np.ones(K)))
model = BayesianMixtureOfGaussiansWithHierarchicalPrior(K, d, gating=gating, components=comp)
# start from the posterior of a k-means-like assignment: one sweep with the sampled likelihood is enough to leave the prior
model.components.posterior.mus = E.to_host(centres).astype(np.float64) + 0.5 * npr.randn(K, d)
model.components.posterior.kappas = np.full(K, N / K)
t = []
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vlb = model.meanfield_coordinate_descent(Z, randomize=False, maxiter=6, maxsubiter=5, tol=0., progress_bar=False)
    torch.cuda.synchronize()
    t.append((time.perf_counter() - t0) / 6)
print('## BayesianMixtureOfGaussiansWithHierarchicalPrior.meanfield_coordinate_descent (hgmm.py:186-215)\n')
print('%.1f ms per outer iteration (5 sub-iterations of the hierarchical prior on K (d + 1) + d (d + 1) / 2 reduced statistics, one fused'
      ' E-step + statistics sweep over the resident data): %.2e points*components/s; lower bound %.6e -> %.6e\n'
      % (1e3 * t[-1], N * K / t[-1], vlb[0], vlb[-1]))

# ---- SVI --------------------------------------------------------------------------------------------------------
def gmm():
    npr.seed(0)
    prior = StackedNormalWisharts(K, d, np.zeros((K, d)), 1e-2 * np.ones(K), np.stack(K * [np.eye(d)]), (d + 1.) * np.ones(K) + 1e-8)
    return BayesianMixtureOfGaussians(gating=CategoricalWithDirichlet(K, Dirichlet(K, np.ones(K))),
                                      components=StackedGaussiansWithNormalWisharts(K, d, prior=prior))

print('## meanfield_stochastic_descent at batch_size = 256, 200 iterations, full-data bound at the end only\n')
print('| route | ms per iteration |')
print('|---|---|')
host_x = None
for route, kw in (('API (minibatch kernels + host blend), 20 iterations', None), ('device-resident, eager', dict(device=True, graph=False)),
                  ('device-resident, CUDA graph', dict(device=True, graph=True)),
                  ('device-resident, CUDA graph, bound every iteration on the 10M points', dict(device=True, graph=True, lower_bound_every=1)),
                  ('device-resident, eager, bound every iteration on the 10M points', dict(device=True, graph=False, lower_bound_every=1))):
    m = gmm()
    random.seed(3)
    npr.seed(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if kw is None:
        if host_x is None:
            host_x = Z[:200_000].double().cpu().numpy()                 # the API route takes host arrays: a 200k-point subset
        it = 20
        m.meanfield_stochastic_descent(host_x, maxiter=it, step_size=5e-2, batch_size=256, progress_bar=False)
    else:
        it = 200 if kw.get('lower_bound_every', 0) != 1 else 30
        every = kw.pop('lower_bound_every', it)
        m.meanfield_stochastic_descent(Z, maxiter=it, step_size=5e-2, batch_size=256, progress_bar=False, lower_bound_every=every, **kw)
    torch.cuda.synchronize()
    print('| %s | %.2f |' % (route, 1e3 * (time.perf_counter() - t0) / it))
