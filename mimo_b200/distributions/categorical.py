"""The label distribution of a mixture: `dim` probabilities, counts as sufficient statistics.

API of mimo/distributions/categorical.py:6-68 (same names, arguments and results).  Counts of hard
labels are a bincount, soft counts are the row sums of a (K, N) responsibility matrix; lists of
shards are summed, which is what the data-sharded driver relies on (utils/abstraction.py:12-14).
Everything here is O(K) or O(N) host work on integers / one pass over the weights; the sweep itself
gets its counts from the last column of the packed statistics instead.
"""
import numpy as np
import numpy.random as npr


class Categorical:

    def __init__(self, dim, probs=None):
        self.dim = dim
        self.probs = np.full((dim,), 1. / dim) if probs is None else probs

    # -- parameters ---------------------------------------------------------------------------
    def _get_params(self):
        return self.probs

    def _set_params(self, values):
        self.probs = values

    params = property(_get_params, _set_params)
    nb_params = property(lambda self: len(self.probs) - 1)

    # -- statistics ---------------------------------------------------------------------------
    def statistics(self, data):
        """hard-label counts per component (categorical.py:35-39)."""
        if not isinstance(data, np.ndarray):
            return sum(self.statistics(shard) for shard in data)
        return np.bincount(data, minlength=self.dim)

    def weighted_statistics(self, data, weights):
        """soft counts: sum over points of the responsibilities (categorical.py:41-46); `data` is unused."""
        if not isinstance(weights, np.ndarray):
            return sum(self.weighted_statistics(None, w) for w in weights)
        return np.atleast_2d(weights).sum(axis=1)

    def max_likelihood(self, data, weights=None):
        """probs = normalised counts (categorical.py:65-68)."""
        counts = self.weighted_statistics(data, weights) if weights is not None else self.statistics(data)
        self.probs = counts / counts.sum()

    # -- density ------------------------------------------------------------------------------
    def log_likelihood(self, x):
        """log probs[x]; NaN entries give 0 (categorical.py:51-59)."""
        x = np.asarray(x)
        keep = ~np.isnan(x)
        out = np.zeros(x.shape, dtype=np.double)
        with np.errstate(invalid='ignore', divide='ignore'):
            table = np.log(self.probs)
        out[keep] = table[x[keep].astype(int)]
        return out

    def rvs(self, size=1):
        return npr.choice(a=self.dim, p=self.probs, size=size)

    def mode(self):
        return np.argmax(self.probs)

    # -- not defined by the reference either ---------------------------------------------------
    def mean(self):
        raise NotImplementedError

    def log_partition(self):
        raise NotImplementedError

    def entropy(self):
        raise NotImplementedError
