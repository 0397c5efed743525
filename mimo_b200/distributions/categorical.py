"""Label distribution (mirrors mimo/distributions/categorical.py)."""
import numpy as np
import numpy.random as npr


class Categorical:

    def __init__(self, dim, probs=None):
        self.dim = dim
        self.probs = np.ones((dim,)) / dim if probs is None else probs

    @property
    def params(self):
        return self.probs

    @params.setter
    def params(self, values):
        self.probs = values

    @property
    def nb_params(self):
        return len(self.probs) - 1

    def mean(self):
        raise NotImplementedError

    def mode(self):
        return np.argmax(self.probs)

    def rvs(self, size=1):
        return npr.choice(a=self.dim, p=self.probs, size=size)

    def statistics(self, data):
        """label counts (categorical.py:35-39); lists of shards are summed."""
        if isinstance(data, np.ndarray):
            return np.bincount(data, minlength=self.dim)
        return sum(self.statistics(d) for d in data)

    def weighted_statistics(self, data, weights):
        """soft counts = row sums of the responsibilities (categorical.py:41-46)."""
        if isinstance(weights, np.ndarray):
            return np.sum(np.atleast_2d(weights), axis=1)
        return sum(self.weighted_statistics(None, w) for w in weights)

    def log_partition(self):
        raise NotImplementedError

    def log_likelihood(self, x):
        x = np.asarray(x)
        out = np.zeros(x.shape, dtype=np.double)
        with np.errstate(invalid='ignore', divide='ignore'):
            good = ~np.isnan(x)
            out[good] = np.log(self.probs)[x[good].astype(int)]
        return out

    def entropy(self):
        raise NotImplementedError

    def max_likelihood(self, data, weights=None):
        counts = self.statistics(data) if weights is None else self.weighted_statistics(data, weights)
        self.probs = counts / counts.sum()
