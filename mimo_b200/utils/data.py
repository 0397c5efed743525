"""Label / minibatch helpers (mirrors mimo/utils/data.py:9-12, 160-169)."""
import random

import numpy as np


def one_hot(z, K):
    """labels -> dense (K, N) float64 indicator (kept for API parity; the sweep
    drivers never build it: labels go straight to the hard-statistics kernel)."""
    z = np.atleast_1d(z).astype(int)
    assert np.all(z >= 0) and np.all(z < K)
    out = np.zeros((K, z.size))
    out[np.ravel(z), np.arange(z.size)] = 1
    return np.reshape(out, (K,) + z.shape)


def batches(batch_size, data_size):
    """One minibatch of `batch_size` distinct indices per call (the reference draws
    batch_size indices and yields them as a single batch: SURVEY q6)."""
    yield random.sample(range(data_size), batch_size)
