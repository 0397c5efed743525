"""Sampling from log-probabilities on the GPU (mirrors mimo/utils/stats.py:8-21)."""
import numpy as np
import numpy.random as npr

from .. import _engine as E


def sample_discrete_from_log(p_log, return_lognorms=False, axis=0, dtype=np.int32, precision='fp64'):
    """Inverse-CDF categorical draw along the component axis.  Consumes exactly the
    RNG call the reference makes -- ``npr.random(size=(1, N))`` -- so a seeded run
    sees the same uniforms; the draw itself runs in the softmax kernel."""
    p_log = np.asarray(p_log)
    if axis != 0 or p_log.ndim != 2:
        raise NotImplementedError('only (K, N) arrays sampled along axis 0 are supported')
    K, N = p_log.shape
    u = npr.random(size=(1, N))
    a = E.to_dev(p_log, E.tdtype(precision))
    out = E.softmax(a, precision, labels=True, lse=return_lognorms, uniforms=u[0])
    labels = E.to_host(out['labels']).astype(dtype)
    if return_lognorms:
        return labels, E.to_host(out['lse']).astype(np.float64)
    return labels
