"""Device plumbing between the NumPy-facing API and the C-ABI kernels.

torch is used for device memory, streams and (in sharded.py) torch.distributed
only; every numerical operation below is a call into libmimo_b200.so.
"""
import numpy as np
import torch

from . import _lib
from ._lib import F32, F64, WRITE_RESP, WRITE_LSE, DRAW_LABELS, ACC_LSE  # noqa: F401

_DEFAULT_PRECISION = 'fp32'


def set_default_precision(precision):
    """'fp32' (FP32 compute, FP64 accumulation; rel 1e-4 parity) or 'fp64' (1e-9 parity)."""
    global _DEFAULT_PRECISION
    assert precision in ('fp32', 'fp64')
    _DEFAULT_PRECISION = precision


def default_precision():
    return _DEFAULT_PRECISION


def tdtype(precision):
    return torch.float32 if precision == 'fp32' else torch.float64


def code(precision):
    return F32 if precision == 'fp32' else F64


def device():
    _lib.require_device()
    return torch.device('cuda', torch.cuda.current_device())


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def ld(t):
    """leading dimension of a (K, N) tensor for the C-ABI; a single row may carry any stride (torch normalises the
    strides of size-1 dimensions when it copies), the kernels want ld >= N."""
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def to_dev(x, dtype=torch.float64):
    if isinstance(x, torch.Tensor):
        return x.to(device=device(), dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x))).to(device=device(), dtype=dtype)


def to_host(t):
    return t.detach().cpu().numpy()


def zeros(shape, dtype=torch.float64):
    return torch.zeros(shape, dtype=dtype, device=device())


def empty(shape, dtype=torch.float64):
    return torch.empty(shape, dtype=dtype, device=device())


def pad_rows(r):
    p = 8
    while p < r:
        p *= 2
    if p > 128:
        raise NotImplementedError('mimo_b200: %d whitening rows per component (> 128) is not supported' % r)
    return p


def pad_cols(dp):
    return (dp + 3) // 4 * 4


# ---- feature tables ---------------------------------------------------------------
_feat_cache = {}


class Features:
    """Index pairs (fi, fj) into zt = [z ; 1] defining the packed statistics."""

    def __init__(self, fi, fj, D):
        self.D = D
        self.F = len(fi)
        self.fi_host = np.asarray(fi, dtype=np.int32)
        self.fj_host = np.asarray(fj, dtype=np.int32)
        self._dev = {}                                   # device index -> (fi, fj) tensors
        # canonical packed triangle f = i (i + 1) / 2 + j?  (the device copy is made from these host arrays, so the sweep can
        # be told instead of reading the tables back)
        il = np.tril_indices(D + 1)
        self.canonical = bool(self.F == (D + 1) * (D + 2) // 2 and np.array_equal(self.fi_host, il[0])
                              and np.array_equal(self.fj_host, il[1]))

    def dev(self):
        idx = device().index
        if idx not in self._dev:
            self._dev[idx] = (to_dev(self.fi_host, torch.int32), to_dev(self.fj_host, torch.int32))
        return self._dev[idx]


def quad_features(D):
    """all pairs j <= i of zt (length D+1), f = i(i+1)/2 + j."""
    key = ('quad', D)
    if key not in _feat_cache:
        fi, fj = [], []
        for i in range(D + 1):
            for j in range(i + 1):
                fi.append(i)
                fj.append(j)
        _feat_cache[key] = Features(fi, fj, D)
    return _feat_cache[key]


def diag_features(D):
    """[z_j * 1 | z_j * z_j | 1 * 1]."""
    key = ('diag', D)
    if key not in _feat_cache:
        fi = list(range(D)) + list(range(D)) + [D]
        fj = [D] * D + list(range(D)) + [D]
        _feat_cache[key] = Features(fi, fj, D)
    return _feat_cache[key]


def tri(a, b):
    a, b = (a, b) if a >= b else (b, a)
    return a * (a + 1) // 2 + b


# ---- operands -----------------------------------------------------------------------
class QuadOperands:
    """W (K, Rp, Dpp), cst (K): a[k][n] = cst[k] - 0.5 || W_k [z_n ; 1] ||^2."""
    family = 0

    def __init__(self, K, D, rows, precision):
        self.K, self.D, self.rows = K, D, rows
        self.Rp, self.Dpp = pad_rows(rows), pad_cols(D + 1)
        self.precision = precision
        self.W = zeros((K, self.Rp, self.Dpp), tdtype(precision))
        self.cst = zeros((K,), tdtype(precision))

    def args(self):
        return ptr(self.W), None, ptr(self.cst), self.K, self.Rp, self.Dpp


class DiagOperands:
    """S, T (K, D), cst (K): a[k][n] = cst[k] - 0.5 sum_j (S_kj z_nj - T_kj)^2."""
    family = 1

    def __init__(self, K, D, precision):
        self.K, self.D = K, D
        self.Rp, self.Dpp = 8, pad_cols(D + 1)
        self.precision = precision
        self.S = zeros((K, D), tdtype(precision))
        self.T = zeros((K, D), tdtype(precision))
        self.cst = zeros((K,), tdtype(precision))

    def args(self):
        return ptr(self.S), ptr(self.T), ptr(self.cst), self.K, self.Rp, self.Dpp


class Info:
    """Device-side status word of the posterior kernels: {code, failing component}."""

    def __init__(self):
        self.t = zeros((2,), torch.int32)

    def check(self):
        h = self.t.cpu().numpy()
        if h[0] != 0:
            self.t.zero_()
            if h[0] == _lib.ENOTPD:
                raise np.linalg.LinAlgError('Matrix is not positive definite (component %d)' % h[1])
            raise AssertionError('posterior kernel rejected its input (code %d, component %d)' % (h[0], h[1]))


def workspace(nbytes):
    return torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=device())


# ---- per-point kernels ----------------------------------------------------------------
def loglik(Z, ops, out=None):
    """(K, N) log-likelihood block for resident data Z (N, D)."""
    N, D = Z.shape
    assert D == ops.D and Z.dtype == tdtype(ops.precision)
    if out is None:
        out = empty((ops.K, N), Z.dtype)
    if ops.family == 0:
        _lib.call('mimo_loglik_quad', code(ops.precision), ptr(Z), N, D, Z.stride(0), ptr(ops.W), ptr(ops.cst),
                  ops.K, ops.Rp, ops.Dpp, ptr(out), out.stride(0), stream())
    else:
        _lib.call('mimo_loglik_diag', code(ops.precision), ptr(Z), N, D, Z.stride(0), ptr(ops.S), ptr(ops.T),
                  ptr(ops.cst), ops.K, ptr(out), out.stride(0), stream())
    return out


def loglik_diag_tc(Z, ops, out=False, labels=False, lse=False, lse_sum=False, uniforms=None, seed=0, offset=0):
    """Diagonal family on the tensor pipe (FP32, D <= 64, K <= 256): log-joints, labels and log-normalisers of the points
    of Z in one kernel.  Returns a dict; 'guard' = 1 means the operands failed the cancellation guard and nothing was
    computed (a sweep then takes the CUDA-core kernels)."""
    N, D = Z.shape
    assert ops.family == 1 and D == ops.D and Z.dtype == torch.float32 and ops.precision == 'fp32'
    out_t = empty((ops.K, N), torch.float32) if out else None
    lab_t = empty((N,), torch.int32) if labels else None
    lse_t = empty((N,), torch.float32) if lse else None
    sum_t = zeros((1,), torch.float64) if lse_sum else None
    uni_t = to_dev(uniforms, torch.float64).reshape(-1) if uniforms is not None else None
    wsb = _lib.load().mimo_loglik_diag_tc_workspace()
    ws = workspace(wsb)
    guard = np.zeros(1, dtype=np.uint32)
    _lib.call('mimo_loglik_diag_tc', ptr(Z), N, D, Z.stride(0), ptr(ops.S), ptr(ops.T), ptr(ops.cst), ops.K,
              ptr(out_t), out_t.stride(0) if out else 0, ptr(lab_t), ptr(uni_t), int(seed), int(offset),
              ptr(lse_t), ptr(sum_t), guard.ctypes.data, ptr(ws), wsb, stream())
    return dict(out=out_t, labels=lab_t, lse=lse_t, lse_sum=sum_t, guard=int(guard[0]))


def studentt_from_quad(a, precision, c0, add, df):
    """In place: a[k][n] <- add[k] + log1p(2 (c0[k] - a[k][n]) / df[k])   (utils/stats.py:53-79 from a Gaussian-form log-joint)."""
    K, n = a.shape
    dev = [to_dev(np.ascontiguousarray(np.broadcast_to(t, (K,)), dtype=np.float64)) for t in (c0, add, df)]   # kept alive over the call
    _lib.call('mimo_studentt_from_quad', code(precision), ptr(a), K, n, a.stride(0), ptr(dev[0]), ptr(dev[1]), ptr(dev[2]), stream())
    return a


def predict_lingauss(X, W, M, Kinv, Sigma, Psi, logdet_psi, df, affine, mode, studentt, precision, Y=None, eps=0.0):
    """Posterior-predictive moments of K linear-Gaussian experts combined with the weights W (K, N): (mu (N, o),
    cov (N, o, o), nlpd (N) or None), all on the device (mixtures/ilr.py:352-411)."""
    N, din = X.shape
    K, o, c = M.shape
    dt = tdtype(precision)
    assert X.dtype == dt and W.dtype == dt and W.shape == (K, N) and c == din + (1 if affine else 0)
    tied = 0                                   # the mirror classes hand over stacked (K, ...) parameters also for tied experts
    dev = [to_dev(np.ascontiguousarray(t, dtype=np.float64)) for t in (M, Kinv, Sigma, Psi, np.atleast_1d(logdet_psi), np.atleast_1d(df))]
    mu, cov = empty((N, o), dt), empty((N, o, o), dt)
    nlpd = empty((N,), dt) if Y is not None else None
    _lib.call('mimo_predict_lingauss', code(precision), ptr(X), N, X.stride(0), din, int(bool(affine)), ptr(W), W.stride(0), K,
              ptr(dev[0]), ptr(dev[1]), ptr(dev[2]), ptr(dev[3]), ptr(dev[4]), ptr(dev[5]), o, tied, int(mode), int(bool(studentt)),
              ptr(Y), Y.stride(0) if Y is not None else 0, float(eps), ptr(mu), ptr(cov), ptr(nlpd), stream())
    return mu, cov, nlpd


def softmax(a, precision, resp=False, lse=False, labels=False, lse_sum=False, uniforms=None, seed=0, offset=0):
    """In-place softmax / label draw over a (K, n) log-joint.  Returns a dict."""
    K, n = a.shape
    flags = (WRITE_RESP if resp else 0) | (WRITE_LSE if lse else 0) | (DRAW_LABELS if labels else 0) \
        | (ACC_LSE if lse_sum else 0)
    out = {}
    lse_t = empty((n,), a.dtype) if lse else None
    lab_t = empty((n,), torch.int32) if labels else None
    sum_t = zeros((1,), torch.float64) if lse_sum else None
    uni_t = to_dev(uniforms, torch.float64).reshape(-1) if uniforms is not None else None
    if uni_t is not None:
        assert uni_t.numel() == n
    _lib.call('mimo_softmax', code(precision), ptr(a), K, n, a.stride(0), flags, ptr(lse_t), ptr(uni_t),
              int(seed), int(offset), ptr(lab_t), ptr(sum_t), stream())
    out['lse'], out['labels'], out['lse_sum'] = lse_t, lab_t, sum_t
    return out


def stats_soft(Z, resp, feats, precision, stat=None):
    N, D = Z.shape
    K = resp.shape[0]
    fi, fj = feats.dev()
    if stat is None:
        stat = zeros((K, feats.F), torch.float64)
    if N == 0:
        return stat
    _lib.call('mimo_stats_soft', code(precision), ptr(Z), N, D, Z.stride(0), ptr(resp), ld(resp), K,
              ptr(fi), ptr(fj), feats.F, ptr(stat), stream())
    return stat


def stats_hard(Z, labels, K, feats, precision, stat=None):
    N, D = Z.shape
    fi, fj = feats.dev()
    if stat is None:
        stat = zeros((K, feats.F), torch.float64)
    if N == 0:
        return stat
    wsb = _lib.load().mimo_stats_hard_workspace(N, K)
    ws = workspace(wsb)
    _lib.call('mimo_stats_hard', code(precision), ptr(Z), N, D, Z.stride(0), ptr(labels), K,
              ptr(fi), ptr(fj), feats.F, ptr(stat), ptr(ws), wsb, stream())
    return stat


class SweepBuffers:
    """Reusable device buffers for mimo_sweep on a fixed (N, K, F)."""

    def __init__(self, N, K, F, precision, hard):
        self.N, self.K, self.F, self.precision, self.hard = N, K, F, precision, hard
        self.wsb, self.ws = 0, None
        # statistics and the lower-bound scalar share one buffer: ONE all-reduce message per sweep
        self.flat = zeros((K * F + 1,), torch.float64)
        self.stat = self.flat[:K * F].view(K, F)
        self.lse_sum = self.flat[K * F:]
        self.labels = empty((N,), torch.int32) if hard else None

    def ensure(self, ops, D):
        """workspace for a sweep with these operands: the operand family / D / Rp decide whether
        the tensor-core kernels (and their operand image / partial-statistics buffers) are used."""
        need = _lib.load().mimo_sweep_workspace(code(self.precision), ops.family, 1 if self.hard else 0,
                                                self.N, D, self.K, ops.Rp)
        if self.ws is None or need > self.wsb:
            self.ws = None
            self.ws = workspace(need)
            self.wsb = need


def sweep(Z, ops, feats, buf, uniforms=None, seed=0, offset=0, ll_out=None, lse_out=None, zero=True,
          phase_ms=None, absmax=None):
    """One E-step + statistics pass over resident Z.  Results land in buf.stat,
    buf.lse_sum and (hard) buf.labels.  phase_ms: optional float64 numpy array (6,) that
    accumulates per-phase device milliseconds, the launch count and the chunk count (synchronises)."""
    N, D = Z.shape
    assert Z.is_cuda and Z.dtype == tdtype(ops.precision) and D == ops.D, 'data must be a resident %s tensor of width %d' % (ops.precision, ops.D)
    fi, fj = feats.dev()
    buf.ensure(ops, D)
    if zero:
        buf.stat.zero_()
        buf.lse_sum.zero_()
    if N == 0:                      # an empty shard: nothing to add (an empty tensor has no device pointer to pass)
        return buf
    a, b, c, K, Rp, Dpp = ops.args()
    if absmax is not None and absmax > 0:      # max |Z| of resident, unchanged data: the sweep skips its own pass over Z for it
        _lib.call('mimo_sweep_absmax_hint', float(absmax))
    if getattr(feats, 'canonical', None) is not None:      # tables this module built itself: no read-back inside the sweep
        _lib.call('mimo_sweep_tables_hint', 1 if feats.canonical else 0)
    args = (code(ops.precision), ops.family, 1 if buf.hard else 0,
            ptr(Z), N, D, Z.stride(0), a, b, c, K, Rp, Dpp, ptr(fi), ptr(fj), feats.F,
            ptr(uniforms), int(seed), int(offset), ptr(buf.stat), ptr(buf.lse_sum), ptr(buf.labels),
            ptr(lse_out), ptr(ll_out), (ll_out.stride(0) if ll_out is not None else 0),
            ptr(buf.ws), buf.wsb, stream())
    if phase_ms is None:
        _lib.call('mimo_sweep', *args)
    else:
        assert phase_ms.dtype == np.float64 and phase_ms.size >= 6
        _lib.call('mimo_sweep_timed', *args, phase_ms.ctypes.data)
    return buf


# ---- per-component kernels --------------------------------------------------------------
def _i32(x):
    return to_dev(np.asarray(x, dtype=np.int32), torch.int32)


def identity_map(d, D):
    """variable j -> column j, constant -> column D (a plain Gaussian over all of z)."""
    return _i32(list(range(d)) + [D])


def nw_posterior(prior, stat, F, stat_idx, Dp, mode=0, tied=False, variates=None, ops=None, row_off=0,
                 col_map=None, want_lik=False, want_vlb=True, info=None, k_range=None):
    """prior = (m0 (K,d), kappa0 (K), psi0 (K,d,d), nu0 (K)) device FP64 tensors.
    k_range = (k0, k1): update only the components [k0, k1) -- the outputs keep their full (K, ...) shapes and only
    that slice (and that slice of the operand block in `ops`) is written; the sharded driver gathers the slices of
    all ranks.  The kernels are one CTA per component, so a slice is bit-identical to the same rows of a full call."""
    m0, k0, p0, n0 = prior
    K, d = m0.shape
    out = dict(m=empty((K, d)), kappa=empty((K,)), psi=empty((K, d, d)), nu=empty((K,)))
    out['vlb'] = empty((K,)) if want_vlb else None
    out['lik_mu'] = empty((K, d)) if want_lik else None
    out['lik_lmbda'] = empty((K, d, d)) if want_lik else None
    lo, hi = (0, K) if k_range is None else k_range
    assert 0 <= lo < hi <= K and (k_range is None or not tied), 'a component slice needs untied components'
    sl = slice(lo, hi)
    Kl = hi - lo

    def part(t):
        return None if t is None else t[sl]
    wsb = _lib.load().mimo_nw_workspace(Kl, d)
    ws = workspace(wsb)
    info = info or Info()
    var_t = to_dev(variates) if variates is not None else None
    W = ops.W if ops is not None else None
    _lib.call('mimo_nw_posterior', Kl, d, int(tied), mode, ptr(m0[sl]), ptr(k0[sl]), ptr(p0[sl]), ptr(n0[sl]),
              ptr(stat[sl]), F, ptr(stat_idx), Dp, ptr(part(var_t)),
              ptr(out['m'][sl]), ptr(out['kappa'][sl]), ptr(out['psi'][sl]), ptr(out['nu'][sl]),
              ptr(part(out['lik_mu'])), ptr(part(out['lik_lmbda'])), ptr(part(out['vlb'])),
              code(ops.precision) if ops is not None else F64, ptr(part(W)), ptr(ops.cst[sl]) if ops is not None else None,
              ops.Rp if ops is not None else 8, ops.Dpp if ops is not None else pad_cols(Dp), row_off, ptr(col_map),
              ptr(ws), wsb, ptr(info.t), stream())
    out['info'] = info
    out['k_range'] = (lo, hi)
    return out


def ng_posterior(prior, stat, F, mode=0, tied=False, bug_compat=False, variates=None, ops=None,
                 want_lik=False, want_vlb=True, info=None):
    m0, k0, a0, b0 = prior
    K, d = m0.shape
    out = dict(m=empty((K, d)), kappa=empty((K, d)), alpha=empty((K, d)), beta=empty((K, d)))
    out['vlb'] = empty((K,)) if want_vlb else None
    out['lik_mu'] = empty((K, d)) if want_lik else None
    out['lik_lmbda'] = empty((K, d)) if want_lik else None
    wsb = _lib.load().mimo_ng_workspace(K, d)
    ws = workspace(wsb)
    info = info or Info()
    var_t = to_dev(variates) if variates is not None else None
    _lib.call('mimo_ng_posterior', K, d, int(tied), mode, int(bug_compat), ptr(m0), ptr(k0), ptr(a0), ptr(b0),
              ptr(stat), F, ptr(var_t), ptr(out['m']), ptr(out['kappa']), ptr(out['alpha']), ptr(out['beta']),
              ptr(out['lik_mu']), ptr(out['lik_lmbda']), ptr(out['vlb']),
              code(ops.precision) if ops is not None else F64,
              ptr(ops.S) if ops is not None else None, ptr(ops.T) if ops is not None else None,
              ptr(ops.cst) if ops is not None else None, ptr(ws), wsb, ptr(info.t), stream())
    out['info'] = info
    return out


def mnw_posterior(prior, stat, F, stat_idx, Dp, mode=0, tied=False, variates=None, ops=None, row_off=0,
                  col_map=None, want_lik=False, want_vlb=True, info=None):
    """prior = (M0 (K,o,c), K0 (K,c,c), psi0 (K,o,o), nu0 (K))."""
    M0, K0, p0, n0 = prior
    K, o, c = M0.shape
    out = dict(M=empty((K, o, c)), K=empty((K, c, c)), psi=empty((K, o, o)), nu=empty((K,)))
    out['vlb'] = empty((K,)) if want_vlb else None
    out['lik_A'] = empty((K, o, c)) if want_lik else None
    out['lik_lmbda'] = empty((K, o, o)) if want_lik else None
    wsb = _lib.load().mimo_mnw_workspace(K, c, o)
    ws = workspace(wsb)
    info = info or Info()
    var_t = to_dev(variates) if variates is not None else None
    _lib.call('mimo_mnw_posterior', K, c, o, int(tied), mode, ptr(M0), ptr(K0), ptr(p0), ptr(n0),
              ptr(stat), F, ptr(stat_idx), Dp, ptr(var_t),
              ptr(out['M']), ptr(out['K']), ptr(out['psi']), ptr(out['nu']),
              ptr(out['lik_A']), ptr(out['lik_lmbda']), ptr(out['vlb']),
              code(ops.precision) if ops is not None else F64, ptr(ops.W) if ops is not None else None,
              ptr(ops.cst) if ops is not None else None,
              ops.Rp if ops is not None else 8, ops.Dpp if ops is not None else pad_cols(Dp), row_off, ptr(col_map),
              ptr(ws), wsb, ptr(info.t), stream())
    out['info'] = info
    return out


def gating_posterior(kind, prior_a, prior_b, stat, F, count_feature, mode=0, variates=None, ops=None, info=None):
    """kind 0 Dirichlet / 1 stick-breaking.  Sets ops.cst to the log-weights."""
    K = prior_a.shape[0]
    out = dict(a=empty((K,)), b=empty((K,)) if kind == 1 else None, probs=empty((K,)), vlb=empty((1,)))
    wsb = _lib.load().mimo_gating_workspace(K)
    ws = workspace(wsb)
    info = info or Info()
    var_t = to_dev(variates) if variates is not None else None
    _lib.call('mimo_gating_posterior', K, kind, mode, ptr(prior_a), ptr(prior_b), ptr(stat), F, count_feature,
              ptr(var_t), ptr(out['a']), ptr(out['b']), ptr(out['probs']), ptr(out['vlb']),
              code(ops.precision) if ops is not None else F64, ptr(ops.cst) if ops is not None else None,
              ptr(ws), wsb, ptr(info.t), stream())
    out['info'] = info
    return out


def set_log_weights(ops, logw):
    """cst[k] <- log-weight (explicit gating probabilities; categorical.py:51-59)."""
    ops.cst.copy_(to_dev(logw, ops.cst.dtype))


def operands_gauss(ops, mu, lmbda, row_off=0, col_map=None, info=None):
    K, d = mu.shape
    wsb = _lib.load().mimo_operands_workspace(K, d)
    ws = workspace(wsb)
    info = info or Info()
    col_map = col_map if col_map is not None else identity_map(d, ops.D)
    _lib.call('mimo_operands_gauss', K, d, ptr(mu), ptr(lmbda), code(ops.precision), ptr(ops.W), ptr(ops.cst),
              ops.Rp, ops.Dpp, row_off, ptr(col_map), ptr(ws), wsb, ptr(info.t), stream())
    return info


def operands_lingauss(ops, A, lmbda, row_off, col_map, info=None):
    K, o, c = A.shape
    wsb = _lib.load().mimo_operands_workspace(K, o)
    ws = workspace(wsb)
    info = info or Info()
    _lib.call('mimo_operands_lingauss', K, c, o, ptr(A), ptr(lmbda), code(ops.precision), ptr(ops.W), ptr(ops.cst),
              ops.Rp, ops.Dpp, row_off, ptr(col_map), ptr(ws), wsb, ptr(info.t), stream())
    return info


def operands_gauss_diag(ops, mu, lam):
    K, d = mu.shape
    _lib.call('mimo_operands_gauss_diag', K, d, ptr(mu), ptr(lam), code(ops.precision), ptr(ops.S), ptr(ops.T),
              ptr(ops.cst), stream())


def mstep_gauss(stat, F, stat_idx, Dp, K, d, tied=False, info=None):
    mu, lmbda = empty((K, d)), empty((K, d, d))
    wsb = _lib.load().mimo_mstep_workspace(K, d)
    ws = workspace(wsb)
    info = info or Info()
    _lib.call('mimo_mstep_gauss', K, d, int(tied), ptr(stat), F, ptr(stat_idx), Dp, ptr(mu), ptr(lmbda),
              ptr(ws), wsb, ptr(info.t), stream())
    return mu, lmbda, info


def mstep_gauss_diag(stat, F, K, d, tied=False):
    mu, lam = empty((K, d)), empty((K, d))
    ws = workspace(8 * (d + 1))
    _lib.call('mimo_mstep_gauss_diag', K, d, int(tied), ptr(stat), F, ptr(mu), ptr(lam), ptr(ws), ws.numel(), stream())
    return mu, lam


def mstep_lingauss(stat, F, stat_idx, Dp, K, c, o, tied=False, info=None):
    A, lmbda = empty((K, o, c)), empty((K, o, o))
    wsb = _lib.load().mimo_mstep_lingauss_workspace(K, c, o)
    ws = workspace(wsb)
    info = info or Info()
    _lib.call('mimo_mstep_lingauss', K, c, o, int(tied), ptr(stat), F, ptr(stat_idx), Dp, ptr(A), ptr(lmbda),
              ptr(ws), wsb, ptr(info.t), stream())
    return A, lmbda, info


# ---- tensor-core variants (stand-alone entry points; mimo_sweep dispatches by itself) ---------------
def set_tensor_cores(mode):
    """0: CUDA-core kernels only; 1 (default): tcgen05 kernels on CTA pairs with the screened E-step;
    2: single-CTA dense tcgen05 kernels; 3: CTA pairs, dense E-step; 4: as 1 with dense statistics;
    5: as 1 with the screening starting on its all-rows tier.  Returns the old mode."""
    return _lib.load().mimo_set_tensor_cores(int(mode))


def set_tc_min_dim(d):
    """Smallest dimension a quad-family sweep takes to the tensor pipe (default 8).  Returns the old value."""
    return _lib.load().mimo_tc_set_min_dim(int(d))


def set_quad_generations(on):
    """Dense E-step for 64 < D <= 128: True = tc_estep4.cu (four components per accumulator generation, zero block of the
    Cholesky factors skipped; measured slower), False (default) = tc_estep2.cu.  Returns the old setting."""
    return _lib.load().mimo_tc_set_quad_generations(int(bool(on)))


def set_triangular(rows):
    """Dense E-step for 64 < D <= 128: rows per step of the triangular skip of tc_estep3.cu (16 or 32; 0 = the kernel
    with both operands in shared memory).  Returns the old value."""
    return _lib.load().mimo_tc_set_triangular(int(rows))


def screen_last():
    """(candidate pairs, dense-fallback flag) of the most recent screened point chunk."""
    out = np.zeros(2, dtype=np.uint32)
    _lib.call('mimo_tc_screen_last', out.ctypes.data)
    return int(out[0]), int(out[1])


def screen_totals():
    """totals of the most recent screened sweep over ALL its point chunks: (candidate pairs of the refined chunks,
    points of the refined chunks, chunks that took the dense pass, chunks, tier the sweep ended on)."""
    out = np.zeros(5, dtype=np.uint64)
    _lib.call('mimo_tc_screen_totals', out.ctypes.data)
    return tuple(int(v) for v in out)


def screen_level():
    """screening tier the most recent screened sweep ended on (0 projection, 1 all rows, 2 none; -1 unknown)."""
    return int(_lib.load().mimo_tc_screen_level())


def sweep_uses_tensor_cores(ops, D):
    return bool(_lib.load().mimo_sweep_uses_tensor_cores(code(ops.precision), ops.family, D, ops.Rp))


def loglik_tc(Z, ops, out=None):
    N, D = Z.shape
    assert ops.family == 0 and ops.precision == 'fp32' and Z.dtype == torch.float32
    if out is None:
        out = empty((ops.K, N), Z.dtype)
    wsb = _lib.load().mimo_loglik_quad_tc_workspace(ops.K, ops.Rp, D)
    ws = workspace(wsb)
    _lib.call('mimo_loglik_quad_tc', ptr(Z), N, D, Z.stride(0), ptr(ops.W), ptr(ops.cst), ops.K, ops.Rp, ops.Dpp,
              ptr(out), out.stride(0), ptr(ws), wsb, stream())
    return out


def stats_soft_tc(Z, resp, feats, stat=None):
    N, D = Z.shape
    K = resp.shape[0]
    assert Z.dtype == torch.float32 and resp.dtype == torch.float32
    if stat is None:
        stat = zeros((K, feats.F), torch.float64)
    wsb = _lib.load().mimo_stats_soft_tc_workspace(N, K)
    ws = workspace(wsb)
    _lib.call('mimo_stats_soft_tc', ptr(Z), N, D, Z.stride(0), ptr(resp), ld(resp), K, feats.F,
              ptr(stat), ptr(ws), wsb, stream())
    return stat
