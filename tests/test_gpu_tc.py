"""Parity of the tensor-core (tcgen05) kernels with the CPU oracle and with the CUDA-core
FP32 kernels (GPU only).  Same tolerance as the FP32 mode: rel 1e-4 against the oracle
(north_star); additionally the 3xFP16-split results must sit within a small multiple of
the FP32 kernels' own error.
"""
import numpy as np
import pytest
import torch

from oracle import mimo_oracle as orc
from test_gpu_kernels import close, eng, spd, unpack_quad

pytestmark = pytest.mark.gpu


def scaled_err(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.max(np.abs(np.asarray(a, float) - b)) / max(1.0, float(np.max(np.abs(b)))))


@pytest.mark.parametrize('K,d,N,shift', [(5, 128, 300, 0.0), (9, 16, 1000, 1.0), (3, 64, 700, 0.0), (70, 100, 513, 3.0),
                                         (33, 128, 2100, 5.0), (4, 2, 500, 0.0), (2, 65, 257, 0.0)])
def test_loglik_tc(K, d, N, shift):
    E = eng()
    rng = np.random.default_rng(d + K)
    x = rng.standard_normal((N, d)) * 2 + rng.standard_normal(d) + shift
    mus = rng.standard_normal((K, d)) * 2 + shift
    lmbdas = np.stack([spd(rng, d) for _ in range(K)])
    logw = np.log(rng.dirichlet(np.ones(K)))
    ops = E.QuadOperands(K, d, d, 'fp32')
    E.set_log_weights(ops, logw)
    E.operands_gauss(ops, E.to_dev(mus), E.to_dev(lmbdas)).check()
    Z = E.to_dev(x, torch.float32)
    ll_tc = E.loglik_tc(Z, ops)                      # default: CTA-pair kernel (cta_group::2)
    old = E.set_tensor_cores(2)                      # single-CTA kernel
    try:
        ll_tc1 = E.loglik_tc(Z, ops)
    finally:
        E.set_tensor_cores(old)
    ll_cc = E.loglik(Z, ops)
    xr = Z.double().cpu().numpy()
    ref = orc.gauss_full_loglik(xr, mus, lmbdas) + logw[:, None]
    close(ll_tc, ref, 1e-4, 'tensor-core log-lik (CTA pairs)')
    close(ll_tc1, ref, 1e-4, 'tensor-core log-lik (single CTA)')
    close(ll_tc, ll_tc1.cpu().numpy(), 1e-6, 'CTA-pair vs single-CTA kernel')
    e_tc, e_cc = scaled_err(ll_tc, ref), scaled_err(ll_cc, ref)
    print('loglik K=%d d=%d N=%d: scaled err tensor-core %.2e, CUDA-core fp32 %.2e' % (K, d, N, e_tc, e_cc))
    assert e_tc <= max(4 * e_cc, 2e-6)


@pytest.mark.parametrize('K,d,rows', [(5, 128, 128), (7, 70, 70), (6, 40, 64), (9, 33, 17)])
def test_loglik_tc_dense_operands(K, d, rows):
    """operand rows that are NOT triangular (what the ILR stack of basis + expert rows looks like):
    the issuer must fall back to full-width MMAs."""
    E = eng()
    rng = np.random.default_rng(5)
    N = 900
    ops = E.QuadOperands(K, d, rows, 'fp32')
    W = np.zeros((K, ops.Rp, ops.Dpp))
    W[:, :rows, :d + 1] = rng.standard_normal((K, rows, d + 1)) / np.sqrt(d)
    cst = rng.standard_normal(K)
    ops.W.copy_(E.to_dev(W, torch.float32))
    ops.cst.copy_(E.to_dev(cst, torch.float32))
    x = rng.standard_normal((N, d)) * 1.5 + 0.5
    Z = E.to_dev(x, torch.float32)
    ll = E.loglik_tc(Z, ops)
    Wr = ops.W.double().cpu().numpy()
    zt = np.concatenate([Z.double().cpu().numpy(), np.ones((N, 1))], axis=1)
    y = np.einsum('kij,nj->kni', Wr[:, :, :d + 1], zt)
    ref = ops.cst.double().cpu().numpy()[:, None] - 0.5 * np.sum(y * y, axis=2)
    close(ll, ref, 1e-4, 'dense-operand tensor-core log-lik')
    close(ll, E.loglik(Z, ops).cpu().numpy(), 2e-5, 'tensor-core vs CUDA-core log-lik')


@pytest.mark.parametrize('quad', [True, False])
@pytest.mark.parametrize('K,d,tri', [(9, 128, True), (4, 100, True), (5, 128, False), (3, 70, False), (1, 65, True), (1030, 128, True)])
def test_loglik_tc_dense_kernels_d_above_64(quad, K, d, tri):
    """the two dense E-step kernels for 64 < d <= 128: tc_estep2.cu (default) and tc_estep4.cu (four components per
    accumulator generation, the zero block of Cholesky-factor operands skipped; dense operands take the fourth step)."""
    E = eng()
    rng = np.random.default_rng(23)
    N = 1300 if K < 100 else 700
    ops = E.QuadOperands(K, d, d, 'fp32')
    W = np.zeros((K, ops.Rp, ops.Dpp))
    blk = rng.standard_normal((K, d, d)) / np.sqrt(d)
    W[:, :d, :d] = np.triu(blk) if tri else blk
    W[:, :d, d] = rng.standard_normal((K, d))
    cst = rng.standard_normal(K)
    ops.W.copy_(E.to_dev(W, torch.float32))
    ops.cst.copy_(E.to_dev(cst, torch.float32))
    Z = E.to_dev(rng.standard_normal((N, d)) * 1.5 + 0.5, torch.float32)
    old = E.set_quad_generations(quad)
    try:
        ll = E.loglik_tc(Z, ops)
    finally:
        E.set_quad_generations(old)
    Wr = ops.W.double().cpu().numpy()
    zt = np.concatenate([Z.double().cpu().numpy(), np.ones((N, 1))], axis=1)
    y = np.einsum('kij,nj->kni', Wr[:, :, :d + 1], zt)
    ref = ops.cst.double().cpu().numpy()[:, None] - 0.5 * np.sum(y * y, axis=2)
    close(ll, ref, 1e-4, 'dense log-lik (%s)' % ('tc_estep4' if quad else 'tc_estep2'))


@pytest.mark.parametrize('gran', [16, 32])
@pytest.mark.parametrize('K,d,tri', [(9, 128, True), (4, 100, True), (5, 128, False), (3, 70, False)])
def test_loglik_tc_points_in_tensor_memory(gran, K, d, tri):
    """tc_estep3.cu (points operand in tensor memory, triangular skip; off by default because it measured slower):
    Cholesky-factor operands take the staircase of narrow MMAs, dense operands the full rows."""
    E = eng()
    rng = np.random.default_rng(17)
    N = 1300
    ops = E.QuadOperands(K, d, d, 'fp32')
    W = np.zeros((K, ops.Rp, ops.Dpp))
    blk = rng.standard_normal((K, d, d)) / np.sqrt(d)
    W[:, :d, :d] = np.triu(blk) if tri else blk
    W[:, :d, d] = rng.standard_normal((K, d))
    cst = rng.standard_normal(K)
    ops.W.copy_(E.to_dev(W, torch.float32))
    ops.cst.copy_(E.to_dev(cst, torch.float32))
    Z = E.to_dev(rng.standard_normal((N, d)) * 1.5 + 0.5, torch.float32)
    old = E.set_triangular(gran)
    try:
        ll = E.loglik_tc(Z, ops)
    finally:
        E.set_triangular(old)
    Wr = ops.W.double().cpu().numpy()
    zt = np.concatenate([Z.double().cpu().numpy(), np.ones((N, 1))], axis=1)
    y = np.einsum('kij,nj->kni', Wr[:, :, :d + 1], zt)
    ref = ops.cst.double().cpu().numpy()[:, None] - 0.5 * np.sum(y * y, axis=2)
    close(ll, ref, 1e-4, 'log-lik, points operand in tensor memory (G=%d)' % gran)
    close(ll, E.loglik(Z, ops).cpu().numpy(), 2e-5, 'tensor-core vs CUDA-core log-lik')


def test_loglik_tc_strided_rows_and_tiny_scale():
    """ldz > D, unaligned row stride (scalar load path), data of magnitude 1e-3."""
    E = eng()
    rng = np.random.default_rng(0)
    K, d, N = 6, 37, 1000
    big = torch.zeros((N, d + 3), dtype=torch.float32, device=E.device())
    x = rng.standard_normal((N, d)) * 1e-3
    big[:, :d] = E.to_dev(x, torch.float32)
    Z = big[:, :d]
    mus = rng.standard_normal((K, d)) * 1e-3
    lmbdas = np.stack([spd(rng, d, 1e6) for _ in range(K)])
    ops = E.QuadOperands(K, d, d, 'fp32')
    E.operands_gauss(ops, E.to_dev(mus), E.to_dev(lmbdas)).check()
    ll = E.loglik_tc(Z, ops)
    ref = orc.gauss_full_loglik(Z.double().cpu().numpy(), mus, lmbdas)
    close(ll, ref, 1e-4, 'strided tensor-core log-lik')


@pytest.mark.parametrize('K,d,N,flush', [(3, 128, 400, 16), (6, 128, 5000, 2), (70, 16, 2100, 16), (5, 40, 1000, 1),
                                         (9, 128, 20000, 16), (1, 3, 130, 16),
                                         # small-dimension feature form (tc_sstats.cu, d <= 21): cfg2 / cfg4 shapes, several component tiles
                                         (128, 9, 5000, 16), (64, 16, 70000, 16), (300, 21, 3000, 16), (2, 1, 64, 16),
                                         # feature-form kernel (64 < d <= 128): ragged K / d / N, several component blocks
                                         (300, 100, 3000, 16), (130, 65, 1001, 1), (1, 128, 63, 16), (257, 127, 9000, 4)])
def test_stats_tc(K, d, N, flush):
    E = eng()
    from mimo_b200 import _lib
    rng = np.random.default_rng(K + d)
    x = rng.standard_normal((N, d)) + 2.0
    w = rng.random((K, N)) ** 4
    w /= w.sum(0)
    Z, R = E.to_dev(x, torch.float32), E.to_dev(w, torch.float32)
    feats = E.quad_features(d)
    _lib.load().mimo_tc_set_flush_tiles(flush)
    try:
        st = E.stats_soft_tc(Z, R, feats).cpu().numpy()
    finally:
        _lib.load().mimo_tc_set_flush_tiles(0)      # 0: back to each kernel's default
    cc = E.stats_soft(Z, R, feats, 'fp32').cpu().numpy()
    ref = orc.gauss_full_wstats(Z.double().cpu().numpy(), R.double().cpu().numpy())
    S, C = unpack_quad(st, d), unpack_quad(cc, d)
    close(S[:, :d, :d], ref[2], 1e-4, 'tc sum r xx')
    close(S[:, d, :d], ref[0], 1e-4, 'tc sum r x')
    close(S[:, d, d], ref[1], 1e-4, 'tc sum r')
    e_tc, e_cc = scaled_err(S[:, :d, :d], ref[2]), scaled_err(C[:, :d, :d], ref[2])
    print('stats K=%d d=%d N=%d flush=%d: scaled err tensor-core %.2e, CUDA-core fp32 %.2e' % (K, d, N, flush, e_tc, e_cc))
    # FP32 accumulation in TMEM over one flush window (16 x 128 points) costs ~1e-5 of the largest entry
    assert e_tc <= max(8 * e_cc, 3e-5)


def test_stats_tc_accumulates_into_stat():
    """stat is accumulated into (the contract of mimo_stats_soft)."""
    E = eng()
    rng = np.random.default_rng(1)
    K, d, N = 4, 32, 600
    Z = E.to_dev(rng.standard_normal((N, d)), torch.float32)
    R = E.to_dev(rng.random((K, N)), torch.float32)
    feats = E.quad_features(d)
    one = E.stats_soft_tc(Z, R, feats)
    two = E.stats_soft_tc(Z, R, feats, stat=one.clone())
    close(two, 2 * one.cpu().numpy(), 1e-12, 'accumulation')


@pytest.mark.parametrize('hard', [False, True])
@pytest.mark.parametrize('K,d,N', [(12, 32, 70000), (7, 128, 45000), (64, 16, 60000), (128, 9, 40000), (20, 21, 9000), (40, 8, 5000),
                                   (9, 50, 20000), (3, 23, 3000)])
def test_sweep_tc_matches_cuda_cores_and_oracle(hard, K, d, N):
    """mimo_sweep on the tensor-core path == the CUDA-core FP32 path within FP32 tolerance,
    and the statistics / lower-bound term match the oracle."""
    E = eng()
    rng = np.random.default_rng(21)
    x = rng.standard_normal((N, d)) + rng.integers(0, 3, size=(N, 1))
    mus = rng.standard_normal((K, d)) + rng.integers(0, 3, size=(K, 1))
    lmbdas = np.stack([spd(rng, d) for _ in range(K)])
    logw = np.log(rng.dirichlet(np.ones(K)))
    ops = E.QuadOperands(K, d, d, 'fp32')
    E.set_log_weights(ops, logw)
    E.operands_gauss(ops, E.to_dev(mus), E.to_dev(lmbdas)).check()
    Z = E.to_dev(x, torch.float32)
    feats = E.quad_features(d)
    assert E.sweep_uses_tensor_cores(ops, d)
    u = E.to_dev(rng.random(N))
    buf = E.SweepBuffers(N, K, feats.F, 'fp32', hard)
    E.sweep(Z, ops, feats, buf, uniforms=u if hard else None)
    old = E.set_tensor_cores(0)
    try:
        assert not E.sweep_uses_tensor_cores(ops, d)
        ref = E.SweepBuffers(N, K, feats.F, 'fp32', hard)
        E.sweep(Z, ops, feats, ref, uniforms=u if hard else None)
    finally:
        E.set_tensor_cores(old)
    xr = Z.double().cpu().numpy()
    ll = orc.gauss_full_loglik(xr, mus, lmbdas) + logw[:, None]
    resp, lse = orc.responsibilities(ll)
    close(buf.lse_sum, [lse.sum()], 1e-6, 'lse sum vs oracle')
    close(buf.lse_sum, ref.lse_sum.cpu().numpy(), 1e-6, 'lse sum vs CUDA cores')
    if hard:
        lab, lab_cc = buf.labels.cpu().numpy(), ref.labels.cpu().numpy()
        lab_ref = orc.sample_discrete_from_log(ll, u.cpu().numpy())
        safe = orc.label_boundary_distance(ll, u.cpu().numpy()) > 1e-3
        assert np.array_equal(lab[safe], lab_ref[safe])
        assert (lab == lab_cc).mean() > 0.999
        st = orc.gauss_full_wstats(xr, orc.one_hot(lab, K))
    else:
        st = orc.gauss_full_wstats(xr, resp)
    S = unpack_quad(buf.stat.cpu().numpy(), d)
    close(S[:, :d, :d], st[2], 1e-4, 'sweep sum r xx')
    close(S[:, d, :d], st[0], 1e-4, 'sweep sum r x')
    close(S[:, d, d], st[1], 1e-4, 'sweep sum r')
    if not hard:
        close(buf.stat, ref.stat.cpu().numpy(), 2e-5, 'tensor-core vs CUDA-core statistics')


@pytest.mark.parametrize('hard', [False, True])
@pytest.mark.parametrize('K,d,N,sep', [(64, 128, 6000, 6.0), (40, 96, 5000, 6.0), (48, 64, 4000, 0.3), (33, 100, 3000, 1.5),
                                     (40, 50, 5000, 6.0), (36, 30, 4000, 8.0), (35, 27, 3000, 8.0)])
def test_sweep_screened_estep(hard, K, d, N, sep):
    """default tensor-core mode: single-pass screening + exact refinement of the candidates (separated components)
    or the device-selected dense second pass (overlapping components).  Either way the sweep must match the dense
    CTA-pair path and the oracle at the FP32 tolerance."""
    E = eng()
    rng = np.random.default_rng(K + d)
    centres = sep * rng.standard_normal((K, d))
    z = rng.integers(0, K, size=N)
    x = centres[z] + rng.standard_normal((N, d))
    mus = centres + 0.2 * rng.standard_normal((K, d))
    lmbdas = np.stack([spd(rng, d) for _ in range(K)])
    logw = np.log(rng.dirichlet(np.ones(K)))
    ops = E.QuadOperands(K, d, d, 'fp32')
    E.set_log_weights(ops, logw)
    E.operands_gauss(ops, E.to_dev(mus), E.to_dev(lmbdas)).check()
    Z = E.to_dev(x, torch.float32)
    feats = E.quad_features(d)
    u = E.to_dev(rng.random(N))
    buf = E.SweepBuffers(N, K, feats.F, 'fp32', hard)
    E.sweep(Z, ops, feats, buf, uniforms=u if hard else None)
    cands, dense = E.screen_last()
    print('screened sweep K=%d d=%d N=%d sep=%.1f: %d candidates of %d pairs (%.2f%%), dense fallback %d'
          % (K, d, N, sep, cands, N * K, 100.0 * cands / (N * K), dense))
    assert cands >= N                                   # every point keeps at least its best component
    if sep >= 6.0:
        assert dense == 0                               # well separated: refined lists
    elif sep < 1.0:
        assert dense == 1                               # overlapping: device-selected dense fallback
    old = E.set_tensor_cores(3)                         # CTA pairs, dense 3-pass E-step
    try:
        ref = E.SweepBuffers(N, K, feats.F, 'fp32', hard)
        E.sweep(Z, ops, feats, ref, uniforms=u if hard else None)
    finally:
        E.set_tensor_cores(old)
    xr = Z.double().cpu().numpy()
    ll = orc.gauss_full_loglik(xr, mus, lmbdas) + logw[:, None]
    resp, lse = orc.responsibilities(ll)
    close(buf.lse_sum, [lse.sum()], 1e-6, 'lse sum vs oracle')
    close(buf.lse_sum, ref.lse_sum.cpu().numpy(), 1e-6, 'lse sum vs dense tensor-core path')
    if hard:
        lab, lab_d = buf.labels.cpu().numpy(), ref.labels.cpu().numpy()
        lab_ref = orc.sample_discrete_from_log(ll, u.cpu().numpy())
        safe = orc.label_boundary_distance(ll, u.cpu().numpy()) > 1e-3
        assert np.array_equal(lab[safe], lab_ref[safe])
        assert (lab == lab_d).mean() > 0.999
        st = orc.gauss_full_wstats(xr, orc.one_hot(lab, K))
    else:
        st = orc.gauss_full_wstats(xr, resp)
    S = unpack_quad(buf.stat.cpu().numpy(), d)
    close(S[:, :d, :d], st[2], 1e-4, 'screened sweep sum r xx')
    close(S[:, d, :d], st[0], 1e-4, 'screened sweep sum r x')
    close(S[:, d, d], st[1], 1e-4, 'screened sweep sum r')
    if not hard:
        close(buf.stat, ref.stat.cpu().numpy(), 2e-5, 'screened vs dense statistics')
        # mode 4: screened E-step with the dense tensor-core statistics instead of the pair-list kernel
        old = E.set_tensor_cores(4)
        try:
            mid = E.SweepBuffers(N, K, feats.F, 'fp32', hard)
            E.sweep(Z, ops, feats, mid)
        finally:
            E.set_tensor_cores(old)
        close(buf.stat, mid.stat.cpu().numpy(), 2e-5, 'pair-list vs dense statistics behind the screened E-step')
        # mode 5: the second screening tier (all operand rows, one FP16 pass) -- a tighter bound than the projection,
        # the tier a multi-chunk sweep moves to by itself when the projection leaves too many candidates
        old = E.set_tensor_cores(5)
        try:
            t1 = E.SweepBuffers(N, K, feats.F, 'fp32', hard)
            E.sweep(Z, ops, feats, t1)
            c1, d1 = E.screen_last()
        finally:
            E.set_tensor_cores(old)
        print('  all-rows tier: %d candidates (%.2f%%), dense fallback %d' % (c1, 100.0 * c1 / (N * K), d1))
        assert c1 <= max(cands, N) or dense == 1
        if sep >= 1.0:
            assert d1 == 0
        close(t1.stat, ref.stat.cpu().numpy(), 2e-5, 'all-rows screening tier vs dense statistics')
        close(t1.lse_sum, ref.lse_sum.cpu().numpy(), 1e-6, 'all-rows screening tier lse sum')


def test_screened_sweep_large_properties():
    """The sweep at a size the oracle cannot reach (N = 1.5 M, K = 256, d = 128: several point chunks of the tensor-core
    path), checked through size-independent properties: responsibilities of every point sum to one (soft counts sum
    to N), the statistics are additive over data shards (the reference's list-of-arrays semantics, gaussian.py:503-505),
    the screened default path and the dense 3-pass path agree, and a Gibbs sweep's counts equal the label histogram."""
    E = eng()
    K, d, N = 256, 128, 1_500_000
    g = torch.Generator(device='cuda')
    g.manual_seed(11)
    centres = 4.0 * torch.randn(K, d, generator=g, device='cuda')
    z = torch.randint(0, K, (N,), generator=g, device='cuda')
    Z = (centres[z] + torch.randn(N, d, generator=g, device='cuda')).contiguous()
    rng = np.random.default_rng(3)
    mus = centres.double().cpu().numpy() + 0.1 * rng.standard_normal((K, d))
    lmbdas = np.stack([spd(rng, d) for _ in range(K)])
    ops = E.QuadOperands(K, d, d, 'fp32')
    E.set_log_weights(ops, np.log(rng.dirichlet(np.ones(K))))
    E.operands_gauss(ops, E.to_dev(mus), E.to_dev(lmbdas)).check()
    feats = E.quad_features(d)
    buf = E.SweepBuffers(N, K, feats.F, 'fp32', False)
    E.sweep(Z, ops, feats, buf)
    cands, dense = E.screen_last()
    assert dense == 0 and cands < 0.04 * N * K
    assert E.screen_level() == 0                        # separated components: the projection tier all the way
    stat = buf.stat.cpu().numpy()
    counts = stat[:, -1]
    assert abs(counts.sum() - N) <= 1e-6 * N
    # additivity over two shards
    h = N // 2 + 12345
    tot = np.zeros_like(stat)
    lse = 0.0
    for lo, hi in ((0, h), (h, N)):
        b = E.SweepBuffers(hi - lo, K, feats.F, 'fp32', False)
        E.sweep(Z[lo:hi], ops, feats, b)
        tot += b.stat.cpu().numpy()
        lse += b.lse_sum.item()
    close(tot, stat, 1e-6, 'statistics additive over shards')
    assert abs(lse - buf.lse_sum.item()) <= 1e-7 * abs(lse)
    # dense path
    old = E.set_tensor_cores(3)
    try:
        ref = E.SweepBuffers(N, K, feats.F, 'fp32', False)
        E.sweep(Z, ops, feats, ref)
    finally:
        E.set_tensor_cores(old)
    close(stat, ref.stat.cpu().numpy(), 2e-5, 'screened vs dense statistics (large)')
    assert abs(ref.lse_sum.item() - buf.lse_sum.item()) <= 1e-6 * abs(ref.lse_sum.item())
    # Gibbs: counts == histogram of the drawn labels, labels independent of sharding (Philox keyed by global index)
    bh = E.SweepBuffers(N, K, feats.F, 'fp32', True)
    E.sweep(Z, ops, feats, bh, seed=99)
    lab = bh.labels.cpu().numpy()
    assert np.array_equal(bh.stat.cpu().numpy()[:, -1], np.bincount(lab, minlength=K))
    b2 = E.SweepBuffers(N - h, K, feats.F, 'fp32', True)
    E.sweep(Z[h:], ops, feats, b2, seed=99, offset=h)
    assert (b2.labels.cpu().numpy() == lab[h:]).mean() > 0.9999


def test_screen_tiers_across_chunks():
    """A sweep of several point chunks over moderately separated components (K = 1024, d = 100, centre spread 1.5): the
    32-row projection leaves too many candidates on the first chunk, which takes the dense pass and moves the sweep to
    the all-rows screening tier; the remaining chunks are screened there.  The device-side tier state must not change
    the results: statistics and log-normalisers agree with the dense path."""
    E = eng()
    K, d, N = 1024, 100, 2_200_000
    g = torch.Generator(device='cuda')
    g.manual_seed(5)
    centres = 1.5 * torch.randn(K, d, generator=g, device='cuda')
    z = torch.randint(0, K, (N,), generator=g, device='cuda')
    Z = (centres[z] + torch.randn(N, d, generator=g, device='cuda')).contiguous()
    rng = np.random.default_rng(8)
    mus = centres.double().cpu().numpy() + 0.05 * rng.standard_normal((K, d))
    lmbdas = np.stack(K * [np.eye(d)]) * (0.8 + 0.4 * rng.random((K, 1, 1)))
    ops = E.QuadOperands(K, d, d, 'fp32')
    E.set_log_weights(ops, np.log(rng.dirichlet(5.0 * np.ones(K))))
    E.operands_gauss(ops, E.to_dev(mus), E.to_dev(lmbdas)).check()
    feats = E.quad_features(d)
    buf = E.SweepBuffers(N, K, feats.F, 'fp32', False)
    E.sweep(Z, ops, feats, buf)
    cands, dense = E.screen_last()                      # the LAST chunk: screened on the all-rows tier
    level = E.screen_level()
    print('tiers across chunks: sweep ended on tier %d; last chunk %d candidates, dense fallback %d' % (level, cands, dense))
    old = E.set_tensor_cores(3)
    try:
        ref = E.SweepBuffers(N, K, feats.F, 'fp32', False)
        E.sweep(Z, ops, feats, ref)
    finally:
        E.set_tensor_cores(old)
    close(buf.stat, ref.stat.cpu().numpy(), 2e-5, 'tiered vs dense statistics')
    assert abs(ref.lse_sum.item() - buf.lse_sum.item()) <= 1e-6 * abs(ref.lse_sum.item())
    assert abs(buf.stat.cpu().numpy()[:, -1].sum() - N) <= 1e-6 * N
    assert dense == 0 and level == 1, 'the sweep should have moved to the all-rows tier and screened the last chunk there (level %d)' % level


def test_absmax_hint_is_scoped_to_its_sweep():
    """mimo_sweep_absmax_hint belongs to the NEXT SWEEP, whatever kernels that sweep runs: a small-dimension sweep stays on
    the CUDA cores and used to leave its max |Z| behind, and the next tensor-core call -- on other data -- scaled its FP16
    operands with it (overflow -> NaN when the new data are larger).  Found by test order in round 2."""
    E = eng()
    rng = np.random.default_rng(3)
    # 1. a d = 2 sweep (no tensor cores) on tiny-valued data, with the hint the Python Session gives
    K0, d0, N0 = 3, 2, 400
    x0 = 1e-3 * rng.standard_normal((N0, d0))
    ops0 = E.QuadOperands(K0, d0, d0, 'fp32')
    E.set_log_weights(ops0, np.log(np.ones(K0) / K0))
    E.operands_gauss(ops0, E.to_dev(1e-3 * rng.standard_normal((K0, d0))), E.to_dev(np.stack(K0 * [1e6 * np.eye(d0)]))).check()
    Z0 = E.to_dev(x0, torch.float32)
    feats0 = E.quad_features(d0)
    assert not E.sweep_uses_tensor_cores(ops0, d0)
    E.sweep(Z0, ops0, feats0, E.SweepBuffers(N0, K0, feats0.F, 'fp32', False), absmax=float(np.abs(x0).max()))
    # 2. the tensor-core log-likelihood on data five orders of magnitude larger
    K, d, N = 6, 64, 500
    x = 40. * rng.standard_normal((N, d))
    mus = 40. * rng.standard_normal((K, d))
    lmbdas = np.stack([spd(rng, d) / 1600. for _ in range(K)])
    ops = E.QuadOperands(K, d, d, 'fp32')
    E.set_log_weights(ops, np.zeros(K))
    E.operands_gauss(ops, E.to_dev(mus), E.to_dev(lmbdas)).check()
    Z = E.to_dev(x, torch.float32)
    ll = E.loglik_tc(Z, ops)
    assert torch.isfinite(ll).all()
    close(ll, orc.gauss_full_loglik(Z.double().cpu().numpy(), mus, lmbdas), 1e-4, 'tensor-core log-lik after a hinted CUDA-core sweep')
    # ... and a tensor-core sweep takes its own hint
    feats = E.quad_features(d)
    buf = E.SweepBuffers(N, K, feats.F, 'fp32', False)
    E.sweep(Z, ops, feats, buf, absmax=float(np.abs(x).max()))
    resp, lse = orc.responsibilities(orc.gauss_full_loglik(Z.double().cpu().numpy(), mus, lmbdas))
    assert abs(buf.lse_sum.item() - lse.sum()) <= 1e-5 * abs(lse.sum())
