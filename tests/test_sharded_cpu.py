"""Data-sharded driver on CPU: world_size-2 gloo processes.  The product's sharding logic (mimo_b200/sharded.py:
contiguous shards, ONE all-reduce of the packed FP64 statistics + lower-bound scalar per sweep, global point offsets
for the Philox label draws) is exercised with per-shard statistics produced by the oracle -- the checker, standing in
for the CUDA kernels that cannot run here.  Reference semantics: a list of arrays is summed shard by shard
(distributions/gaussian.py:503-505, categorical.py:45-46, utils/abstraction.py:12-14)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pack(st, lse):
    """statistics tuple (sum r x, sum r, sum r xx^T, sum r) + scalar -> one flat FP64 message."""
    return torch.from_numpy(np.concatenate([st[0].ravel(), st[1].ravel(), st[2].ravel(), [lse]]))


def _worker(rank, world, port, N, K, d, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from mimo_b200.sharded import Communicator, init_from_env, shard_bounds
    from oracle import mimo_oracle as orc
    r, w = init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    rng = np.random.default_rng(5)                       # every rank sees the same data set and model
    x = rng.standard_normal((N, d)) + 2.0 * rng.integers(0, 3, size=(N, 1))
    mus = rng.standard_normal((K, d)) * 2
    lmbdas = np.stack([np.eye(d) * (0.5 + rng.random()) for _ in range(K)])
    logw = np.log(rng.dirichlet(np.ones(K)))
    comm = Communicator(N_global=N)
    lo, hi = shard_bounds(N, rank, world)
    assert comm.point_offset == lo and comm.world == world
    ll = orc.gauss_full_loglik(x[lo:hi], mus, lmbdas) + logw[:, None]
    resp, lse = orc.responsibilities(ll)
    msg = _pack(orc.gauss_full_wstats(x[lo:hi], resp), lse.sum())
    comm.allreduce(msg)
    assert comm.messages == 1 and comm.bytes == msg.numel() * 8
    # labels drawn from uniforms indexed by GLOBAL point: independent of the shard count
    u = np.random.default_rng(9).random(N)
    lab = orc.sample_discrete_from_log(ll, u[lo:hi])
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), msg=msg.numpy(), lab=lab, lo=lo, hi=hi)
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_allreduce_matches_whole_data(tmp_path):
    from oracle import mimo_oracle as orc
    N, K, d, world = 1001, 5, 3, 2                       # odd N: shards of 501 and 500 points
    mp.spawn(_worker, args=(world, _free_port(), N, K, d, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(5)
    x = rng.standard_normal((N, d)) + 2.0 * rng.integers(0, 3, size=(N, 1))
    mus = rng.standard_normal((K, d)) * 2
    lmbdas = np.stack([np.eye(d) * (0.5 + rng.random()) for _ in range(K)])
    logw = np.log(rng.dirichlet(np.ones(K)))
    ll = orc.gauss_full_loglik(x, mus, lmbdas) + logw[:, None]
    resp, lse = orc.responsibilities(ll)
    whole = _pack(orc.gauss_full_wstats(x, resp), lse.sum()).numpy()
    outs = [np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r)) for r in range(world)]
    for o in outs:                                        # every rank holds the same reduced message
        assert np.allclose(o['msg'], whole, rtol=1e-12, atol=1e-9)
    assert (int(outs[0]['lo']), int(outs[0]['hi']), int(outs[1]['lo']), int(outs[1]['hi'])) == (0, 501, 501, 1001)
    lab = orc.sample_discrete_from_log(ll, np.random.default_rng(9).random(N))
    assert np.array_equal(np.concatenate([o['lab'] for o in outs]), lab)


def test_shard_bounds_cover_and_balance():
    sys.path.insert(0, ROOT)
    from mimo_b200.sharded import shard_bounds
    for N in (0, 1, 7, 50_000_000, 100_000_001):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(N, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == N
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
