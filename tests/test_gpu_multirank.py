"""Multi-rank GPU parity (needs >= 2 GPUs; skipped otherwise): the data-sharded Session over NCCL -- statistics
all-reduce, posterior update split over ranks, operand all-gather -- equals the single-process run on statistics,
posteriors and lower bound (1e-9 in FP64); reference semantics: distributions/gaussian.py:503-505, utils/abstraction.py:12-14."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(600)
def test_two_rank_nccl_session_matches_single_process(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    out = str(tmp_path / 'verdict.json')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', '_multirank_worker.py'), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=570)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    verdict = json.load(open(out))
    print(json.dumps(verdict[0]))            # (shown with pytest -s / on failure)
    assert len(verdict) == 2
    for rank_result in verdict:
        for precision, res in rank_result.items():
            assert res['ok'], (precision, res)
            assert res['messages'] >= (1 if precision == 'abi_allreduce' else 4)
