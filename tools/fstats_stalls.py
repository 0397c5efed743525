"""Where the MMA issuer of tc_fstats_kernel waits (cfg5 shape, one 1 M-point chunk): prints the stall clocks by cause."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mimo_b200 import _engine as E, _lib

K, d, N = 1024, 128, 1 << 20
g = torch.Generator(device='cuda'); g.manual_seed(0)
Z = torch.randn((N, d), generator=g, device='cuda', dtype=torch.float32)
R = torch.rand((K, N), generator=g, device='cuda', dtype=torch.float32)
R /= R.sum(0, keepdim=True)
feats = E.quad_features(d)
out = np.zeros(8, dtype=np.uint64)
for rep in range(3):
    st = E.stats_soft_tc(Z, R, feats)
    torch.cuda.synchronize()
    _lib.call('mimo_tc_fstats_stall_clocks', out.ctypes.data)
    tot, stages = float(out[5]), float(out[6])
    print('rep %d: issuer clocks %.3g (per cluster %.3g), stages %d; waits: A %.1f%%  peer A %.1f%%  B %.1f%%  peer B %.1f%%  drain %.1f%%; MMA issue floor 1536 clk x stages = %.1f%%'
          % (rep, tot, tot / 74, stages, 100 * out[0] / tot, 100 * out[1] / tot, 100 * out[2] / tot, 100 * out[3] / tot, 100 * out[4] / tot, 100 * 1536 * stages / tot))
