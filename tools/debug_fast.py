"""Stage test of the tensor-core path: raw accumulators of tile 0 (Z and Tt) against numpy."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import torch

import viabel_b200 as vb
from _problems import logistic_problem

N, d, S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 512, int(sys.argv[3]) if len(sys.argv) > 3 else 256
X, y, beta = logistic_problem(N, d, seed=3)
rs = np.random.RandomState(0)
base = rs.randn(S, d).astype(np.float16).astype(np.float64)
vp = np.concatenate([0.5 * beta, -1.0 * np.ones(d)])
theta = vp[:d] + np.exp(vp[d:]) * base
model = vb.LogisticRegression(X, y).enable_fast_path()
print('absmax', model.absmax)
th = torch.as_tensor(theta, device='cuda')
bs = torch.as_tensor(base, device='cuda')
grid = min((N + 127) // 128, vb._lib.lib.vb_device_sm_count())
dbg = torch.zeros(grid * 49152, dtype=torch.float32, device='cuda')
out = torch.zeros(S + 2 * d, dtype=torch.float64, device='cuda')
model._sweep_fast(th, bs, None, True, out, debug=dbg)
torch.cuda.synchronize()
dbg = dbg.cpu().numpy().reshape(grid, 49152)
Xy = X * y[:, None]
Z = Xy[:128] @ theta.T
Zg = dbg[0, :32768].reshape(128, 256)[:min(128, N), :S]
print('Z err', np.abs(Zg - Z[:, :S]).max(), 'scale', np.abs(Z).max())
a = Xy @ theta.T
R = 1 / (1 + np.exp(a))
T = (R[:128] @ base)          # [n, j]
Tg = dbg[0, 32768:].reshape(128, 128)       # [j, n]
print('Tt err', np.abs(Tg[:min(128, d), :min(128, N)] - T[:, :128].T[:min(128, d)]).max(), 'scale', np.abs(T).max())
ll = -np.log1p(np.exp(-a)).sum(0)
gmu = (Xy * R.sum(1)[:, None]).sum(0)
ge = (Xy * (R @ base)).sum(0)
o = out.cpu().numpy()
rel = lambda x, r: np.linalg.norm(x - r) / np.linalg.norm(r)
print('ll', rel(o[:S], ll), 'gmu', rel(o[S:S + d], gmu), 'ge', rel(o[S + d:], ge))
