"""Phase timestamps (clock64) of CTA 0 of the fast kernel, first 8 tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import viabel_b200 as vb
N, d, S = 200000, 512, 256
g = torch.Generator(device='cuda'); g.manual_seed(1)
X = torch.randn(N, d, generator=g, device='cuda', dtype=torch.float64)
y = torch.where(torch.rand(N, generator=g, device='cuda', dtype=torch.float64) < 0.5, 1.0, -1.0)
model = vb.LogisticRegression(X, y).enable_fast_path()
th = torch.randn(S, d, generator=g, device='cuda', dtype=torch.float64) * 0.05
bs = torch.randn(S, d, generator=g, device='cuda', dtype=torch.float64)
grid = min((N + 127) // 128, vb._lib.lib.vb_device_sm_count())
dbg = torch.zeros(grid * 49152 + 256, dtype=torch.float32, device='cuda')
out = torch.zeros(S + 2 * d, dtype=torch.float64, device='cuda')
for _ in range(2):
    model._sweep_fast(th, bs, None, True, out, debug=dbg)
torch.cuda.synchronize()
tim = dbg[grid * 49152:].view(torch.int64).cpu().numpy().reshape(8, 16)
t0 = tim[0, 0]
names = ['G1start', 'G1issued', 'r_full', 'G2b0', 'G2b1', 'G2b2', 'G2b3', '-', 'E1start', 'E1end', 'E2b0', 'E2b1', 'E2b2', 'E2b3', 'E2end', '-']
for t in range(6):
    print('tile', t, ' '.join('%s=%d' % (n, tim[t, i] - t0) for i, n in enumerate(names) if n != '-'))
