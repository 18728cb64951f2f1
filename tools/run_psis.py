"""Small driver for profiling: PSIS on n float64 log-weights (default 1e8)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import viabel_b200 as vb

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
gen = torch.Generator(device='cuda')
gen.manual_seed(7)
lw = torch.zeros(n, device='cuda', dtype=torch.float64)
for _ in range(2):
    z = torch.randn(n, generator=gen, device='cuda', dtype=torch.float64)
    lw += -5.5 * torch.log1p(z * z / 10.0) + 20.5 * torch.log1p(z * z / 40.0)
    del z
out = torch.empty_like(lw)
for _ in range(reps):
    _, res, _, _ = vb.psislw_device(lw, out)
torch.cuda.synchronize()
print(res.cpu().numpy())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(10):
    e0.record()
    vb.psislw_device(lw, out)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print('psislw n=%d: median %.4f ms  min %.4f ms  -> %.3e draws/s, %.1f%% of 6543.7 GB/s at 24 B/draw'
      % (n, ts[len(ts) // 2], ts[0], n / ts[len(ts) // 2] * 1e3, 100 * 24.0 * n / ts[len(ts) // 2] * 1e3 / 6543.7e9))

# k-hat / moments only (no output array): 16 B/draw algorithmic
_, res0, _, _ = vb.psislw_device(lw, None)
torch.cuda.synchronize()
print('moments-only result', res0.cpu().numpy())
r_full, r_mom = res.cpu().numpy(), res0.cpu().numpy()
for name, i in (('khat', 0), ('lse', 4), ('sumv', 7), ('sumexp2v', 8)):
    print('  %-9s full %.17g  moments-only %.17g  rel diff %.2e'
          % (name, r_full[i], r_mom[i], abs(r_full[i] - r_mom[i]) / max(abs(r_full[i]), 1e-300)))
ts = []
for _ in range(10):
    e0.record()
    vb.psislw_device(lw, None)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print('moments-only n=%d: median %.4f ms  min %.4f ms  -> %.1f%% of 6543.7 GB/s at 16 B/draw'
      % (n, ts[len(ts) // 2], ts[0], 100 * 16.0 * n / ts[len(ts) // 2] * 1e3 / 6543.7e9))
