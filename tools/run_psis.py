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
