"""FASO's convergence statistics at BASELINE configs[3] scale: the full-rank MultivariateT family at d = 2048 has
P = 2 100 224 variational parameters; the iterate ring lives on the device and split-R-hat / ESS / MCSE run batched over
all parameters (csrc/faso.cu).  Runs a short FASO loop (the R-hat checks fire) and times one check and one MCSE."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import viabel_b200 as vb
from viabel_b200._mc_diagnostics import RingStats

G, p, n_per, S = 65, 31, 200, 64
rs = np.random.RandomState(20260118)
N = G * n_per
group = np.repeat(np.arange(G), n_per)
X = rs.randn(N, p); m = rs.randn(p); beta = m + 0.5 * rs.randn(G, p)
y = np.sum(X * beta[group], axis=1) + 0.3 * rs.randn(N)
model = vb.HierarchicalLinearRegression(X, y, group, G)
d = model.dim
approx = vb.MultivariateT(d, 100, seed=1)
vp0 = approx.init_param()
F = np.zeros((d, d)); F[np.diag_indices(d)] = 0.5 * np.log(0.01)
vp0[d:] = F[np.tril_indices(d)]; vp0[:G * p] = beta.reshape(-1); vp0[G * p:G * p + p] = m
objective = vb.AlphaDivergence(approx, model, S, 2.0)
sgo = vb.RMSProp(0.0005); sgo.progress = False
faso = vb.FASO(sgo, W_min=60, k_check=40, mcse_threshold=0.1)
n_iters = int(sys.argv[1]) if len(sys.argv) > 1 else 170
torch.cuda.synchronize(); t0 = time.perf_counter()
res = faso.optimize(n_iters, objective, vp0)
torch.cuda.synchronize(); t = time.perf_counter() - t0
print('FASO at P = %d parameters: %d iterations in %.2f s (%.1f ms/iteration incl. the checks); k_conv %s k_stopped %s'
      % (vp0.size, len(res['value_history']), t, 1e3 * t / len(res['value_history']), res['k_conv'], res['k_stopped']))
# the statistics alone, on a ring of the iterates just produced
hist = torch.as_tensor(res['variational_param_history'], device='cuda')
st = RingStats(hist, hist.shape[0], hist.shape[1])
end = hist.shape[0]
wins = np.linspace(60, int(0.95 * (end - 1)), num=5, dtype=int)
for name, fn in (('split R-hat, 5 windows', lambda: st.convergence_check(end, wins)), ('ESS + MCSE, W = 100', lambda: st.mcse(end, 100))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    print('%-24s %.1f ms' % (name, 1e3 * (time.perf_counter() - t0)))
