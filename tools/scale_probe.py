"""Where the multi-GPU step time goes: C2 sharded over the ranks, with and without the all-reduce,
eager and as a captured CUDA graph.  torchrun --nproc-per-node N tools/scale_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import viabel_b200 as vb
from viabel_b200.parallel import shard_rows
dev = torch.device('cuda', local)
N, d, S = 1000000, 512, 256
lo, hi = shard_rows(N, rank, world)
gen = torch.Generator(device=dev); gen.manual_seed(rank)
X = torch.randn(hi - lo, d, generator=gen, device=dev, dtype=torch.float64)
y = torch.where(torch.rand(hi - lo, generator=gen, device=dev, dtype=torch.float64) < 0.5, 1.0, -1.0)
model = vb.LogisticRegression(X, y, prior_scale=10.0, sharded=world > 1).enable_fast_path()
del X
approx = vb.MFGaussian(d, seed=3); approx.quantize_draws = 2
objective = vb.ExclusiveKL(approx, model, S)
opt = vb.RMSProp(0.01)
vp = torch.as_tensor(approx.init_param(), device=dev)

def step():
    value, grad = objective(vp)
    opt._fused_step(vp, grad, False)

def timeit(fn, n=50):
    for _ in range(5): fn()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

res = {}
res['eager'] = timeit(step)
orig = model._allreduce
model._allreduce = lambda buf: None
res['eager_no_allreduce'] = timeit(step)
model._allreduce = orig
theta = approx.sample(vp, S); base = approx.last_base
res['sweep_only_no_allreduce'] = timeit(lambda: (setattr(model, '_allreduce', lambda b: None), model.sweep(theta, base, None, True, ll_total_only=True), setattr(model, '_allreduce', orig)))
buf = torch.zeros(S + 2 * d, dtype=torch.float64, device=dev)
if world > 1:
    res['nccl_allreduce_10KB'] = timeit(lambda: dist.all_reduce(buf))
try:
    if world > 1: raise RuntimeError('skipped with NCCL')
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): step()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            step()
    res['graph'] = timeit(g.replay)
except Exception as e:
    res['graph'] = 'failed: %r' % (e,)
if rank == 0:
    print(world, 'GPUs:', {k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()})
if world > 1:
    dist.destroy_process_group()
