"""Multi-GPU check of the sharded hot path (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/check_sharded_step.py

* the peer-memory communicator maps every rank's buffer (CUDA IPC) and its all-reduce equals NCCL's;
* the fused step on row-sharded observations (in-kernel exchange) reproduces the single-GPU run on all rows:
  float64 path to 1e-10, tensor-core path to 1e-4, and the parameters are BIT-IDENTICAL on every rank;
* AlphaDivergence on a sharded model evaluates the same samples on every rank (rank 0's seed);
* draw-sharded PSIS equals the single-GPU result.
Prints one PASS / FAIL line per check on rank 0 and exits non-zero on failure.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import viabel_b200 as vb
    from viabel_b200.engine import FusedStep
    from viabel_b200.parallel import get_communicator, shard_rows
    from _problems import logistic_problem
    ok_all = True

    def report(name, ok, detail=''):
        nonlocal ok_all
        flag = torch.tensor([1 if ok else 0], device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok_all = ok_all and bool(flag.item())
        if rank == 0:
            print('%s %s %s' % ('PASS' if flag.item() else 'FAIL', name, detail), flush=True)

    # 1. communicator
    comm = get_communicator(None, 4096 * 8)
    report('communicator (CUDA IPC peer mapping)', comm is not None)
    if comm is None:
        dist.destroy_process_group()
        sys.exit(1)
    rs = np.random.RandomState(100 + rank)
    for n in (1, 7, 1280, 4096):
        a = torch.as_tensor(rs.randn(n), device='cuda')
        b = a.clone()
        comm.allreduce_sum_(a)
        dist.all_reduce(b)
        gathered = [torch.empty_like(a) for _ in range(world)]
        dist.all_gather(gathered, a)
        same = all(torch.equal(g, gathered[0]) for g in gathered)
        report('all-reduce n=%d' % n, same and float((a - b).abs().max()) <= 1e-12 * max(1.0, float(b.abs().max())))

    # 2. fused step, sharded vs all rows on one GPU
    N, d, S = 30000, 96, 48
    X, y, beta = logistic_problem(N, d, seed=7)
    lo, hi = shard_rows(N, rank, world)
    vp0 = np.concatenate([0.5 * beta, -1.0 * np.ones(d)])
    for path, tol in (('f64', 1e-10), ('fast', 1e-4)):
        res = {}
        for sharded in (False, True):
            model = vb.LogisticRegression(X[lo:hi], y[lo:hi], sharded=True) if sharded else vb.LogisticRegression(X, y)
            approx = vb.MFGaussian(d, seed=31)
            if path == 'fast':
                model.enable_fast_path(approx)
            eng = FusedStep(vb.ExclusiveKL(approx, model, S), vb.RMSProp(0.02), hist_len=12)
            eng.set_param(vp0)
            eng.run(12)
            torch.cuda.synchronize()
            eng.check_comm()
            res[sharded] = (eng.vp.clone(), eng.value_hist.clone())
        vp_s, val_s = res[True]
        vp_1, val_1 = res[False]
        gathered = [torch.empty_like(vp_s) for _ in range(world)]
        dist.all_gather(gathered, vp_s)
        identical = all(torch.equal(g, gathered[0]) for g in gathered)
        ev = float((val_s - val_1).norm() / val_1.norm())
        ep = float(((vp_s - vp_1)).norm() / (vp_1 - torch.as_tensor(vp0, device='cuda')).norm())
        report('fused step %s: sharded == single' % path, ev < tol and ep < 20 * tol, 'value %.2e param-move %.2e' % (ev, ep))
        report('fused step %s: parameters bit-identical on all ranks' % path, identical)

    # 3. AlphaDivergence on a sharded model: same draws everywhere (each rank seeds numpy differently on purpose)
    np.random.seed(1000 + rank)
    model = vb.LogisticRegression(X[lo:hi], y[lo:hi], sharded=True)
    approx = vb.MFGaussian(d, seed=5)
    v, g = vb.AlphaDivergence(approx, model, S, 2.0)(torch.as_tensor(vp0, device='cuda'))
    gathered = [torch.empty_like(g) for _ in range(world)]
    dist.all_gather(gathered, g)
    report('AlphaDivergence sharded: identical gradient on all ranks', all(torch.equal(x, gathered[0]) for x in gathered))
    full = vb.LogisticRegression(X, y)
    v1, g1 = vb.AlphaDivergence(vb.MFGaussian(d, seed=5), full, S, 2.0)(torch.as_tensor(vp0, device='cuda'),
                                                                        base=approx.last_base)
    report('AlphaDivergence sharded == single', float((g - g1).norm() / g1.norm()) < 1e-9)

    # 4. draw-sharded PSIS
    n = 2000000
    gen = torch.Generator(device='cuda')
    gen.manual_seed(3)
    z = torch.randn(n, generator=gen, device='cuda', dtype=torch.float64)
    lw = -5.5 * torch.log1p(z * z / 10.0) + 20.5 * torch.log1p(z * z / 40.0)
    a, b = shard_rows(n, rank, world)
    out, khat, res_ = vb.psislw_sharded(lw[a:b].contiguous())
    ref_out, ref_k = vb.psislw(lw)
    report('psislw sharded == single', abs(khat - ref_k) < 1e-10 * abs(ref_k) and
           float((out - ref_out[a:b]).abs().max()) < 1e-9, 'khat %.6f' % khat)

    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == '__main__':
    main()
