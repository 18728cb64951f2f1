"""Draw-sharded PSIS under torchrun: parity with the single-GPU result at a small size, then timing.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/run_psis_sharded.py --draws 200000000
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--draws', dest='n', type=int, default=200_000_000, help='global number of draws')
    ap.add_argument('--reps', type=int, default=5)
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    if world > 1:
        dist.init_process_group('nccl')
    import viabel_b200 as vb
    from viabel_b200.parallel import shard_rows

    # parity: every rank builds the same global column, smooths its shard, rank 0 compares with one-GPU psislw
    g = torch.Generator(device='cuda').manual_seed(7)
    n0 = 1_000_003
    full = torch.randn(n0, dtype=torch.float64, device='cuda', generator=g)
    full = 1.5 * full + 0.3 * full ** 2
    lo, hi = shard_rows(n0, rank, world)
    out, k, res = vb.psislw_sharded(full[lo:hi].contiguous())
    ref, kref = vb.psislw(full.clone())
    err = float((out - ref[lo:hi]).abs().max())
    from viabel_b200._psis import psislw_device
    _, r1, _, _ = psislw_device(full, None, 1.0)
    r1 = r1.cpu().numpy()
    ok = (k == kref) and err < 1e-11 and abs(res[7] - r1[7]) <= 1e-10 * abs(r1[7]) and abs(res[8] - r1[8]) <= 1e-10 * abs(r1[8])
    if world > 1:
        flag = torch.tensor([float(ok)], device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())

    # timing: n/world draws per rank (strong scaling of one column)
    lo, hi = shard_rows(a.n, rank, world)
    g.manual_seed(100 + rank)
    x = torch.randn(hi - lo, dtype=torch.float64, device='cuda', generator=g)
    x = 1.5 * x + 0.3 * x ** 2
    o = torch.empty_like(x)
    sizes = [shard_rows(a.n, r, world)[1] - shard_rows(a.n, r, world)[0] for r in range(world)]
    ms = []
    for it in range(a.reps + 2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, kk, _ = vb.psislw_sharded(x, out=o, sizes=sizes)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it >= 2:
            ms.append(float(t.item()))
    if rank == 0:
        best = float(np.median(ms))
        print(json.dumps({'metric': 'psis_sharded_draws_per_s', 'n_global': a.n, 'n_gpus': world, 'parity_ok': bool(ok),
                          'parity_max_abs_err': err, 'khat': kk, 'ms': best, 'value': a.n / best * 1e3,
                          'note': 'end to end incl. the record all-gather, the moment all-reduce and the final status read'}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
