"""CPU, world_size 2, gloo: the host-side logic of the sharded path.

The CUDA kernels cannot run here, so each rank's sweep is computed by the numpy oracle on its row
shard; what is under test is the product's sharding / exchange code (viabel_b200.parallel):
row partition, the all-reduced [ll, gmu, ge] sum feeding the same objective assembly, identical
draws on every rank, and the ragged all-gather used by draw-sharded PSIS."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from viabel_b200 import parallel
        from oracle import viabel_oracle as vo
        from _problems import logistic_problem
        N, d, S = 1001, 7, 5
        X, y, beta = logistic_problem(N, d, seed=3)
        rs = np.random.RandomState(0)                     # identical draws on every rank
        base = rs.randn(S, d)
        vp = np.concatenate([beta, -1.0 * np.ones(d)])
        theta = vo.mfg_sample(vp, base)

        lo, hi = parallel.shard_rows(N, rank, world)
        # per-rank sweep outputs (what vb_glm_sweep_* returns for this rank's rows)
        Xr, yr = X[lo:hi], y[lo:hi]
        M = (Xr @ theta.T) * yr[:, None]
        ll = -vo.softplus(-M).sum(axis=0)
        R = vo.sigmoid(-M) * yr[:, None]
        gmu = Xr.T @ R.sum(axis=1)
        ge = np.sum(Xr * (R @ base), axis=0)
        buf = torch.from_numpy(np.concatenate([ll, gmu, ge]))
        parallel.allreduce_sum_(buf)
        ll, gmu, ge = (buf[:S].numpy(), buf[S:S + d].numpy(), buf[S + d:].numpy())

        # replicated objective assembly from the reduced sums
        sig = np.exp(vp[d:])
        prior = -0.5 * (theta ** 2).sum(axis=1) / 100.0 - d * np.log(10 * np.sqrt(2 * np.pi))
        value = -(np.mean(ll + prior) + vo.mfg_entropy(vp, d))
        a = gmu - theta.sum(axis=0) / 100.0
        b = ge - (theta * base).sum(axis=0) / 100.0
        grad = np.concatenate([-a / S, -b / S * sig - 1.0])
        v0, g0, _ = vo.exclusive_kl_meanfield(vp, base, lambda th: vo.logistic_logp_grad(th, X, y, 10.0))
        ok = abs(value - v0) < 1e-10 * abs(v0) and np.linalg.norm(grad - g0) < 1e-10 * np.linalg.norm(g0)

        # ragged all-gather (candidate lists of different lengths)
        mine = torch.arange(3 + 2 * rank, dtype=torch.float64) + 100 * rank
        allc = parallel.allgather_ragged(mine)
        expect = torch.cat([torch.arange(3 + 2 * r, dtype=torch.float64) + 100 * r for r in range(world)])
        ok = ok and torch.equal(allc, expect)
        ok = ok and parallel.world() == (rank, world) and parallel.is_distributed()

        # draw-sharded PSIS exchange: fixed-size records (values + int64 indices as bit patterns)
        # all-gathered in rank order, merged, compared with the one-column algorithm
        from _problems import psis_case
        lw = psis_case('t5_t7_1e5')
        n = lw.size
        lo, hi = parallel.shard_rows(n, rank, world)
        Mt = vo.psis_tail_len(n, 1.0)
        head, vals, idx = vo.psis_shard_record(lw[lo:hi], lo, Mt)
        rec = torch.from_numpy(np.concatenate([head, vals, idx.view(np.float64)]))
        bufs = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(bufs, rec)
        recs = []
        for b in bufs:
            b = b.numpy()
            recs.append((b[:4], b[4:4 + Mt + 1], b[4 + Mt + 1:].view(np.int64)))
        mx, cutoff, tidx, tv, body = vo.psis_merge_records(recs, Mt, n)
        x = lw - lw.max()
        xcut = max(np.partition(x, n - Mt - 1)[n - Mt - 1], np.log(np.finfo(float).tiny))
        ok = ok and mx == lw.max() and cutoff == xcut and np.array_equal(np.sort(tidx), np.flatnonzero(x > xcut))
        ok = ok and abs(body - np.exp(x[x <= xcut]).sum()) < 1e-12 * body
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_sharded_elbo_and_gather_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_shard_rows_partition():
    sys.path.insert(0, ROOT)
    from viabel_b200 import parallel
    for n in (0, 1, 7, 1000, 1000001):
        for ws in (1, 2, 3, 8):
            parts = [parallel.shard_rows(n, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_rows(10, 2, 2)
    assert parallel.world() == (0, 1) and not parallel.is_distributed()
