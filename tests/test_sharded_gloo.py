"""world_size-2 gloo test (CPU) of the multi-GPU path's host logic: every rank takes its contiguous row shard
(ks.shard_range), computes the partial sums the GPU path all-reduces ([A^T A | A^T R | colsum A | colsum R] with a
shared shift), the sums are reduced with torch.distributed, and every rank solves redundantly.  The result must
equal the single-process oracle -- this is the algebra of DESIGN.md section 4 end to end, incl. the NCCL-id broadcast
plumbing shape (an object broadcast from rank 0)."""
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
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import keystone_b200 as ks
    from oracle import keystone_oracle as ko
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)                       # same data on every rank, each takes its shard
        n, d, k, bs, lam, iters = 1003, 40, 4, 16, 0.3, 2
        F = rng.standard_normal((n, d)) + 1.5
        Y = ko.class_label_indicators(rng.integers(0, k, n), k)
        lo, hi = ks.shard_range(n, rank, world)
        Fl, Yl = F[lo:hi], Y[lo:hi]
        # object broadcast in the shape Context.from_torch_distributed uses for the NCCL id
        ids = [b"x" * 128 if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        assert ids[0] == b"x" * 128

        def allreduce(a):
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)); dist.all_reduce(t); return t.numpy()

        ntot = int(allreduce(np.array([hi - lo]))[0])
        ymean = allreduce(Yl.sum(0)) / ntot
        R = Yl - ymean
        bounds = ko.block_bounds(d, bs)
        xs = [np.zeros((e - s, k)) for s, e in bounds]
        shifts = []
        for it in range(iters):
            for j, (s, e) in enumerate(bounds):
                if it == 0:   # shift = mean of a per-rank row sample, identical on all ranks after the reduce
                    ns = min(64, hi - lo)
                    ssum = allreduce(np.r_[Fl[:ns, s:e].sum(0), ns])
                    shifts.append(ssum[:-1] / ssum[-1])
                A = Fl[:, s:e] - shifts[j]
                G, Cm, sa, sr = (allreduce(x) for x in ko.block_ls_partial_sums(A, R))
                dW, delta = ko.block_ls_solve_from_sums(G, Cm, sa, sr, ntot, lam, xs[j] if it > 0 else None)
                xs[j] = xs[j] + dW
                R = R - (A - delta) @ dW
        ref, b0, mus = ko.block_ls_fit(F, Y, bs, iters, lam)
        err = max(np.abs(a - b).max() for a, b in zip(xs, ref))
        ret[rank] = float(err)
    finally:
        dist.destroy_process_group()


def test_sharded_blockls_algebra_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world and all(v < 1e-9 for v in ret.values()), dict(ret)
