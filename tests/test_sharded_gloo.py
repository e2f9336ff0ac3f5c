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


def _solve_worker(rank, world, port, ret):
    """Host logic of the column-sharded solve and of the global residual scale (engine.cu::fit_blockls, world > 1)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        b, k = 48, 7                                          # k not divisible by the world size: uneven slices
        A = rng.standard_normal((200, b))
        H = A.T @ A + 0.5 * np.eye(b)
        rhs = rng.standard_normal((b, k))
        L = np.linalg.cholesky(H)
        col0 = lambda r: k * r // world                       # rank r owns columns [k r / world, k (r+1) / world)
        m0, m1 = col0(rank), col0(rank + 1)
        full = np.asfortranarray(rhs.copy())                  # column-major like the device buffer: a slice is contiguous
        y = np.linalg.solve(L, full[:, m0:m1])
        full[:, m0:m1] = np.linalg.solve(L.T, y)              # in-place solve of the own slice
        for r in range(world):                                # one broadcast per owner, in place (grouped on the device)
            t = torch.from_numpy(np.ascontiguousarray(full[:, col0(r):col0(r + 1)].T))
            dist.broadcast(t, src=r)
            full[:, col0(r):col0(r + 1)] = t.numpy().T
        ref = np.linalg.solve(H, rhs)
        gathered = [None] * world
        dist.all_gather_object(gathered, full.tobytes())
        # residual scale: every rank must use the power of two of the GLOBAL max |R| (ncclMax all-reduce of the bit pattern)
        local_max = np.float32(3.0 if rank == 0 else 0.002)
        bits = torch.tensor([int(np.float32(local_max).view(np.uint32))], dtype=torch.int64)
        dist.all_reduce(bits, op=dist.ReduceOp.MAX)           # non-negative floats order like their bit patterns
        gmax = np.array([bits.item()], dtype=np.uint32).view(np.float32)[0]
        ret[rank] = (float(np.abs(full - ref).max()), all(g == gathered[0] for g in gathered), float(gmax))
    finally:
        dist.destroy_process_group()


def test_column_sharded_solve_and_global_scale_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_solve_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world
    for err, identical, gmax in ret.values():
        assert err < 1e-10 and identical and gmax == 3.0      # same bytes on every rank; the global maximum everywhere
