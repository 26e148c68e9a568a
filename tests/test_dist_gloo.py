"""Multi-rank host logic of pygimli_b200.dist on CPU (gloo, world_size 2): shard arithmetic, the
padded all-gather layout of the potential exchange and the row permutation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pygimli_b200.dist import split_range, padded_range, chunk_width, max_chunk, row_order
from pygimli_b200.scheme import create_dd


def test_partitions_cover_everything():
    for n in (1, 7, 100, 1056, 9700):
        for world in (1, 2, 3, 8):
            seen = np.zeros(n, int)
            for r in range(world):
                a, b = split_range(n, world, r)
                seen[a:b] += 1
            assert np.all(seen == 1)
            seen[:] = 0
            w = max_chunk(n, world)
            for r in range(world):
                a, b = padded_range(n, world, r)
                assert b - a <= w and a % 2 == 0
                seen[a:b] += 1
            assert np.all(seen == 1)


def test_row_order_groups_current_dipoles():
    s = np.zeros((21, 3)); s[:, 0] = np.arange(21)
    sch = create_dd(s)
    o = row_order(sch)
    a = sch.a[o]
    assert np.all(np.diff(a) >= 0)
    lo, hi = split_range(sch.size, 4, 1)
    assert np.unique(a[lo:hi]).size < np.unique(sch.a).size / 2


def _worker(rank, world, port, n_nodes, n_src, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank owns the source columns padded_range(rank) of a global [n_nodes x n_src] block
    rng = np.random.default_rng(0)
    U = rng.standard_normal((n_nodes, n_src))
    w = max_chunk(n_src, world)
    a, b = padded_range(n_src, world, rank)
    send = torch.zeros(n_nodes * w, dtype=torch.float64)
    send[: n_nodes * (b - a)] = torch.from_numpy(np.ascontiguousarray(U[:, a:b]).ravel())     # pack
    recv = [torch.zeros(n_nodes * w, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(recv, send)
    got = np.zeros_like(U)
    for r in range(world):
        ra, rb = padded_range(n_src, world, r)
        got[:, ra:rb] = recv[r][: n_nodes * (rb - ra)].numpy().reshape(n_nodes, rb - ra)          # unpack
    ok = np.array_equal(got, U)
    # partial electrode matrices sum to the full one
    pm = torch.from_numpy(U[:5, a:b].sum(1).copy())
    dist.all_reduce(pm)
    ok = ok and np.allclose(pm.numpy(), U[:5].sum(1))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_potential_exchange_layout_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, 37, 11, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def _coverage_worker(rank, world, port, out):
    """row-sharded coverageDCtrans: local column sums of |J_ij dd_i| over this rank's (a,b)-ordered rows, one
    all-reduce, division by |mm| afterwards -- the exchange ShardedERT.coverage_trans performs with NCCL"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.ert_oracle import coverage_dc_trans
    rng = np.random.default_rng(3)
    D, M = 23, 17
    J, dd, mm = rng.standard_normal((D, M)), rng.standard_normal(D), rng.standard_normal(M) + 2.0
    perm = rng.permutation(D)                                           # stands in for row_order(scheme)
    lo, hi = split_range(D, world, rank)
    rows = perm[lo:hi]
    part = torch.from_numpy(np.abs(J[rows] * dd[rows, None]).sum(0))
    dist.all_reduce(part)
    got = part.numpy() / np.abs(mm)
    out[rank] = bool(np.allclose(got, coverage_dc_trans(J, dd, mm), rtol=1e-13, atol=0.0))
    dist.destroy_process_group()


def test_sharded_coverage_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_coverage_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
