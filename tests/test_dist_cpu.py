"""Gene-shard communicator on CPU (gloo, world_size 2): the N > 1 host-side plumbing of the hot path."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scgeneclust_b200.dist import GeneShardComm, shard_bounds


def test_shard_bounds_cover_and_align():
    for total, world in [(20000, 8), (30000, 8), (13714, 2), (100, 4), (64, 3)]:
        b = shard_bounds(total, world)
        assert b[0][0] == 0 and b[-1][1] == total and len(b) == world
        for (lo, hi), (lo2, _) in zip(b, b[1:]):
            assert hi == lo2 and lo <= hi
        assert all(lo % 64 == 0 or lo == total for lo, _ in b)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = GeneShardComm(total)
        lo, hi = comm.gene_lo, comm.gene_hi
        full = torch.arange(total * 3, dtype=torch.float32).reshape(total, 3)
        got = comm.all_gather_rows(full[lo:hi].clone())
        assert torch.equal(got, full)
        # cell-space panel = sum over gene shards
        rs = np.random.RandomState(0)
        A = torch.from_numpy(rs.standard_normal((40, total)).astype(np.float32))
        Q = torch.from_numpy(rs.standard_normal((total, 4)).astype(np.float32))
        part = A[:, lo:hi].double() @ Q[lo:hi].double()
        comm.all_reduce_sum(part)
        assert torch.allclose(part, A.double() @ Q.double(), rtol=0, atol=1e-9)
        # svd_flip merge: largest magnitude, first (lowest gene index) on ties
        v = torch.tensor([1.0, -3.0, 2.0]) if rank == 0 else torch.tensor([-2.0, 3.0, -2.0])
        m = comm.absmax_merge(v)
        assert m.tolist() == [-2.0, -3.0, 2.0]
        # pair ranges
        n_pairs = 1001
        plo, phi = comm.split_range(n_pairs)
        vec = torch.arange(n_pairs, dtype=torch.float64)
        assert torch.equal(comm.all_gather_ranges(vec[plo:phi].clone(), n_pairs), vec)
        # row-sliced cell-space panel: reduce-scatter of the partial sums, all-gather of the processed chunks
        n_rows = 41
        chunk = comm.row_chunk(n_rows)
        assert chunk == 21
        partial = torch.zeros((chunk * world, 4), dtype=torch.float32)
        partial[:n_rows] = torch.from_numpy(rs.standard_normal((2, n_rows, 4)).astype(np.float32))[rank]
        both = torch.from_numpy(np.random.RandomState(0).standard_normal((40, total)).astype(np.float32))  # same stream as above
        mine = torch.empty((chunk, 4), dtype=torch.float32)
        expect = partial.clone()
        dist.all_reduce(expect)
        comm.reduce_scatter_rows(partial.clone(), mine)
        assert torch.equal(mine, expect[rank * chunk:(rank + 1) * chunk])
        gathered = torch.empty((chunk * world, 4), dtype=torch.float32)
        comm.all_gather_chunks(mine * 2.0, gathered)
        assert torch.equal(gathered, expect * 2.0)
        # flags raised on one rank reach every rank (rank-consistent errors)
        flag = torch.tensor([0, 1 if rank == 1 else 0], dtype=torch.int32)
        assert comm.all_reduce_max(flag).tolist() == [0, 1]
        comm.barrier()
    finally:
        dist.destroy_process_group()


def _worker_too_few_genes(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        try:
            GeneShardComm(64)              # 64 genes over 2 ranks: the second shard would be empty
        except ValueError as exc:
            assert "cannot be split" in str(exc)
        else:
            raise AssertionError("empty shard accepted")
    finally:
        dist.destroy_process_group()


def test_empty_shard_rejected_on_every_rank():
    mp.spawn(_worker_too_few_genes, args=(2, _free_port()), nprocs=2, join=True)


def test_gene_shard_comm_gloo_world2():
    mp.spawn(_worker, args=(2, _free_port(), 200), nprocs=2, join=True)
