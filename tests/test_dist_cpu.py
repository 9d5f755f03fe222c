"""CPU-only (gloo, world_size 2): the host-side logic of the multi-GPU path (dots_socp_b200/dist.py).

No kernels run here; what is checked is the sharding arithmetic and the exchange plumbing that the GPU ranks use:
slab / mode partition, the rank-local transform matrices (sharded transforms == full transforms), neighbour halo
shifts, slab gathers and the deterministic cross-rank sums."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dots_socp_b200 import dist as dd
from dots_socp_b200.engine import time_basis


def test_partition_covers_all_levels_once():
    for n_time in (6, 7, 15, 31, 63, 127):
        for world in (1, 2, 4, 8):
            if (world - 1) * -(-(n_time + 1) // world) >= n_time + 1:
                continue
            parts = [dd.partition(n_time, r, world) for r in range(world)]
            lv = np.concatenate([np.arange(p.lvl_begin, p.lvl_end) for p in parts])
            assert np.array_equal(lv, np.arange(n_time + 1))
            st = np.concatenate([np.arange(p.lvl_begin, p.t_end) for p in parts])
            assert np.array_equal(st, np.arange(n_time))
            assert all(p.m_pad in dd.MODE_PADS and p.m_pad >= p.n_modes for p in parts)
    with pytest.raises(ValueError):
        dd.partition(2, 0, 8)


@pytest.mark.parametrize("n_time,world", [(7, 2), (31, 4), (63, 8), (20, 3)])
def test_sharded_time_transforms_equal_the_full_ones(n_time, world):
    """hat = Q^T rhs restricted to a rank's modes, and phi = Q hat assembled from the gathered per-rank solutions."""
    Q, _ = time_basis(n_time)
    rng = np.random.default_rng(0)
    V = 5
    rhs = rng.standard_normal((n_time + 1, V))
    hat_full = Q.T @ rhs                                    # (modes, V)
    parts = [dd.partition(n_time, r, world) for r in range(world)]
    gathered = []
    for p in parts:
        qf, qb, n_out = dd.transform_matrices(Q, p)
        rhs_pad = np.zeros((qf.shape[0], V)); rhs_pad[:n_time + 1] = rhs
        hat_loc = (rhs_pad.T @ qf)                          # (V, m_pad) == the forward GEMM of this rank
        assert np.allclose(hat_loc[:, :p.n_modes], hat_full[p.lvl_begin:p.lvl_end].T, atol=1e-13)
        assert np.all(hat_loc[:, p.n_modes:] == 0)
        gathered.append(hat_loc)
    phi_full = Q @ hat_full
    for p in parts:
        qf, qb, n_out = dd.transform_matrices(Q, p)
        A = np.concatenate(gathered, axis=1)                # (V, world*m_pad): gathered layout seen per vertex
        phi_loc = A @ qb                                    # (V, nb)
        assert np.allclose(phi_loc[:, :n_out].T, phi_full[p.lvl_begin:p.lvl_begin + n_out], atol=1e-12)
        assert n_out == min(n_time + 1, p.lvl_end + 1) - p.lvl_begin


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_time):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = dd.Comm()
        assert comm.enabled and comm.rank == rank and comm.world == world
        part = dd.partition(n_time, rank, world)
        V = 3
        # slab store with one halo step in front, filled with the global step index
        mu = dd.SlabStore(part.lvl_begin - 1, part.t_end, (V,), "cpu")
        for t in range(part.lvl_begin, part.t_end):
            mu.level(t)[:] = float(t)
        # halo: last owned step -> next rank's step lvl_begin-1
        send = mu.level(part.t_end - 1).clone().reshape(1, V) if rank + 1 < world else None
        recv = torch.zeros(1, V, dtype=torch.float64) if rank > 0 else None
        comm.shift(send_next=send, recv_prev=recv)
        if rank > 0:
            mu.level(part.lvl_begin - 1).copy_(recv[0])
            assert float(mu.level(part.lvl_begin - 1)[0]) == part.lvl_begin - 1
        # backward halo: first owned level -> previous rank's level lvl_end
        cn = dd.SlabStore(part.lvl_begin, part.lvl_end + 1, (2,), "cpu")
        for lv in range(part.lvl_begin, part.lvl_end):
            cn.level(lv)[:] = 100.0 + lv
        send = cn.level(part.lvl_begin).clone() if rank > 0 else None
        recv = torch.zeros(2, dtype=torch.float64) if rank + 1 < world else None
        comm.shift(send_prev=send, recv_next=recv)
        if rank + 1 < world:
            assert float(recv[0]) == 100.0 + part.lvl_end
        # gather of the owned slabs reproduces the global field on every rank
        full = dd.gather_levels(comm, part, mu, n_time, owned_hi=part.t_end)
        assert torch.equal(full[:, 0], torch.arange(n_time, dtype=torch.float64))
        # all_gather_into with equal chunks (the rhs / hat exchange)
        rhs = torch.zeros(world * part.chunk, V, dtype=torch.float64)
        rhs[rank * part.chunk:(rank + 1) * part.chunk] = rank + 1.0
        comm.all_gather_into(rhs, rhs[rank * part.chunk:(rank + 1) * part.chunk].clone())
        for r in range(world):
            assert torch.all(rhs[r * part.chunk:(r + 1) * part.chunk] == r + 1.0)
        # deterministic sum in rank order
        tot = comm.sum_in_rank_order(np.array([rank + 1.0, 0.1 * (rank + 1)]), "cpu")
        assert tot[0] == sum(range(1, world + 1))
        comm.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_time", [(2, 7), (2, 6)])
def test_exchanges_with_gloo(world, n_time):
    mp.spawn(_worker, args=(world, _free_port(), n_time), nprocs=world, join=True)
