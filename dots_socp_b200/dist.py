"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

How the path shards (reference coupling in time: utils/laplacian_inverse_socp.py:54-61, one-level
stencils: socp/solver_socp.py:884,892,934-940):

* **time slabs** for every streaming kernel: rank g owns the time levels ``[lvl_begin, lvl_end)`` (and the staggered
  steps ``[lvl_begin, min(lvl_end, nT))``) of all state arrays;
* **time modes** for the batched Laplacian solves: rank g factorises and sweeps only its ``n_modes`` modes;
* the two layouts meet in the time transforms: ``all_gather`` of the rhs slabs before the forward transform and of the
  per-rank solutions before the inverse one (every rank then sums over the full time/mode axis locally, in the same
  order as a single GPU would: results do not depend on the number of ranks);
* neighbour halos of ONE time level: (lam, A, lam_c, mu) of the last owned step go to the next rank after the vertex
  kernel, the side-1 corner norms of the first owned level go to the previous rank after the triangle kernel;
* residual sums: ``all_gather`` of the 8 local partial sums, added in rank order on the host (deterministic).

Everything here works on torch tensors of any device, so the same code is exercised with gloo/CPU in the tests.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

MODE_PADS = (8, 16, 32, 64, 96, 128)


def pad_modes(n: int) -> int:
    for p in MODE_PADS:
        if n <= p:
            return p
    raise ValueError(f"{n} time modes per rank exceed the sweep kernels' batch width (128)")


@dataclass
class Partition:
    """Slab / mode ownership of one rank."""
    n_time: int
    rank: int
    world: int
    chunk: int          # levels (= modes) per rank, last rank may own fewer
    lvl_begin: int
    lvl_end: int
    m_pad: int          # padded local mode count

    @property
    def t_end(self):
        return min(self.lvl_end, self.n_time)

    @property
    def n_levels(self):
        return self.lvl_end - self.lvl_begin

    @property
    def n_steps(self):
        return max(0, self.t_end - self.lvl_begin)

    @property
    def n_modes(self):
        return self.lvl_end - self.lvl_begin

    def owner_ranges(self):
        n = self.n_time + 1
        return [(r * self.chunk, min(n, (r + 1) * self.chunk)) for r in range(self.world)]


def partition(n_time: int, rank: int = 0, world: int = 1, min_pad: int = 0) -> Partition:
    n = n_time + 1
    chunk = -(-n // world)
    lo, hi = rank * chunk, min(n, (rank + 1) * chunk)
    if (world - 1) * chunk >= n:
        raise ValueError(f"{world} ranks cannot each own a time level of {n} levels")
    return Partition(n_time, rank, world, chunk, lo, hi, max(pad_modes(chunk), int(min_pad)))


def transform_matrices(Q: np.ndarray, part: Partition):
    """Host-built operands of the two time-transform GEMMs for this rank.

    qf (kf x m_pad): rows = all time levels (padded to 4), columns = this rank's modes.
    qb (kb x nb)   : rows = gathered modes in rank-major padded order, columns = the phi levels this rank writes
                     (its slab plus the halo level lvl_end when that exists)."""
    n = part.n_time + 1
    kf = (n + 3) & ~3
    qf = np.zeros((kf, part.m_pad))
    qf[:n, :part.n_modes] = Q[:, part.lvl_begin:part.lvl_end]
    n_out = min(n, part.lvl_end + 1) - part.lvl_begin
    nb = (n_out + 7) & ~7
    kb = part.world * part.m_pad
    qb = np.zeros((kb, nb))
    for r, (k0, k1) in enumerate(part.owner_ranges()):
        qb[r * part.m_pad:r * part.m_pad + (k1 - k0), :n_out] = Q[part.lvl_begin:part.lvl_begin + n_out, k0:k1].T
    return qf, qb, n_out


class Comm:
    """Thin wrapper around a torch.distributed process group (or nothing, for one rank)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.group = group
        self.rank = dist.get_rank(group) if self.enabled else 0
        self.world = dist.get_world_size(group) if self.enabled else 1

    # -- collectives -------------------------------------------------------------------------------------------
    def all_gather_into(self, out: torch.Tensor, mine: torch.Tensor):
        """out (world * n) <- concatenation of every rank's ``mine`` (n elements each); no-op copy on one rank."""
        if not self.enabled:
            if out.data_ptr() != mine.data_ptr():
                out.view(-1)[:mine.numel()].copy_(mine.reshape(-1))
            return
        self.dist.all_gather_into_tensor(out.view(-1), mine.reshape(-1), group=self.group)

    def shift(self, send_next=None, recv_prev=None, send_prev=None, recv_next=None):
        """Neighbour exchange along the rank line (no wrap-around).  Any argument may be None."""
        if not self.enabled:
            return
        d, ops = self.dist, []
        if send_next is not None and self.rank + 1 < self.world:
            ops.append(d.P2POp(d.isend, send_next, self._peer(self.rank + 1), self.group))
        if recv_prev is not None and self.rank > 0:
            ops.append(d.P2POp(d.irecv, recv_prev, self._peer(self.rank - 1), self.group))
        if send_prev is not None and self.rank > 0:
            ops.append(d.P2POp(d.isend, send_prev, self._peer(self.rank - 1), self.group))
        if recv_next is not None and self.rank + 1 < self.world:
            ops.append(d.P2POp(d.irecv, recv_next, self._peer(self.rank + 1), self.group))
        if ops:
            for req in d.batch_isend_irecv(ops):
                req.wait()

    def _peer(self, group_rank):
        return self.dist.get_global_rank(self.group, group_rank) if self.group is not None else group_rank

    def sum_in_rank_order(self, local: np.ndarray, device) -> np.ndarray:
        """Deterministic cross-rank sum of a few host doubles (gather, then add in rank order)."""
        if not self.enabled:
            return local
        mine = torch.as_tensor(local, dtype=torch.float64, device=device)
        out = torch.empty((self.world,) + tuple(mine.shape), dtype=torch.float64, device=device)
        self.dist.all_gather_into_tensor(out.view(-1), mine.view(-1), group=self.group)
        parts = out.cpu().numpy()
        total = parts[0].copy()
        for r in range(1, self.world):
            total += parts[r]
        return total

    def any_true(self, flag: bool, device) -> bool:
        """Logical OR of a host flag over the ranks (every rank gets the same answer)."""
        if not self.enabled:
            return bool(flag)
        t = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device=device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return bool(t.item() > 0.0)

    def barrier(self):
        if self.enabled:
            self.dist.barrier(group=self.group)


class SlabStore:
    """Level-indexed field of which a rank backs only ``[lo, hi)``; ``base_ptr`` is the virtual address of level 0."""

    def __init__(self, lo: int, hi: int, row_shape, device, dtype=torch.float64):
        self.lo, self.hi = lo, max(hi, lo + 1)
        self.row_shape = tuple(row_shape)
        self.row_elems = int(np.prod(self.row_shape))
        # two elements of padding behind the data: the bulk copies of the triangle kernel round their size up to 16 bytes and
        # may read one element past the last plane when the triangle count is odd
        n = (self.hi - self.lo) * self.row_elems
        self._flat = torch.zeros(n + 2, dtype=dtype, device=device)
        self.data = self._flat[:n].view((self.hi - self.lo,) + self.row_shape)

    @property
    def base_ptr(self):
        return self.data.data_ptr() - self.lo * self.row_elems * self.data.element_size()

    def level(self, lv):
        return self.data[lv - self.lo]

    def levels(self, lo, hi):
        return self.data[lo - self.lo:hi - self.lo]


def gather_levels(comm: Comm, part: Partition, store: SlabStore, n_levels_total: int, owned_hi=None, root=None):
    """Assemble the full (n_levels_total, ...) field from the owned slabs of all ranks: on every rank (``root`` None,
    all_gather) or on rank ``root`` only (gather; the other ranks get None)."""
    lo, hi = part.lvl_begin, part.lvl_end if owned_hi is None else owned_hi
    mine = torch.zeros((part.chunk,) + store.row_shape, dtype=store.data.dtype, device=store.data.device)
    if hi > lo:
        mine[:hi - lo] = store.levels(lo, hi)
    shape = (part.world * part.chunk,) + store.row_shape
    # Always the ring all_gather, also for a single consumer: its connections exist since setup (the IPC handle exchange
    # is an all_gather), whereas a first NCCL gather sets up point-to-point channels to every peer, which took 0.5 s at
    # 4 ranks and 1.3 s at 8.  The extra copies cross NVSwitch in a few ms; only ``root`` keeps and downloads the result.
    full = torch.empty(shape, dtype=store.data.dtype, device=store.data.device)
    comm.all_gather_into(full, mine)
    if root is not None and comm.rank != root:
        return None
    return full[:n_levels_total]
