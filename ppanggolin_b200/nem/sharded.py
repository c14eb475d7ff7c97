"""Stepwise EM driver: one NEM instance on one GPU, or row-sharded over the ranks of a ``torch.distributed`` group.

Sharding: families (rows) are split into contiguous blocks, one per rank; every rank keeps its
block of the bit-packed matrix in HBM.  Per EM iteration there are two exchange steps:

* M-step: each rank accumulates the column statistics of its rows (``nemb200_mstep_accumulate``); the small K x D
  statistics (int32 one-hot counts, fp64 fuzzy sums, class sizes) are all-reduced (NCCL over NVLink) and every rank
  derives identical parameters (``nemb200_mstep_finalize``).
* E-step: each rank computes the log-densities of its rows (``nemb200_density``); they are all-gathered (N x K fp64,
  4.8 MB for 200 000 x 3) and every rank runs the same level-scheduled Gauss-Seidel sweep over all rows, so the
  posteriors are bit-identical on every rank and identical to the single-GPU result -- no block-Jacobi approximation
  at shard boundaries.

The reference has no counterpart (it is single-threaded C, one process per instance, partition.py:505,837); the loop
itself follows NEM/nem_alg.c:1174-1193 and :1783-1938.

``backend`` abstracts the kernels so that the exchange logic can be exercised on CPU with gloo in the tests; the
product backend is :class:`CudaBackend` (the C ABI).  There is no CPU backend in this package.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np


def shard_rows(N: int, world: int, rank: int):
    """Contiguous row block of ``rank``: sizes differ by at most one row."""
    base, rem = divmod(N, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class CudaBackend:
    """Kernels through libnemb200.so (engine.Plan)."""

    def __init__(self, bits_local, D: int, N_total: int, K: int, free_dispersion: bool, flags: int, graph, peer_group=None,
                 row_offset: int = 0):
        """``peer_group``: a torch.distributed group -> the two exchange steps of an iteration go over NVLink peer memory
        inside the kernels (nemb200_peer_*) instead of NCCL collectives; None -> single GPU, or NCCL in ShardedEM."""
        import torch
        from . import engine
        self.torch, self.engine = torch, engine
        self.bits = bits_local
        self.N_local = int(bits_local.shape[0])
        self.N, self.K, self.D = N_total, K, D
        self.row_offset = row_offset
        self.peer = peer_group is not None
        if self.peer:
            flags |= engine.PEER_EXCHANGE
        self.plan = engine.Plan(self.N_local, N_total, D, int(bits_local.shape[1]), K, free_dispersion, flags)
        self.graph = graph  # engine.DeviceInput carrying the full CSR + schedule (replicated), or None
        dev = bits_local.device
        if self.peer:
            import torch.distributed as dist
            world, rank = dist.get_world_size(peer_group), dist.get_rank(peer_group)
            mine = self.plan.peer_export()
            allx = [None] * world
            dist.all_gather_object(allx, mine, group=peer_group)
            self.plan.peer_attach(world, rank, [h for h, _ in allx], [o for _, o in allx])
            dist.barrier(group=peer_group)
            self.logf_full = self.plan.peer_logf()
        else:
            self.logf_full = torch.empty((N_total, K), dtype=torch.float64, device=dev)
        self.T = [torch.zeros((N_total, K), dtype=torch.float32, device=dev) for _ in range(2)]
        self.launches = 0
        self.density_ms, self.density_launches = 0.0, 0

    def init_params(self):
        self.plan.init_params(); self.launches += 1

    def accumulate(self, T_local):
        self.plan.mstep_accumulate(self.bits, T_local); self.launches += 4   # classify, scatter, one-hot counts, fuzzy sums

    def stats(self):
        return self.plan.stats_blocks()   # two contiguous blocks: int32 counts + class sizes, fp64 fuzzy sums + sizes

    def finalize(self):
        self.plan.mstep_finalize(); self.launches += 2 if self.peer else 1   # peer mode: + the signal-and-wait kernel

    def density(self, logf_local):
        if self.peer:   # stores into every rank's buffer and publishes the epoch: the all-gather, fused
            self.plan.density_peers(self.bits, self.row_offset); self.launches += 2   # density + shard push (its last CTA signals)
        else:
            self.plan.density(self.bits, logf_local); self.launches += 1

    def sweep(self, T_old, T_new, beta_now, maxdiff=None):
        """maxdiff: optional zeroed int32 device word receiving max |T_new - T_old| as float bits (a pipelined driver
        reads it one iteration late); None uses the plan's own word, read with read_flags()."""
        if self.peer:
            self.plan.peer_wait_densities()   # no launch: the sweep's CTAs check the flags themselves
        self.plan.sweep(self.N, self.logf_full, T_old, T_new, self.graph, beta_now, maxdiff); self.launches += 1

    def read_flags(self):
        """(max |T_new - T_old|, status) -- the only per-iteration device->host read."""
        return self.plan.read_flags()

    def criteria(self, T, beta):
        return self.plan.criteria(self.N, self.logf_full, T, self.graph, beta)

    def run_device_loop(self, beta: float, itermax: int, cv_th: float):
        """The whole run with the device-resident loop (peer mode only): no host synchronisation between iterations, the
        convergence test runs on the replicated posteriors of every rank.  Returns (T, crit, n_iter, converged, status)."""
        res = self.plan.nem_sharded(self.bits, self.row_offset, self.graph, beta, itermax, cv_th, self.T[0])
        self.density_ms, self.density_launches = self.density_ms + res.density_ms, self.density_launches + res.density_launches
        self.launches += 5 + 5 * (res.n_iter + 1) + 3     # see csrc/nemb200_batch.inc: start, per iteration, criteria + collect
        return self.T[0], list(res.crit), int(res.n_iter), bool(res.converged), int(res.status)


@dataclass
class EmOutcome:
    T: object
    crit: list
    n_iter: int
    converged: bool
    status: int


class ShardedEM:
    """EM loop over a backend; ``group`` is a torch.distributed process group or None (single GPU)."""

    def __init__(self, backend, N: int, K: int, rank: int = 0, world: int = 1, group=None):
        self.b, self.N, self.K, self.rank, self.world, self.group = backend, N, K, rank, world, group
        self.lo, self.hi = shard_rows(N, world, rank)
        self.cur = 0

    # -- exchange steps ------------------------------------------------------------------------------------------
    def _allreduce_stats(self):
        if self.world == 1 or getattr(self.b, "peer", False):   # peer mode: fused into nemb200_mstep_finalize
            return
        import torch.distributed as dist
        for t in self.b.stats():
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def _allgather_logf(self):
        if self.world == 1 or getattr(self.b, "peer", False):   # peer mode: fused into the density kernel
            return
        import torch
        import torch.distributed as dist
        full = self.b.logf_full
        spans = [shard_rows(self.N, self.world, r) for r in range(self.world)]
        rows = max(hi - lo for lo, hi in spans)
        if all((hi - lo) == rows for lo, hi in spans):
            dist.all_gather_into_tensor(full, full[self.lo:self.hi].clone(), group=self.group)
            return
        # uneven shards (N not a multiple of the world size): gather equal-size padded blocks, then unpack
        if getattr(self, "_pad", None) is None:
            self._pad = (torch.zeros((rows, self.K), dtype=full.dtype, device=full.device),
                         torch.zeros((self.world, rows, self.K), dtype=full.dtype, device=full.device))
        mine, gathered = self._pad
        mine[: self.hi - self.lo].copy_(full[self.lo:self.hi])
        dist.all_gather_into_tensor(gathered.view(self.world * rows, self.K), mine, group=self.group)
        for r, (lo, hi) in enumerate(spans):
            if r != self.rank:
                full[lo:hi].copy_(gathered[r, : hi - lo])

    # -- algorithm -----------------------------------------------------------------------------------------------
    def e_step(self, beta_now: float):
        b = self.b
        b.density(b.logf_full[self.lo:self.hi])
        self._allgather_logf()
        b.sweep(b.T[self.cur], b.T[1 - self.cur], beta_now)
        self.cur = 1 - self.cur

    def m_step(self):
        b = self.b
        b.accumulate(b.T[self.cur][self.lo:self.hi])
        self._allreduce_stats()
        b.finalize()

    def start(self, beta: float):
        """nem_alg.c:2023-2065 with Needinit = 1: blind sweep (beta = 0), then a sweep with beta."""
        self.b.init_params()
        self.b.T[self.cur].zero_()
        self.e_step(0.0)
        if beta != 0.0:
            b = self.b
            b.sweep(b.T[self.cur], b.T[1 - self.cur], beta)
            self.cur = 1 - self.cur

    def iterate(self, beta: float):
        """One EM iteration (nem_alg.c:1844-1868); returns (maxdiff, status)."""
        self.m_step()
        self.e_step(beta)
        return self.b.read_flags()

    def run(self, beta: float, itermax: int = 100, cv_th: float = 0.01) -> EmOutcome:
        if getattr(self.b, "peer", False) and hasattr(self.b, "run_device_loop") and self.world > 1:
            T, crit, it, converged, status = self.b.run_device_loop(beta, itermax, cv_th)
            return EmOutcome(T, crit, it, converged, status)
        self.start(beta)
        it, converged, status = 0, False, 0
        while it < itermax and not converged and status == 0:
            it += 1
            md, status = self.iterate(beta)
            converged = md < cv_th
        crit = self.b.criteria(self.b.T[self.cur], beta) if status == 0 else [float("nan")] * 5
        if status == 0 and crit[0] == float("-inf"):
            status = 2   # NEMB200_STS_INFINITE: the reference returns STS_E_INFINITE and writes no .uf (nem_alg.c:2482-2487)
        return EmOutcome(self.b.T[self.cur], crit, it, converged, status)
