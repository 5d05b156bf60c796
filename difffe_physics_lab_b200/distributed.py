"""Batch sharding of independent FEM solves over the GPUs of one box.

The hot path shards ONLY over the batch: every (f, kappa) sample is an independent linear system on a
mesh that is replicated on every rank, so the solve itself needs no collective and a single mesh never
leaves one GPU.  The one collective of the path is the sum all-reduce of shared-parameter gradients and
losses — a handful of float64 numbers per optimisation step — issued through ``torch.distributed``
(NCCL over NVLink on GPUs, gloo in the CPU tests).  One process per GPU; the rendezvous is the
launcher's (``torchrun``), this module never creates process groups.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_total: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of ``n_total`` samples for ``rank``: sizes differ by at most one
    and the shards tile [0, n_total) in rank order."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    base, extra = divmod(n_total, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t: torch.Tensor, rank: Optional[int] = None, world_size: Optional[int] = None) -> torch.Tensor:
    """View of this rank's rows of a batch tensor (dim 0 = samples)."""
    lo, hi = shard_bounds(t.shape[0], rank, world_size)
    return t[lo:hi]


def allreduce_sum_(values: Sequence[torch.Tensor]) -> Sequence[torch.Tensor]:
    """Sum-all-reduce several small float64 tensors with ONE collective (they are packed into one buffer,
    reduced, and copied back in place).  No-op for a single process."""
    _, w = world()
    if w == 1 or not values:
        return values
    flat = torch.cat([v.detach().reshape(-1).to(torch.float64) for v in values])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for v in values:
        n = v.numel()
        v.detach().copy_(flat[off:off + n].reshape(v.shape))
        off += n
    return values


def sharded_loss_and_grad(solver_factory, kappa: torch.Tensor, f_local: torch.Tensor, u_data_local: torch.Tensor,
                          n_total: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """One step of the kappa inverse-problem sweep (BASELINE config 5; examples/poisson_1d_demo.py:104-110
    batched): mean-squared mismatch over ALL ``n_total`` samples and its gradient w.r.t. the shared ``kappa``.

    ``solver_factory(kappa)`` returns the module to call on this rank's shard ``f_local``
    (``DifferentiableFESolver(mesh, kappa=kappa)``).  Forward + adjoint run on the local shard only; the
    partial loss and partial gradient are summed over ranks with one all-reduce of ``1 + kappa.numel()``
    doubles.  Returns (loss, dloss/dkappa), identical on every rank.
    """
    k = kappa.detach().clone().requires_grad_(True)
    u = solver_factory(k)(f_local)
    n_nodes = u.shape[-1]
    loss_local = ((u - u_data_local) ** 2).sum() / (n_total * n_nodes)
    (g_local,) = torch.autograd.grad(loss_local, k)
    loss = loss_local.detach().clone()
    allreduce_sum_([loss, g_local])
    return loss, g_local
