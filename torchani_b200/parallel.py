"""Multi-GPU sharding of one energy+force step (one process per GPU, torch.distributed/NCCL).

The reference has no distributed runtime at all; the path shards naturally
because AEV rows and atomic energies are independent per central atom:

* every rank holds the full (species, coords, cell) -- 120 KB at 10 k atoms -- and builds the
  same bucket grid (deterministic, so all ranks agree on the bucket-sorted order),
* rank r owns the bucket-sorted atoms ``[n*r/W, n*(r+1)/W)``: it runs neighbour search + AEV +
  MLP + AEV-backward for those central atoms only and scatters dE_owned/dx_j into a
  full-length gradient buffer (contributions land on non-owned neighbours too),
* ONE all-reduce(sum) of ``3N + C`` float64 values (gradient + per-conformer energies) over
  NVLink completes the step.  Nothing else is exchanged.
"""
from __future__ import annotations

import typing as tp

import torch
import torch.distributed as dist
from torch import Tensor

from .engine import Engine, StepResult


def shard_bounds(n: int, rank: int, world: int) -> tp.Tuple[int, int]:
    """Slice of bucket-sorted positions owned by ``rank`` (same formula as Engine.step)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def pack_partials(grad: Tensor, energies: Tensor) -> Tensor:
    """[grad (3N) | energies (C)] as one float64 message."""
    return torch.cat([grad.reshape(-1).to(torch.float64), energies.reshape(-1).to(torch.float64)])


def unpack_totals(buf: Tensor, n_conf: int, n_per_conf: int) -> tp.Tuple[Tensor, Tensor]:
    n3 = n_conf * n_per_conf * 3
    return buf[:n3].view(n_conf, n_per_conf, 3), buf[n3:n3 + n_conf]


def allreduce_partials(grad: Tensor, energies: Tensor, group=None) -> tp.Tuple[Tensor, Tensor]:
    """Sum the per-rank partial gradients / energies with a single collective."""
    buf = pack_partials(grad, energies)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return unpack_totals(buf, grad.shape[0], grad.shape[1])


class ShardedEngine:
    """Runs ``Engine.step`` on this rank's slice and all-reduces the partial results."""

    def __init__(self, engine: Engine, group=None):
        self.engine = engine
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1

    def step(self, species: Tensor, coords: Tensor, cell: tp.Optional[Tensor] = None, pbc: bool = False
             ) -> tp.Tuple[Tensor, Tensor]:
        """-> (energies f64 (C,), dE/dcoords f64 (C, A, 3)), identical on every rank."""
        res: StepResult = self.engine.step(species, coords, cell, pbc, want_grad=True,
                                           shard=(self.rank, self.world))
        grad, energies = allreduce_partials(res.grad, res.energies, self.group)
        return energies, grad
